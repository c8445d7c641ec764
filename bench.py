#!/usr/bin/env python
"""bench.py -- atom-steps/s of the ExaMiniMD hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

A "step" is one MD timestep (initial_integrate -> halo update or, every 20th step, exchange + sort + halo + binning +
neighbor build -> LJ force -> final_integrate, and the thermo reductions every 10th step exactly as the reference's
run loop times them) over the configuration BASELINE.json quotes the metric on: LJ fcc 2 048 000 atoms (in.lj with
`region 0 80 0 80 0 80`), cutoff 2.5, skin 0.3, half CSR list, one B200.  N > 1: that brick per GPU (weak scaling).

  value      atom-steps/s, state resident in HBM, CUDA events on the module stream, thermo passes at the deck's cadence
             INSIDE the timed region (the reference's Atomsteps/s spans them, src/examinimd.cpp:252-267,278);
             value_no_thermo = the same steps without them
  e2e        same metric through the host-buffer session API: every step copies x,v,f from pinned host memory to the
             device, advances one step and copies x,v,f (+ id,type after a re-sort) back
  roofline   dominant kernel (LJ force): algorithmic bytes of SURVEY.md 8(d) / CUDA-event duration of that kernel,
             against MEASURED_PEAKS.json hbm_gbs
  parity     N > 1 only, before the timed region: small decomposed LJ and SNAP runs compared by atom id with the
             single-rank CPU oracle (tests/mgpu_check.py)
  snap, weak16M, snap_strong   the other named configurations (configs[2], [3], [4]) as nested sections
  cpu_baseline / --impl reference: the reference's own CPU path (oracle/_ref/ExaMiniMD_ref_omp = the unmodified sources
             over an OpenMP Kokkos stand-in, all host cores) on the SAME region, a bounded number of steps
"""
from __future__ import annotations

import os

# every kernel of the library is loaded when its CUDA module is (default: at its first launch, ~1 ms each, which a short timed
# window would otherwise meet at its first re-neighboring / thermo step); must be set before CUDA initialises
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import argparse
import contextlib
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
DECK = REPO / "input" / "in.lj"
SNAP_DIR = REPO / "input" / "snap"
METRIC = "atom_steps_per_s_lj_2M_half_csr"
UNIT = "atom-steps/s"
BRICK = (80, 80, 80)
GRIDS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}  # the reference's minimum-surface rule (comm_mpi.cpp:58-89)


def workload(region, world):
    n = 4 * region[0] * region[1] * region[2]
    s = f"LJ fcc {n} atoms (in.lj, region {region[0]}x{region[1]}x{region[2]}), rc 2.5 + skin 0.3, half CSR list, re-neighbor every 20 steps, thermo every 10, newton off"
    if world > 1:
        s += f"; weak scaling: {4 * BRICK[0] * BRICK[1] * BRICK[2]} atoms per GPU"
    return s


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ CPU reference
REF_OMP = REPO / "oracle" / "_ref" / "ExaMiniMD_ref_omp"  # the UNMODIFIED reference over the host-only Kokkos stand-in (oracle/Makefile.ref)
PERF_RE = re.compile(r"^(\d+) (\d+) \| (\S+) (\S+) (\S+) (\S+) (\S+) \| (\S+) (\S+) (\S+) PERFORMANCE", re.M)


def cpu_env(cores):
    return dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close", OMP_PLACES="cores")


def edit_deck(src, region, nsteps):
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), Path(src).read_text())
    return re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)


def run_reference_cpu(region, nsteps, snap=False):
    """the reference's own CPU implementation of the path on all host cores: oracle/_ref when it was built (kind
    "reference"), else the oracle port (kind "port").  Returns the numbers of its PERFORMANCE line."""
    cores = os.cpu_count() or 1
    t0 = time.time()
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "in.deck").write_text(edit_deck(SNAP_DIR / "in.snap.W" if snap else DECK, region, nsteps))
        if snap:
            for f in SNAP_DIR.glob("*.snap*"):
                (td / f.name).write_bytes(f.read_bytes())
        flags = ["--comm-type", "SERIAL", "--neigh-type", "CSR"] + ([] if snap else ["--force-iteration", "NEIGH_HALF"])
        if REF_OMP.exists():
            exe, kind, what = REF_OMP, "reference", "oracle/_ref/ExaMiniMD_ref_omp (unmodified ExaMiniMD sources over the OpenMP Kokkos stand-in)"
        else:
            exe = REPO / "oracle" / "oracle_md_omp"
            if not exe.exists():
                subprocess.run(["make", "-C", str(REPO / "oracle"), "oracle_md_omp"], check=True, capture_output=True)
            kind, what = "port", "oracle/oracle_md_omp (OpenMP restatement of the reference)"
        out = subprocess.run([str(exe), "-il", "in.deck", *flags], capture_output=True, text=True, env=cpu_env(cores), check=True, cwd=td).stdout
    m = PERF_RE.search(out)
    return {"value": float(m.group(9)), "atoms": int(m.group(2)), "loop_s": float(m.group(3)), "cores": cores, "kind": kind, "what": what,
            "steps": nsteps, "wall_s": time.time() - t0}


def cpu_baseline(region, nsteps):
    r = run_reference_cpu(region, nsteps)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
            "sample": f"{r['what']}, same deck and region ({r['atoms']} atoms) x {nsteps} steps incl. one re-neighboring and the thermo passes, "
                      f"loop {r['loop_s']:.2f} s"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    g = GRIDS.get(world, (1, 1, 1))
    brick = tuple(args.region) if args.region else BRICK
    region = tuple(b * k for b, k in zip(brick, g))
    # bounded sample: the same system, at most 40 steps at N=1 (always across one re-neighboring), fewer for the larger
    # weak-scaling boxes so that the run ends within a few minutes on the host cores
    total = max(21, min(args.steps + args.warmup, 40 if world == 1 else 21))
    r = run_reference_cpu(region, total)
    ms = 1e3 * r["loop_s"] / total
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(region, world), "atoms_total": r["atoms"],
                       "sample": f"{total} steps of the workload (one re-neighboring, thermo every 10) on {r['cores']} host cores"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": f"{r['what']}, {r['atoms']} atoms x {total} steps, half CSR, loop {r['loop_s']:.2f} s"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


@contextlib.contextmanager
def stdout_to_stderr():
    """the C++ layer prints the reference's own banner lines (e.g. `CommSerial`, src/comm_types/comm_serial.cpp:42) on
    stdout; bench.py's stdout must hold exactly one JSON line, so they are sent to stderr while an App is created"""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ measurement
class Env:
    """process-wide state of one bench run: torch.distributed, the device, the barrier"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        self.gloo = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
            self.gloo = dist.new_group(backend="gloo")  # python objects of the parity check travel over gloo

    def barrier(self, app=None):
        if app is not None:
            app.sync()
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed(env, app, fn):
    """device time (ms, max over ranks) of fn() bracketed by barrier + synchronize, and the kernels it launched"""
    import examinimd_b200 as emd
    L = emd.lib()
    ms = C.c_float()
    env.barrier(app)
    l0 = app.launches()
    emd.check(L.emd_ctx_tic(app.ctx))
    fn()
    emd.check(L.emd_ctx_toc(app.ctx, C.byref(ms)))
    launches = app.launches() - l0
    env.barrier(app)
    return env.max_over_ranks(ms.value), launches


def lj_app(env, region, iteration="NEIGH_HALF"):
    import examinimd_b200 as emd
    argv = ["-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", iteration, "--comm-type", "MPI" if env.world > 1 else "SERIAL",
            "--region", *map(str, region)]
    with stdout_to_stderr():
        return emd.App(argv, device=env.local_rank)


def lj_throughput(env, app, K, W):
    """W warm-up steps, then K timed steps with and without the thermo passes.  The timed region starts one step before a
    re-neighboring, so K steps hold ceil(K/20) re-neighborings (exactly their share when K is a multiple of 20)."""
    n_atoms = app.get("N")
    rate = app.get("exchange_rate")
    app.run(W)
    app.thermo()  # (warm-up also meets the thermo kernels: CUDA loads a kernel at its first launch, ~1 ms)
    # the steps up to the re-neighboring boundary run like the timed ones (thermo at the deck's cadence), so that the warm-up
    # has met a thermo step (the force + energy launch) whatever W is
    app.run((rate - 1 - app.get("step")) % rate)
    t_ms, launches = timed(env, app, lambda: app.run(K))
    app.advance((rate - 1 - app.get("step")) % rate)
    t2_ms, _ = timed(env, app, lambda: app.advance(K))
    return {"value": n_atoms * K / (t_ms * 1e-3), "ms_per_step": t_ms / K, "launches": launches,
            "value_no_thermo": n_atoms * K / (t2_ms * 1e-3), "ms_per_step_no_thermo": t2_ms / K, "atoms": n_atoms}


def phase_split(app, nsteps):
    ph = app.advance_timed(nsteps)
    return {k: 1e3 * v / nsteps for k, v in ph.items()}


def main_section(env, args):
    """configs[1] (per GPU): throughput, dominant-kernel roofline, e2e, phases"""
    import examinimd_b200 as emd
    torch = env.torch
    L = emd.lib()
    P = C.c_void_p
    world = env.world
    brick = tuple(args.region) if args.region else BRICK
    grid = GRIDS[world]
    region = tuple(b * k for b, k in zip(brick, grid))
    if world > 1:
        dec = emd.Decomp()
        emd.check(L.emd_comm_decompose(world, env.rank, emd.vec3([r * 1.0 for r in region]), C.byref(dec)))
        assert tuple(dec.grid) == grid, (tuple(dec.grid), grid)
    half = 1 if args.iteration == "NEIGH_HALF" else 0
    W, K = max(args.warmup, 3), args.steps
    sampler = ClockSampler(env.local_rank)
    sampler.start()  # nvidia-smi needs ~0.1 s for its first sample: started before the warm-up, its samples cover the timed region
    app = lj_app(env, region, args.iteration)
    ctx = app.ctx
    thr = lj_throughput(env, app, K, W)
    clocks = sampler.stop()
    n_atoms = thr["atoms"]
    rate = app.get("exchange_rate")

    # ---- dominant kernel: LJ force, timed alone on the live state
    n_local, n_ghost = app.get("N_local"), app.get("N_ghost")
    total_neighs = app.get("total_neighs")
    nbar = total_neighs / n_local
    g = (n_local + n_ghost) / n_local
    lst = emd.NeighList(app.device_ptr("row_map"), None, app.device_ptr("neighs"), 1)
    tiles = app.device_ptr("tiles")  # the product path: tile lists (kernels/tiles.cu); 0 = generic list kernels in use
    ms = C.c_float()

    def force_call():
        if tiles:
            emd.check(L.emd_force_lj_compute_tiles(ctx, P(tiles), P(app.device_ptr("x")), P(app.device_ptr("type")),
                                                   P(app.device_ptr("f")), None))
        else:
            emd.check(L.emd_force_lj_compute(ctx, P(app.device_ptr("x")), P(app.device_ptr("type")), P(app.device_ptr("f")), n_local,
                                             n_local + n_ghost, C.byref(lst), half, 1))

    # the kernel streams the tile lists (> 300 MB at 2 M atoms) + x + f: more than the 126 MB L2, no flush needed
    fk_ms = []
    for _ in range(3):
        force_call()
    for _ in range(20):
        emd.check(L.emd_ctx_tic(ctx))
        force_call()
        emd.check(L.emd_ctx_toc(ctx, C.byref(ms)))
        fk_ms.append(ms.value)
    force_ms = statistics.mean(fk_ms)
    force_bytes = n_local * (8 + 4 * nbar + 28 * g + 48 * g + 24 * g)  # row_map + list + x,type + f RMW + zero-f
    peak, peak_kind = peaks()
    achieved = force_bytes / (force_ms * 1e-3) / 1e9
    b_lj = 208 + 100 * g + 52 * (g - 1) + 4 * nbar
    per_gpu = thr["value"] / world
    traffic, traffic_src = None, None
    tj = REPO / "profiles" / "lj_force_traffic.json"
    if tiles and world == 1 and not args.region and tj.exists():
        tr = json.loads(tj.read_text())
        traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": force_bytes,
                "kernel": "lj_tiles_kernel" if tiles else "lj_force_kernel<half> + zero_rows_kernel", "kernel_ms": force_ms, "peak_kind": peak_kind,
                "algorithmic_bytes_per_atom": force_bytes / n_local, "nbar": nbar, "g": g,
                "note": "the kernel is bound by the FP64 pipe (0.146 ms floor) and the shared-memory pipe, not by HBM (DESIGN.md 4.3); the HBM fraction is what BASELINE.json's metric asks for",
                "whole_step": {"B_LJ_bytes_per_atom_step": b_lj, "achieved_gbs": b_lj * per_gpu / 1e9, "frac": b_lj * per_gpu / 1e9 / peak,
                               "frac_no_thermo": b_lj * thr["value_no_thermo"] / world / 1e9 / peak}}

    # ---- e2e: host buffers in and out every step
    st = app.download()
    cap = int(n_local * 1.05) + 4096  # head-room: with CommMPI the owned-atom count drifts as atoms migrate between bricks

    def pinned(a, width):
        t = torch.zeros((cap, width) if width > 1 else (cap,), dtype=torch.from_numpy(a).dtype).pin_memory()
        t[: a.shape[0]] = torch.from_numpy(a)
        return t

    hx, hv, hf = pinned(st["x"], 3), pinned(st["v"], 3), pinned(st["f"], 3)
    hid, htype = pinned(st["id"], 1), pinned(st["type"], 1)
    Ke = min(K, 40)
    h2d = 72 * n_local * world      # whole job, every step: x, v, f of the owned atoms of every rank
    d2h = 72 * n_local * world
    d2h_rebuild = 8 * n_local * world

    def e2e_step():
        emd.check(L.emd_app_upload(app.handle, P(hx.data_ptr()), P(hv.data_ptr()), P(hf.data_ptr())))
        app.advance(1)
        resort = app.get("step") % rate == 0
        emd.check(L.emd_app_download(app.handle, P(hid.data_ptr()) if resort else None, P(htype.data_ptr()) if resort else None, None,
                                     P(hx.data_ptr()), P(hv.data_ptr()), P(hf.data_ptr())))

    for _ in range(3):
        e2e_step()
    env.barrier(app)
    t0 = time.perf_counter()
    emd.check(L.emd_ctx_tic(ctx))
    for _ in range(Ke):
        e2e_step()
    emd.check(L.emd_ctx_toc(ctx, C.byref(ms)))
    env.barrier(app)
    e2e_wall = time.perf_counter() - t0
    t_e = env.max_over_ranks(max(ms.value * 1e-3, e2e_wall))
    e2e = {"value": n_atoms * Ke / t_e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + d2h_rebuild / rate,
           "steps": Ke, "api": "emd_app_upload -> emd_app_advance(1) -> emd_app_download (pinned host x,v,f)"}

    # device-event phase split of 2*rate further steps (reference timers: force / neigh / comm / other)
    app.advance((-app.get("step")) % rate)
    phases = phase_split(app, 2 * rate)
    T, PE, KE = app.thermo()
    line = {"metric": METRIC, "value": thr["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": thr["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "value_no_thermo": thr["value_no_thermo"], "ms_per_step_no_thermo": thr["ms_per_step_no_thermo"],
            "config": {"workload": workload(region, world), "atoms_total": n_atoms, "atoms_per_gpu": n_atoms // world,
                       "ghosts_rank0": n_ghost, "neigh_entries_rank0": total_neighs,
                       "l2": "state + lists (>400 MB per GPU) exceed the 126 MB L2; no flush between steps",
                       "parallelism": "1 GPU" if world == 1 else
                       f"3-D domain decomposition {grid[0]}x{grid[1]}x{grid[2]} (CommMPI), halo exchange over NVLink every step, one process per GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": thr["launches"], "roofline": roofline, "phase_ms_per_step": phases,
            "thermo_after": {"T": T, "PE": PE, "E": PE + KE}}
    app.close()
    return line


def fp64_peak(env):
    """FP64 FMA peak measured in this run (DFMA loop on every SM, best of 10): the SNAP roofline's denominator"""
    import examinimd_b200 as emd
    ctx = emd.Context(env.local_rank)
    best, mean = C.c_double(), C.c_double()
    emd.check(emd.lib().emd_microbench_fp64(ctx.handle, 10, C.byref(best), C.byref(mean)))
    ctx.close()
    return {"tflops": best.value, "tflops_mean": mean.value, "how": "emd_microbench_fp64: 8 DFMA chains/thread, 8x256 threads per SM, best of 10, measured by this run"}


def snap_section(env, region, steps, name, decomposed):
    """SNAP tungsten (input/snap/in.snap.W, 2J=8), full list.  FP64-pipe roofline: flops of the formulation actually executed
    (adjoint ui/yi/duidrj/deidrj, SURVEY.md 8(d): 464 546 + n_in * 28 446 per atom-step) against the FP64 peak measured by this run."""
    import examinimd_b200 as emd
    L = emd.lib()
    n_expected = region[0] * region[1] * region[2]
    out = {"metric": name, "unit": UNIT,
           "workload": f"SNAP W (in.snap.W: sc 3.1803, 2J=8, rcut 4.73442 + skin 1.0, re-neighbor every step, newton on), region "
                       f"{region[0]}x{region[1]}x{region[2]} = {n_expected} atoms, full CSR list" +
                       (f", strong scaling over {env.world} GPUs (CommMPI)" if decomposed and env.world > 1 else "")}
    td = Path(tempfile.mkdtemp())
    (td / "in.deck").write_text(edit_deck(SNAP_DIR / "in.snap.W", region, steps))
    for f in SNAP_DIR.glob("*.snap*"):
        (td / f.name).write_bytes(f.read_bytes())
    with stdout_to_stderr():
        app = emd.App(["-il", str(td / "in.deck"), "--neigh-type", "CSR", "--comm-type", "MPI" if (decomposed and env.world > 1) else "SERIAL"],
                      device=env.local_rank)
    n = app.get("N")
    app.advance(3)
    t_ms, launches = timed(env, app, lambda: app.advance(steps))
    ph = phase_split(app, steps)
    snap = C.c_void_p(L.emd_app_device_ptr(app.handle, b"snap"))
    npairs = C.c_int()
    emd.check(L.emd_snap_info(snap, None, None, None, None, C.byref(npairs), None))
    n_local = app.get("N_local")
    T, _, _ = app.thermo()
    app.close()
    value = n * steps / (t_ms * 1e-3)
    n_in = npairs.value / max(n_local, 1)
    flops = 464546 + n_in * 28446
    pk = fp64_peak(env)
    force_ms = ph["force"]
    ranks = env.world if decomposed else 1
    out.update({"value": value, "ms_per_step": t_ms / steps, "steps": steps, "atoms": n, "n_gpus": ranks, "gpu_launches": launches,
                "phase_ms_per_step": ph, "T_after": T,
                "roofline": {"bound": "fp64", "formulation": "adjoint (ui, yi, duidrj, deidrj)", "n_inside_mean": n_in,
                             "flops_per_atom_step": flops, "achieved": flops * n_local / (force_ms * 1e-3) / 1e12, "peak": pk["tflops"],
                             "unit": "TFLOP/s per GPU", "frac": flops * n_local / (force_ms * 1e-3) / 1e12 / pk["tflops"],
                             "frac_whole_step": flops * value / ranks / 1e12 / pk["tflops"], "peak_kind": pk["how"],
                             "kernel": "ForceSNAP::compute (pair list + ui + yi + deidrj kernels)", "kernel_ms": force_ms}})
    return out


def weak16m_section(env, steps):
    """configs[3]: LJ fcc weak scaling with the 160^3 brick (16 384 000 atoms) per GPU"""
    brick = (160, 160, 160)
    grid = GRIDS[env.world]
    region = tuple(b * k for b, k in zip(brick, grid))
    app = lj_app(env, region)
    thr = lj_throughput(env, app, steps, 3)
    rate = app.get("exchange_rate")
    n_local, n_ghost, total = app.get("N_local"), app.get("N_ghost"), app.get("total_neighs")
    app.advance((-app.get("step")) % rate)
    phases = phase_split(app, rate)
    app.close()
    g, nbar = (n_local + n_ghost) / n_local, total / n_local
    b_lj = 208 + 100 * g + 52 * (g - 1) + 4 * nbar
    peak, kind = peaks()
    return {"metric": "atom_steps_per_s_lj_16M_per_gpu_half_csr", "unit": UNIT, "workload": workload(region, env.world).replace(
                f"{4 * BRICK[0] * BRICK[1] * BRICK[2]} atoms per GPU", "16384000 atoms per GPU"),
            "value": thr["value"], "ms_per_step": thr["ms_per_step"], "value_no_thermo": thr["value_no_thermo"], "steps": steps, "atoms": thr["atoms"],
            "n_gpus": env.world, "gpu_launches": thr["launches"], "phase_ms_per_step": phases,
            "whole_step": {"B_LJ_bytes_per_atom_step": b_lj, "frac": b_lj * thr["value"] / env.world / 1e9 / peak, "peak_kind": kind}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--region", type=int, nargs=3, default=None, help="override the per-GPU brick (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="only the headline configuration (no snap / weak16M / snap_strong sections)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the decomposed-run parity check before the timed region")
    ap.add_argument("--iteration", default="NEIGH_HALF", help="force iteration (debug; the headline config is NEIGH_HALF)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly one JSON line: everything native code prints on fd 1 while the job runs (the C++ layer's
    # banner lines, NCCL's version line at communicator creation) goes to stderr; the line is written to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    env = Env()
    if env.world not in GRIDS:
        raise SystemExit("bench.py: --gpus must be 1, 2, 4 or 8")

    parity = None
    if env.world > 1 and not args.no_parity:
        # the decomposed path proves itself before anything is timed: LJ half list across two re-neighborings and SNAP with its
        # reverse force fold, by atom id against the single-rank CPU oracle (1e-10 of the RMS)
        sys.path.insert(0, str(REPO / "tests"))
        import mgpu_check
        cases = [("lj", 12, 14, 14, 45, "NEIGH_HALF"), ("snap", 4, 4, 8, 6, "NEIGH_FULL")]
        res = [mgpu_check.parity(*c, group=env.gloo, device=env.local_rank) for c in cases]
        if env.rank == 0:
            parity = {"ok": all(r["ok"] for r in res), "err_x": max(r.get("err_x", 1.0) for r in res), "err_v": max(r.get("err_v", 1.0) for r in res),
                      "err_f": max(r.get("err_f", 1.0) for r in res), "tolerance": 1e-10, "against": "single-rank CPU oracle, by atom id",
                      "cases": res}
        env.barrier()

    line = main_section(env, args)
    if parity is not None:
        line["parity"] = parity
    if not args.no_extra and not args.region:
        extra_steps = max(5, min(args.steps, 20))
        try:
            if env.world == 1:
                line["snap"] = snap_section(env, (50, 50, 100), min(extra_steps, 10), "atom_steps_per_s_snap_W_250k", decomposed=False)
            line["weak16M"] = weak16m_section(env, 20)
            line["snap_strong"] = snap_section(env, (100, 100, 200), min(extra_steps, 10), "atom_steps_per_s_snap_W_2M_strong", decomposed=True)
        except Exception as e:  # the nested sections never take the headline down
            line.setdefault("extra_error", str(e))
    if env.rank == 0:
        if env.world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(tuple(args.region) if args.region else BRICK, 25)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
            if "snap" in line and not args.no_extra:
                try:
                    r = run_reference_cpu((20, 20, 40), 1, snap=True)
                    line["snap"]["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                                    "sample": f"{r['what']}, in.snap.W region 20x20x40 (16000 atoms) x 1 step, loop {r['loop_s']:.2f} s"}
                except Exception as e:
                    line["snap"]["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if env.dist:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
