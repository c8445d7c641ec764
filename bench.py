#!/usr/bin/env python
"""bench.py -- atom-steps/s of the ExaMiniMD LJ hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one MD timestep (initial_integrate -> halo update or, every 20th step, exchange + sort +
halo + binning + neighbor build -> LJ force -> final_integrate) over the configuration BASELINE.json
quotes the metric on: LJ fcc 2 048 000 atoms (in.lj with `region 0 80 0 80 0 80`), cutoff 2.5, skin
0.3, half CSR list, one B200.  The timed region starts one step before a re-neighboring, so K steps hold
ceil(K/20) re-neighborings (exactly their share when K is a multiple of 20).

  value     atom-steps/s, state resident in HBM, timed with CUDA events on the module stream
  e2e       same metric through the host-buffer session API: every step copies x,v,f from pinned
            host memory to the device, advances one step and copies x,v,f (+ id,type after a
            re-sort) back
  roofline  dominant kernel (LJ force incl. its fused zero-f): algorithmic bytes of SURVEY.md 8(d)
            / CUDA-event duration of that kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the CPU oracle restatement of the reference (OpenMP, all host
            cores) on a bounded sample of the same workload (in.lj 256 000 atoms)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
DECK = REPO / "input" / "in.lj"
METRIC = "atom_steps_per_s_lj_2M_half_csr"
UNIT = "atom-steps/s"
WORKLOAD = "LJ fcc 2048000 atoms (in.lj, region 80^3), rc 2.5 + skin 0.3, half CSR list, re-neighbor every 20 steps, newton off"


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------ CPU baseline
REF_OMP = REPO / "oracle" / "_ref" / "ExaMiniMD_ref_omp"  # the UNMODIFIED reference over the host-only Kokkos stand-in (oracle/Makefile.ref)
PERF_RE = re.compile(r"^(\d+) (\d+) \| (\S+) (\S+) (\S+) (\S+) (\S+) \| (\S+) (\S+) (\S+) PERFORMANCE", re.M)


def cpu_env(cores):
    return dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close", OMP_PLACES="cores")


def run_reference_cpu(region, nsteps, threads=None):
    """the reference's own CPU implementation of the path on all host cores: oracle/_ref when it was built (kind
    "reference"), else the oracle port (kind "port").  Returns the numbers of its PERFORMANCE line."""
    cores = threads or os.cpu_count() or 1
    t0 = time.time()
    if REF_OMP.exists():
        import tempfile
        with tempfile.TemporaryDirectory() as td:  # the reference has no --region/--nsteps flags: edit the deck's region/run lines
            deck = Path(td) / "in.deck"
            txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), DECK.read_text())
            deck.write_text(re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt))
            out = subprocess.run([str(REF_OMP), "-il", str(deck), "--comm-type", "SERIAL", "--neigh-type", "CSR", "--force-iteration",
                                  "NEIGH_HALF"], capture_output=True, text=True, env=cpu_env(cores), check=True).stdout
        kind, what = "reference", "oracle/_ref/ExaMiniMD_ref_omp (unmodified ExaMiniMD sources over the OpenMP Kokkos stand-in)"
    else:
        exe = REPO / "oracle" / "oracle_md_omp"
        if not exe.exists():
            subprocess.run(["make", "-C", str(REPO / "oracle"), "oracle_md_omp"], check=True, capture_output=True)
        out = subprocess.run([str(exe), "-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--region",
                              *map(str, region), "--nsteps", str(nsteps)], capture_output=True, text=True, env=cpu_env(cores), check=True).stdout
        kind, what = "port", "oracle_md_omp (OpenMP restatement of the reference)"
    m = PERF_RE.search(out)
    return {"value": float(m.group(9)), "atoms": int(m.group(2)), "loop_s": float(m.group(3)), "cores": cores, "kind": kind, "what": what,
            "wall_s": time.time() - t0}


def cpu_baseline(nsteps):
    r = run_reference_cpu((40, 40, 40), nsteps)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
            "sample": f"{r['what']}, in.lj 256000 atoms x {nsteps} steps, half CSR, loop {r['loop_s']:.2f} s"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    r = run_reference_cpu(tuple(args.region) if args.region else (40, 40, 40), total)
    ms = 1e3 * r["loop_s"] / total
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "256000-atom in.lj sample of the workload per step (CPU-bounded)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": f"{r['what']}, in.lj 256000 atoms x {total} steps, half CSR"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


import contextlib


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


# ------------------------------------------------------------------------------------- SNAP
SNAP_DIR = REPO / "input" / "snap"


def snap_deck(td, region, nsteps):
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), (SNAP_DIR / "in.snap.W").read_text())
    txt = re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)
    (td / "in.deck").write_text(txt)
    for f in SNAP_DIR.glob("*.snap*"):
        (td / f.name).write_bytes(f.read_bytes())
    return td / "in.deck"


def snap_section(device, steps=10):
    """BASELINE.json configs[2]: SNAP tungsten (input/snap/in.snap.W, 2J=8), full list, 250 000 atoms, one B200.
    FP64-pipe roofline: flops of the formulation actually executed (adjoint ui/yi/duidrj/deidrj, SURVEY.md 8(d):
    464 546 + n_in * 28 446 per atom-step) against the FP64 peak measured on this pool (profiles/r01_microbench_b200.json)."""
    import tempfile
    import examinimd_b200 as emd
    L = emd.lib()
    out = {"metric": "atom_steps_per_s_snap_W_250k", "unit": UNIT,
           "workload": "SNAP W (in.snap.W: sc 3.1803, 2J=8, rcut 4.73442 + skin 1.0, re-neighbor every step, newton on), region 50x50x100 = 250000 atoms, full CSR list"}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        with stdout_to_stderr():
            app = emd.App(["-il", str(snap_deck(td, (50, 50, 100), steps)), "--neigh-type", "CSR", "--comm-type", "SERIAL"], device=device)
        n = app.get("N")
        app.advance(3)
        ms = C.c_float()
        l0 = app.launches()
        emd.check(L.emd_ctx_tic(app.ctx))
        app.advance(steps)
        emd.check(L.emd_ctx_toc(app.ctx, C.byref(ms)))
        launches = app.launches() - l0
        ph = app.advance_timed(steps)
        snap = C.c_void_p(L.emd_app_device_ptr(app.handle, b"snap"))
        npairs = C.c_int()
        emd.check(L.emd_snap_info(snap, None, None, None, None, C.byref(npairs), None))
        T, _, _ = app.thermo()
        app.close()
    value = n * steps / (ms.value * 1e-3)
    n_in = npairs.value / n
    flops = 464546 + n_in * 28446
    fp64_peak, kind = 36.8, "measured (tools/microbench.cu on this pool, profiles/r01_microbench_b200.json)"
    mb = REPO / "profiles" / "r01_microbench_b200.json"
    if mb.exists():
        fp64_peak = json.loads(mb.read_text())["fp64_tflops_sustained"]
    force_ms = 1e3 * ph["force"] / steps
    out.update({"value": value, "ms_per_step": ms.value / steps, "steps": steps, "atoms": n, "gpu_launches": launches,
                "phase_ms_per_step": {k: 1e3 * v / steps for k, v in ph.items()}, "T_after": T,
                "roofline": {"bound": "fp64", "formulation": "adjoint (ui, yi, duidrj, deidrj)", "n_inside_mean": n_in,
                             "flops_per_atom_step": flops, "achieved": flops * n / (force_ms * 1e-3) / 1e12, "peak": fp64_peak,
                             "unit": "TFLOP/s", "frac": flops * n / (force_ms * 1e-3) / 1e12 / fp64_peak, "peak_kind": kind,
                             "kernel": "ForceSNAP::compute = snap_pairs + snap_ui + snap_yi + snap_deidrj", "kernel_ms": force_ms}})
    return out


def snap_cpu_baseline(nsteps=4):
    """the reference's own SNAP on the box's host cores: in.snap.W as shipped (256 atoms)"""
    import tempfile
    if not REF_OMP.exists():
        return None
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        deck = snap_deck(td, (4, 8, 8), nsteps)
        outp = subprocess.run([str(REF_OMP), "-il", str(deck), "--comm-type", "SERIAL", "--neigh-type", "CSR"], capture_output=True, text=True,
                              env=cpu_env(cores), check=True, cwd=td).stdout
    m = PERF_RE.search(outp)
    return {"value": float(m.group(9)), "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"oracle/_ref/ExaMiniMD_ref_omp, in.snap.W as shipped (256 atoms) x {nsteps} steps, loop {float(m.group(3)):.2f} s"}


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


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--region", type=int, nargs=3, default=None, help="override the lattice (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-snap", action="store_true", help="skip the SNAP (configs[2]) section of the single-GPU line")
    ap.add_argument("--iteration", default="NEIGH_HALF", help="force iteration (debug; the headline config is NEIGH_HALF)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly one JSON line: everything native code prints on fd 1 while the job runs (the C++ layer's
    # banner lines, NCCL's version line at communicator creation) goes to stderr; the line is written to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import numpy as np
    import torch
    import examinimd_b200 as emd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None

    W = max(args.warmup, 3)
    K = args.steps
    brick = tuple(args.region) if args.region else (80, 80, 80)
    # N > 1: weak scaling, ONE global system cut into `world` bricks of the single-GPU size by CommMPI (3-D domain
    # decomposition, NCCL halo exchange over NVLink every step).  The processor grid is the reference's minimum-surface
    # rule (comm_mpi.cpp:58-89), which for these boxes is (1,1,2), (1,2,2), (2,2,2).
    L = emd.lib()
    grid = (1, 1, 1)
    if world > 1:
        g = {2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
        if g is None:
            raise SystemExit("bench.py: --gpus must be 1, 2, 4 or 8")
        region = tuple(b * k for b, k in zip(brick, g))
        dec = emd.Decomp()
        emd.check(L.emd_comm_decompose(world, rank, emd.vec3([r * 1.0 for r in region]), C.byref(dec)))
        assert tuple(dec.grid) == g, (tuple(dec.grid), g)
        grid = g
    else:
        region = brick
    half = 1 if args.iteration == "NEIGH_HALF" else 0
    argv = ["-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", args.iteration, "--comm-type", "MPI" if world > 1 else "SERIAL",
            "--region", *map(str, region)]
    with stdout_to_stderr():
        app = emd.App(argv, device=local_rank)
    ctx = app.ctx
    n_atoms = app.get("N")  # global atom count

    def barrier():
        app.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    # ---- resident throughput ------------------------------------------------------------
    # nvidia-smi needs ~0.1 s to deliver its first sample: it is started before the warm-up so that its samples cover the
    # timed region (both run the same steps under the same load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    app.advance(W)
    # land on the step just BEFORE a re-neighboring: the K timed steps then hold ceil(K / rate) rebuilds -- exactly K / rate for a
    # multiple of the deck's cadence (20), and never fewer than their share for any other K
    rate = app.get("exchange_rate")
    app.advance((rate - 1 - app.get("step")) % rate)
    barrier()
    launches0 = app.launches()
    ms = C.c_float()
    emd.check(L.emd_ctx_tic(ctx))
    app.advance(K)
    emd.check(L.emd_ctx_toc(ctx, C.byref(ms)))
    launches = app.launches() - launches0
    barrier()
    clocks = sampler.stop()
    t_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    step_ms = float(t_ms.item()) / K
    value = n_atoms * K / (float(t_ms.item()) * 1e-3)

    # ---- dominant kernel: LJ force (with its fused zero-f), timed alone on the live state ----
    n_local, n_ghost = app.get("N_local"), app.get("N_ghost")
    total_neighs = app.get("total_neighs")
    nbar = total_neighs / n_local
    g = (n_local + n_ghost) / n_local
    lst = emd.NeighList(app.device_ptr("row_map"), None, app.device_ptr("neighs"), 1)
    P = C.c_void_p
    reps = 20
    tiles = app.device_ptr("tiles")  # the product path: tile lists (kernels/tiles.cu); 0 = generic list kernels in use

    def force_call():
        if tiles:
            emd.check(L.emd_force_lj_compute_tiles(ctx, P(tiles), P(app.device_ptr("x")), P(app.device_ptr("type")),
                                                   P(app.device_ptr("f")), None))
        else:
            emd.check(L.emd_force_lj_compute(ctx, P(app.device_ptr("x")), P(app.device_ptr("type")), P(app.device_ptr("f")), n_local,
                                             n_local + n_ghost, C.byref(lst), half, 1))

    # the kernel streams the tile lists (~330 MB at 2 M atoms) + x + f: more than the 126 MB L2, no flush needed
    fk_ms = []
    for _ in range(3):
        force_call()
    for _ in range(reps):
        emd.check(L.emd_ctx_tic(ctx))
        force_call()
        emd.check(L.emd_ctx_toc(ctx, C.byref(ms)))
        fk_ms.append(ms.value)
    force_ms = statistics.mean(fk_ms)
    force_bytes = n_local * (8 + 4 * nbar + 28 * g + 48 * g + 24 * g)  # row_map + list + x,type + f RMW + zero-f
    peak, peak_kind = peaks()
    achieved = force_bytes / (force_ms * 1e-3) / 1e9
    b_lj = 208 + 100 * g + 52 * (g - 1) + 4 * nbar
    value_per_gpu = value / world
    # DRAM bytes of one launch of this kernel on this workload from the committed `ncu --set full` capture (profiles/)
    traffic, traffic_src = None, None
    tj = REPO / "profiles" / "r01_lj_tiles_traffic.json"
    if tiles and world == 1 and not args.region and tj.exists():
        tr = json.loads(tj.read_text())
        traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": force_bytes,
                "limiter": "FP64 pipe + LSU, not HBM: 17 FP64 instructions per pair, every pair evaluated from both sides (ncu: FP64 pipe 48 % active, "
                           "DRAM 17 % of peak); the HBM fraction is reported because BASELINE.json's metric asks for it",
                "kernel": "lj_tiles_kernel" if tiles else "lj_force_kernel<half> + zero_rows_kernel", "kernel_ms": force_ms, "peak_kind": peak_kind,
                "algorithmic_bytes_per_atom": force_bytes / n_local, "nbar": nbar, "g": g,
                "whole_step": {"B_LJ_bytes_per_atom_step": b_lj, "achieved_gbs": b_lj * value_per_gpu / 1e9,
                               "frac": b_lj * value_per_gpu / 1e9 / peak}}

    # ---- e2e: host buffers in and out every step ---------------------------------------------
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
    barrier()
    t0 = time.perf_counter()
    emd.check(L.emd_ctx_tic(ctx))
    for _ in range(Ke):
        e2e_step()
    emd.check(L.emd_ctx_toc(ctx, C.byref(ms)))
    barrier()
    e2e_wall = time.perf_counter() - t0
    t_e = torch.tensor([max(ms.value * 1e-3, e2e_wall)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = n_atoms * Ke / float(t_e.item())
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + d2h_rebuild / rate,
           "steps": Ke, "api": "emd_app_upload -> emd_app_advance(1) -> emd_app_download (pinned host x,v,f)"}

    # device-event phase split of 2*rate further steps (reference timers: force / neigh / comm / other)
    app.advance((-app.get("step")) % rate)
    ph = app.advance_timed(2 * rate)
    phases = {k: 1e3 * v / (2 * rate) for k, v in ph.items()}  # ms per step
    T, PE, KE = app.thermo()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (WORKLOAD if not args.region else f"LJ fcc brick {brick} per GPU (debug override)") +
                       ("" if world == 1 else f"; weak scaling: that brick per GPU, global region {region}"),
                       "atoms_total": n_atoms, "atoms_per_gpu": n_atoms // world, "ghosts_rank0": n_ghost, "neigh_entries_rank0": total_neighs,
                       "l2": "state + list (>400 MB per GPU) exceed the 126 MB L2; no flush between steps",
                       "parallelism": "1 GPU" if world == 1 else
                       f"3-D domain decomposition {grid[0]}x{grid[1]}x{grid[2]} (CommMPI), NCCL halo exchange every step, one process per GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "phase_ms_per_step": phases,
            "thermo_after": {"T": T, "PE": PE, "E": PE + KE}}
    app.close()
    app = None
    if world == 1 and not args.no_snap and not args.region:
        try:
            line["snap"] = snap_section(local_rank)
            if not args.no_cpu_baseline:
                line["snap"]["cpu_baseline"] = snap_cpu_baseline()
        except Exception as e:
            line["snap"] = {"error": str(e)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(20)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
