out=gpurun_out/${1:-r02lat}; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
python - > $out/init_time.log 2>&1 <<'PY'
import time, os, sys
sys.path.insert(0, os.getcwd())
import examinimd_b200 as emd
for env in ("0", "1"):
    os.environ["EMD_HOST_LATTICE"] = env
    for region in ((80, 80, 80), (160, 160, 160)):
        t0 = time.time()
        app = emd.App(["-il", "input/in.lj", "--comm-type", "SERIAL", "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--region", *map(str, region)])
        t1 = time.time()
        print(f"EMD_HOST_LATTICE={env} region {region}: N={app.get('N')} create (lattice + velocities + first neighbor build + forces) {t1 - t0:.2f} s", flush=True)
        app.close()
PY
cat $out/init_time.log
