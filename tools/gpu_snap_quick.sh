# SNAP check: parity tests of the SNAP path + per-phase timing at 250 000 atoms + per-kernel durations
tag=${1:-r02snap}
mkdir -p gpurun_out/$tag
timeout 400 python -m pytest tests/test_gpu_snap.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/snap_time.py 50 50 100 10 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:snap_ -c 12 --csv --log-file gpurun_out/$tag/snap_launches.csv python tools/snap_time.py 50 50 100 2 > gpurun_out/$tag/ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/$tag/snap_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][-22:], r[vi])
PY
