mkdir -p gpurun_out/r02s7
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r02s7/launches.csv python tools/snap_time.py 50 50 100 2 > gpurun_out/r02s7/ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02s7/launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][-40:], r[vi])
PY
