export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest subset"; timeout 900 python -m pytest tests/test_formats_pool_gpu.py tests/test_pointcloud_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>gpurun_out/r02e_pc.err | tail -1 > gpurun_out/r02e_pc.log; python -c "
import json; l=json.loads(open('gpurun_out/r02e_pc.log').read()); print(l['value'], l['checked_against_oracle']); print(l['binary_cloud'])"; tail -3 gpurun_out/r02e_pc.err
echo "== sweep"; timeout 600 python profiles/sweep_geometry.py > gpurun_out/r02e_sweep_geometry.txt 2>&1; cat gpurun_out/r02e_sweep_geometry.txt
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
$NCU -k regex:compact -c 12 --log-file gpurun_out/r02e_compact_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > gpurun_out/r02e_ncu_pc.log 2>&1
grep -c compact gpurun_out/r02e_compact_launches.csv; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02e_compact_launches.csv')) if r and r[0].isdigit()]
for r in rows[:16]: print(r[0], r[4][-40:], r[8], r[12], r[14])
PY
