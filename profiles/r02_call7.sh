export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest subset"; timeout 900 python -m pytest tests/test_pointcloud_gpu.py tests/test_reference_pinning.py tests/test_cpp_host_api.py -m gpu -x -q 2>&1 | tail -4
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>gpurun_out/r02g_pc.err | tail -1 > gpurun_out/r02g_pc.log; python -c "
import json; l=json.loads(open('gpurun_out/r02g_pc.log').read()); print(l['value'], l['roofline']['frac'], l['checked_against_oracle'], l['e2e']['value'])"; tail -3 gpurun_out/r02g_pc.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
$NCU -k regex:pc_text -c 6 --log-file gpurun_out/r02g_text_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > gpurun_out/r02g_ncu_pc.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_text_launches.csv')) if r and r[0].isdigit()]
for r in rows[:8]: print(r[0], r[4][-30:], r[8], r[12], r[14])
PY
echo "== app"; timeout 600 python bench.py --path app 2>/dev/null | tail -1 | cut -c1-400
