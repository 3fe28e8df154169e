export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest dynamic"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_pinning.py tests/test_cpp_host_api.py tests/test_formats_pool_gpu.py -m gpu -x -q -k "dynamic or dyna or reference or cpp" 2>&1 | tail -4
echo "== dynamic"; for i in 1 2; do timeout 300 python bench.py --path dynamic 2>gpurun_out/r02i_dyn.err | tail -1 > gpurun_out/r02i_dyn.log; python -c "
import json; l=json.loads(open('gpurun_out/r02i_dyn.log').read()); print(l['value'], l['roofline']['frac'], l['checked_against_oracle'], l['e2e']['value'], l['e2e_compact']['depth']['value'])"; done; tail -3 gpurun_out/r02i_dyn.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
$NCU -c 12 --log-file gpurun_out/r02i_dyna_launches.csv python bench.py --path dynamic --steps 2 --warmup 1 > gpurun_out/r02i_ncu_dyn.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02i_dyna_launches.csv')) if r and r[0].isdigit()]
for r in rows[8:24]: print(r[0], r[4][:60], r[8], r[12], r[14])
PY
