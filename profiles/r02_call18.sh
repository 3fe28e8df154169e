export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest dynamic"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_pinning.py tests/test_cpp_host_api.py tests/test_formats_pool_gpu.py -m gpu -x -q -k "dynamic or dyna or cpp or reference" 2>&1 | tail -3
for i in 1 2 3; do timeout 300 python bench.py --path dynamic 2>gpurun_out/r02r.err | tail -1 > gpurun_out/r02r_dyn.json; python -c "
import json; l=json.loads(open('gpurun_out/r02r_dyn.json').read()); print(round(l['value']), round(l['roofline']['frac'],4), l['checked_against_oracle'], round(l['e2e']['value']), round(l['e2e_compact']['depth']['value']))"; done; tail -2 gpurun_out/r02r.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
$NCU -c 8 --log-file gpurun_out/r02r_dyna_launches.csv python bench.py --path dynamic --steps 2 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02r_dyna_launches.csv')) if r and r[0].isdigit()]
for r in rows[8:24]: print(r[0], r[4][:40], r[8], r[12], r[14])
PY
