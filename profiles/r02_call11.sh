export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest dynamic (warp kernel)"; SLC_DYNA_WARP=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_pinning.py tests/test_cpp_host_api.py -m gpu -x -q -k "dynamic or dyna or cpp" 2>&1 | tail -3
for v in "SLC_DYNA_TILE_W=128" "SLC_DYNA_WARP=1" "SLC_DYNA_TILE_W=128" "SLC_DYNA_WARP=1"; do
  env $v timeout 300 python bench.py --path dynamic 2>gpurun_out/r02k.err | tail -1 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('$v:', round(l['value']), round(l['roofline']['frac'],4), l['checked_against_oracle'])"; tail -2 gpurun_out/r02k.err
done
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
env SLC_DYNA_WARP=1 $NCU -k regex:"dyna_fused|dyna_warp" -c 3 --log-file gpurun_out/r02k_dyna_warp.csv python bench.py --path dynamic --steps 2 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02k_dyna_warp.csv')) if r and r[0].isdigit()]
for r in rows[:8]: print(r[0], r[4][:30], r[8], r[12], r[14])
PY
