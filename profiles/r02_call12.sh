export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest generic-N subset"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused_kernel_matches_oracle or random_geometries" 2>&1 | tail -3
echo "== sweep"; timeout 600 python profiles/sweep_geometry.py > gpurun_out/r02l_sweep_geometry.txt 2>&1; cat gpurun_out/r02l_sweep_geometry.txt
for c in config3 config5; do timeout 300 python bench.py --config $c --steps 12 --warmup 4 --no-cpu-baseline --no-next-rows --sustained-seconds 0 2>/dev/null | tail -1 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('$c', round(l['value']), round(l['roofline']['frac'],3), l['checked_against_oracle'])"; done
