export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02h_pytest_gpu.log
echo "== dynamic"; timeout 300 python bench.py --path dynamic 2>gpurun_out/r02h_dyn.err | tail -1 > gpurun_out/r02h_dyn.log; python -c "
import json; l=json.loads(open('gpurun_out/r02h_dyn.log').read()); print(l['value'], l['roofline']['frac'], l['checked_against_oracle'], l['e2e']['value'], l['e2e_compact'])"; tail -3 gpurun_out/r02h_dyn.err
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>/dev/null | tail -1 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['checked_against_oracle'], l['binary_cloud']['us_per_frame'], l['binary_cloud']['roofline_frac'])"
