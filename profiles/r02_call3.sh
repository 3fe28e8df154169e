export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02c_pytest_gpu.log
echo "== e2e sweep"
for v in "1 4" "1 6" "1 8" "2 4" "2 6" "4 3"; do set -- $v
  timeout 300 python bench.py --steps 15 --no-cpu-baseline --no-next-rows --sustained-seconds 0 --e2e-chunk $1 --e2e-slots $2 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $1 slots $2: e2e', round(l['e2e']['value']), 'depth', round(l['e2e_compact']['depth']['value']), 'points', round(l['e2e_compact']['points']['value']))"
done
echo "== e2e 96 sets per call"; timeout 300 python bench.py --steps 6 --no-cpu-baseline --no-next-rows --sustained-seconds 0 --e2e-chunk 1 --e2e-slots 4 --e2e-stacks 96 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E=96 chunk 1 slots 4: e2e', round(l['e2e']['value']), 'depth', round(l['e2e_compact']['depth']['value']))"
echo "== curve pxt4"; timeout 300 python profiles/launch_size_curve.py --pxt 4 --out gpurun_out/r02c_curve_pxt4.txt 2>&1 | tail -9
echo "== curve depth"; timeout 300 python profiles/launch_size_curve.py --format depth --out gpurun_out/r02c_curve_depth.txt 2>&1 | tail -9
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>gpurun_out/r02c_pc.err | tail -1 > gpurun_out/r02c_pc.log; python -c "
import json; l=json.loads(open('gpurun_out/r02c_pc.log').read()); print(l['value'], l['binary_cloud'])"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
$NCU -c 80 --log-file gpurun_out/r02c_pointcloud_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > gpurun_out/r02c_ncu_pc.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02c_pointcloud_launches.csv | tail -12
