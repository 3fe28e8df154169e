export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02d_pytest_gpu.log
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>gpurun_out/r02d_pc.err | tail -1 > gpurun_out/r02d_pc.log; python -c "
import json; l=json.loads(open('gpurun_out/r02d_pc.log').read()); print(l['value'], l['roofline']['frac'], l['e2e']['value'], l['checked_against_oracle'], l['gpu_launches']); print(l['binary_cloud'])"; tail -3 gpurun_out/r02d_pc.err
echo "== curve"; timeout 300 python profiles/launch_size_curve.py --out gpurun_out/r02d_curve.txt 2>&1 | tail -9
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
$NCU -c 80 --log-file gpurun_out/r02d_pointcloud_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > gpurun_out/r02d_ncu_pc.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02d_pointcloud_launches.csv | tail -8
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02d_bench.log 2> gpurun_out/r02d_bench.err; python -c "
import json; l=json.loads(open('gpurun_out/r02d_bench.log').read().strip().splitlines()[-1]); print('value', l['value'], 'roof', l['roofline']['frac'], 'e2e', l['e2e']['value'], l['e2e'].get('roofline'), 'compact', {k:(v['value'] if isinstance(v,dict) else v) for k,v in l['e2e_compact'].items()}); print('sustained', l['sustained']); print(l['next_rows'])"; tail -3 gpurun_out/r02d_bench.err
