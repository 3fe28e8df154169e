export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_formats_pool_gpu.py tests/test_pointcloud_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== pointcloud"; timeout 300 python bench.py --path pointcloud 2>gpurun_out/r02o_pc.err | tail -1 > gpurun_out/r02o_pc.json; python -c "
import json; l=json.loads(open('gpurun_out/r02o_pc.json').read()); b=l['binary_cloud']; print(l['value'], l['checked_against_oracle'], b['us_per_frame'], b['roofline_frac'], b['row_major_us_per_frame'], b['row_major_roofline_frac'], b['row_major_checked'], b['checked'])"; tail -2 gpurun_out/r02o_pc.err
