export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2
echo "== pytest new"; timeout 900 python -m pytest tests/test_formats_pool_gpu.py -m gpu -x -q 2>&1 | tail -15
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02a_pytest_gpu.log
echo "== curve"; timeout 300 python profiles/launch_size_curve.py --out gpurun_out/r02a_launch_size_curve.txt 2>&1 | tail -12
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02a_bench.log 2> gpurun_out/r02a_bench.err; tail -c 6000 gpurun_out/r02a_bench.log; tail -5 gpurun_out/r02a_bench.err
echo "== config4"; timeout 600 python bench.py --config config4 --steps 2 --warmup 3 > gpurun_out/r02a_config4.log 2> gpurun_out/r02a_config4.err; tail -c 3000 gpurun_out/r02a_config4.log; tail -5 gpurun_out/r02a_config4.err
echo "== probe (1 gpu)"; timeout 300 python profiles/hostlink_probe.py --out gpurun_out/r02a_hostlink_probe_1gpu 2>&1 | tail -40
