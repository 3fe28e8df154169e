export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== sweep"; timeout 600 python profiles/sweep_geometry.py > gpurun_out/r02s_sweep_geometry.txt 2>&1; tail -13 gpurun_out/r02s_sweep_geometry.txt
