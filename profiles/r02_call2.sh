# 8-GPU box: host-link probe, then the bench under torchrun exactly as the driver launches it
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" 
nvidia-smi topo -m 2>/dev/null | head -14
echo "== probe"; timeout 600 python profiles/hostlink_probe.py --out gpurun_out/r02_hostlink_probe > gpurun_out/r02_hostlink_probe.stdout 2>&1; tail -75 gpurun_out/r02_hostlink_probe.stdout
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=8"; timeout 900 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err; tail -c 5000 gpurun_out/r02_bench_n8.log; tail -3 gpurun_out/r02_bench_n8.err
echo "== config4 N=8"; timeout 900 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --config config4 --steps 3 --warmup 3 > gpurun_out/r02_config4_n8.log 2> gpurun_out/r02_config4_n8.err; tail -c 3000 gpurun_out/r02_config4_n8.log; tail -3 gpurun_out/r02_config4_n8.err
echo "== bench N=2"; timeout 600 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.log 2> gpurun_out/r02_bench_n2.err; tail -c 2500 gpurun_out/r02_bench_n2.log; tail -3 gpurun_out/r02_bench_n2.err
