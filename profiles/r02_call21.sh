# 8-GPU box, final build: configs[3] with the one-process pool, then the default bench at N = 8
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== config4 N=8"; timeout 900 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --config config4 --steps 3 --warmup 3 > gpurun_out/r02u_config4_n8.log 2> gpurun_out/r02u_config4_n8.err; tail -c 1800 gpurun_out/r02u_config4_n8.log; tail -3 gpurun_out/r02u_config4_n8.err
echo "== bench N=8"; timeout 900 $TR --nproc-per-node 8 --master-port 29632 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02u_bench_n8.log 2> gpurun_out/r02u_bench_n8.err; tail -c 600 gpurun_out/r02u_bench_n8.log; tail -3 gpurun_out/r02u_bench_n8.err
