# 8-GPU box: the bench under torchrun exactly as the driver launches it (N = 8 and 4), with the on-demand pool
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=8"; timeout 900 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n8.log 2> gpurun_out/r02f_bench_n8.err; tail -c 1500 gpurun_out/r02f_bench_n8.log; tail -3 gpurun_out/r02f_bench_n8.err
echo "== bench N=4"; timeout 900 $TR --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n4.log 2> gpurun_out/r02f_bench_n4.err; tail -c 600 gpurun_out/r02f_bench_n4.log; tail -3 gpurun_out/r02f_bench_n4.err
echo "== reference arm N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29623 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02f_ref_n8.log 2> gpurun_out/r02f_ref_n8.err; tail -c 800 gpurun_out/r02f_ref_n8.log; tail -3 gpurun_out/r02f_ref_n8.err
