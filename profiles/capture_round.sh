#!/bin/bash
# One 1-GPU call that reproduces the per-round evidence under profiles/:
#   gpurun --timeout 2400 -- 'bash profiles/capture_round.sh r02'
# Plain runs first (bench numbers never come from a run under ncu), then the ncu launch lists
# of the same commands, then one `--set full` capture per kernel.  Everything lands in
# gpurun_out/<tag>_*; copy what is to be judged into profiles/.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1

echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

echo "== bench lines"
: > $out/${tag}_bench_lines.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $out/${tag}_bench_ref.err | tail -1 >> $out/${tag}_bench_lines.jsonl
timeout 900 python bench.py 2> $out/${tag}_bench_main.err | tail -1 >> $out/${tag}_bench_lines.jsonl
timeout 600 python bench.py --config config4 --steps 2 --warmup 3 2> $out/${tag}_bench_config4.err | tail -1 >> $out/${tag}_bench_lines.jsonl
for path in dynamic pointcloud ingest app; do
    timeout 600 python bench.py --path $path 2> $out/${tag}_bench_$path.err | tail -1 >> $out/${tag}_bench_lines.jsonl
done
for cfgname in config1 config3 config5 reference_default; do
    timeout 600 python bench.py --config $cfgname --steps 12 --warmup 4 --no-cpu-baseline 2> /dev/null | tail -1 >> $out/${tag}_bench_lines.jsonl
done
cut -c1-260 $out/${tag}_bench_lines.jsonl
echo "== launch-size curve"
timeout 300 python profiles/launch_size_curve.py --out $out/${tag}_launch_size_curve.txt > /dev/null 2>&1; cat $out/${tag}_launch_size_curve.txt
timeout 300 python profiles/launch_size_curve.py --format depth --out $out/${tag}_launch_size_curve_depth.txt > /dev/null 2>&1; cat $out/${tag}_launch_size_curve_depth.txt

echo "== geometry sweep"
timeout 600 python profiles/sweep_geometry.py > $out/${tag}_sweep_geometry.txt 2>&1; cat $out/${tag}_sweep_geometry.txt

echo "== ncu launch lists"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
$NCU -c 60 --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows --sustained-seconds 0 > $out/${tag}_ncu_main.log 2>&1
$NCU -c 60 --log-file $out/${tag}_dyna_launches.csv python bench.py --path dynamic --steps 2 --warmup 1 > $out/${tag}_ncu_dyn.log 2>&1
$NCU -c 80 --log-file $out/${tag}_pointcloud_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > $out/${tag}_ncu_pc.log 2>&1
$NCU -c 60 --log-file $out/${tag}_ingest_launches.csv python bench.py --path ingest --steps 2 --warmup 1 > $out/${tag}_ncu_ing.log 2>&1

echo "== ncu --set full (one capture per kernel)"
FULL="ncu --set full --clock-control none --import-source on"
$FULL -k regex:reconstruct_vec_kernel -s 3 -c 2 -o $out/${tag}_prof_main -f python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --no-next-rows --sustained-seconds 0 > $out/${tag}_ncu_full_main.log 2>&1
$FULL -k regex:"compact_cols_kernel|compact_rows_kernel" -s 2 -c 2 -o $out/${tag}_prof_compact -f python bench.py --path pointcloud --steps 1 --warmup 1 > $out/${tag}_ncu_full_compact.log 2>&1
$FULL -k regex:pc_text_kernel -s 2 -c 2 -o $out/${tag}_prof_pc -f python bench.py --path pointcloud --steps 1 --warmup 1 > $out/${tag}_ncu_full_pc.log 2>&1
$FULL -k regex:"strip_regression21_kernel|dyna_fused_kernel" -s 2 -c 2 -o $out/${tag}_prof_dyn -f python bench.py --path dynamic --steps 1 --warmup 1 --dyna-frames 30 > $out/${tag}_ncu_full_dyn.log 2>&1
ls -la $out/${tag}_* | awk '{print $5, $9}'
