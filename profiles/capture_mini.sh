out=gpurun_out; tag=r01g
: > $out/${tag}_bench_lines.jsonl
for path in pointcloud ingest; do timeout 600 python bench.py --path $path 2>/dev/null | tail -1 >> $out/${tag}_bench_lines.jsonl; done
cut -c1-140 $out/${tag}_bench_lines.jsonl
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
$NCU -c 60 --log-file $out/${tag}_pointcloud_launches.csv python bench.py --path pointcloud --steps 2 --warmup 1 > $out/${tag}_ncu_pc.log 2>&1
$NCU -c 60 --log-file $out/${tag}_ingest_launches.csv python bench.py --path ingest --steps 2 --warmup 1 > $out/${tag}_ncu_ing.log 2>&1
FULL="ncu --set full --clock-control none --import-source on"
$FULL -k regex:bmp_unpack_batch_kernel -s 2 -c 2 -o $out/${tag}_prof_ingest -f python bench.py --path ingest --steps 1 --warmup 1 > $out/${tag}_ncu_full_ing.log 2>&1
$FULL -k regex:pc_emit_kernel -s 4 -c 2 -o $out/${tag}_prof_pc -f python bench.py --path pointcloud --steps 1 --warmup 1 > $out/${tag}_ncu_full_pc.log 2>&1
ls $out/${tag}_*
