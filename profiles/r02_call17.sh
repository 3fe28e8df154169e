export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
for sk in 0 1 2 4 6 7 3 5; do
SLC_DYNA_SKIP=$sk $NCU -k regex:"dyna_fused" -c 2 --log-file gpurun_out/r02q_skip.csv python bench.py --path dynamic --steps 2 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02q_skip.csv')) if r and r[0].isdigit()]
d={}
for r in rows:
    if r[0]=='1': d[r[12]]=r[14]
print("skip=$sk", d)
PY
done
