export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== sweep"; timeout 600 python profiles/sweep_geometry.py > gpurun_out/r02m_sweep_geometry.txt 2>&1; tail -7 gpurun_out/r02m_sweep_geometry.txt
for c in config3 config5; do timeout 300 python bench.py --config $c --steps 12 --warmup 4 --no-cpu-baseline --no-next-rows --sustained-seconds 0 2>gpurun_out/r02m_$c.err | tail -1 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('$c', round(l['value']), round(l['roofline']['frac'],3), l['checked_against_oracle'], round(l['e2e']['value'],1), {k:(round(v['value'],1) if isinstance(v,dict) else v) for k,v in l['e2e_compact'].items()})"; tail -2 gpurun_out/r02m_$c.err; done
