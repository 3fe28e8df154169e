export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in config3 config3 config1 config5; do timeout 300 python bench.py --config $c --steps 12 --warmup 4 --no-cpu-baseline --no-next-rows --sustained-seconds 0 2>gpurun_out/r02n_$c.err | tail -1 > gpurun_out/r02n_$c.json; python -c "
import sys,json; l=json.loads(open('gpurun_out/r02n_$c.json').read()); print('$c', round(l['value']), round(l['roofline']['frac'],3), l['kernel'], l['checked_against_oracle'], round(l['e2e']['value'],1), {k:(round(v['value'],1) if isinstance(v,dict) and 'value' in v else v) for k,v in l['e2e_compact'].items()})"; tail -2 gpurun_out/r02n_$c.err; done
echo "== sweep"; timeout 600 python profiles/sweep_geometry.py > gpurun_out/r02n_sweep_geometry.txt 2>&1; tail -7 gpurun_out/r02n_sweep_geometry.txt
