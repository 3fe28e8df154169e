export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02t_bench.log 2> gpurun_out/r02t_bench.err; python -c "
import json; l=json.loads(open('gpurun_out/r02t_bench.log').read().strip().splitlines()[-1]); print('value', round(l['value']), 'roof', round(l['roofline']['frac'],3), 'e2e', round(l['e2e']['value']), round(l['e2e']['roofline']['frac'],3), 'compact', {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else v) for k,v in l['e2e_compact'].items()}, 'sustained', round(l['sustained']['frac'],3), 'checked', l['checked_against_oracle'])"; tail -2 gpurun_out/r02t_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
