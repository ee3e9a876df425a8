#!/bin/bash
# round 2, call U: full parity suite on the final build (incl. the larger member-wise models), smoke(), default bench line
mkdir -p gpurun_out
TAG=${TAG:-r2u}
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -6 | tee gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err ) 2>&1 | tail -3
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_default.json').read().strip().splitlines()[-1])
print('default step_us=%.1f'%(1e3*d['ms_per_step']), 'value=%.3g'%d['value'], 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'frac=%.3f'%d['roofline']['frac'], 'check', (d.get('check') or {}).get('ok'), 'launches', d['gpu_launches'])
for k,v in d['configs'].items():
    print(' ', k, 'step_us=%.1f'%(1e3*v['ms_per_step']), 'frac=%.3f'%v['roofline']['frac'], 'check', (v.get('check') or {}).get('ok'))
r=json.loads(open('gpurun_out/r2u_reference.json').read().strip().splitlines()[-1])
print('reference value=%.3g'%r['value'], r['cpu_baseline']['cores'], 'cores', r['config'].get('rows_per_step'), 'ms_per_step=%.1f'%r['ms_per_step'], 'ratio value %.0f e2e %.0f'%(d['value']/r['value'], d['e2e']['value']/r['value']))
PY
