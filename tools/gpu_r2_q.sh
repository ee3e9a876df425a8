#!/bin/bash
# round 2, call Q: parity suite; finish sums with batched loads; chain batches with the prior kernels on a parallel
# branch (default) against in line (c5 at 128 chains = rank 0 of 8); the default bench line with all sub-records
mkdir -p gpurun_out
TAG=${TAG:-r2q}
B="--no-cpu-baseline --no-sub"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    w=d.get('warm_l2',{}).get('ms_per_step')
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'warm_us=%s'%(('%.1f'%(1e3*w)) if w else '-'), 'launches/step=%.1f'%(d['gpu_launches']/d['steps']), 'frac=%.3f'%d['roofline']['frac'], 'check_ok=%s err_g=%.1e'%(c.get('ok'), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run() { name=$1; shift; timeout 300 python bench.py "$@" $B --steps 30 --warmup 5 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; summ $name gpurun_out/${TAG}_$name.json; }
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
run c4emulateworld8 --config c4 --emulate-world 8
run c3 --config c3
run c2 --config c2
run c5emulateworld8 --config c5 --emulate-world 8
SKK_PRIORS_IN_LINE=1 run c5emulateworld8_inline --config c5 --emulate-world 8
run c5 --config c5
( time timeout 900 python bench.py > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_default.json').read().strip().splitlines()[-1])
print('default', d['config']['workload'][:40], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'frac=%.3f'%d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'check', (d.get('check') or {}).get('ok'))
for k,v in (d.get('sub') or {}).items():
    print(' sub', k, 'step_us=%.1f'%(1e3*v['ms_per_step']), 'frac=%.3f'%v['roofline']['frac'], 'canonical', (v['roofline'].get('canonical') or {}).get('frac'), 'check', (v.get('check') or {}).get('ok'))
PY
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 2>&1 | grep -v "slowest warp" | tail -8 | tee gpurun_out/${TAG}_trace_c4_w8.txt
