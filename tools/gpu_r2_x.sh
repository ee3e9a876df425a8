#!/bin/bash
# round 2, call X: reordered prologue of the row kernel (value-table loads first, tables cleared under their latency)
mkdir -p gpurun_out
TAG=${TAG:-r2x}
B="--no-cpu-baseline --no-sub"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    w=d.get('warm_l2',{}).get('ms_per_step')
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'warm_us=%s'%(('%.1f'%(1e3*w)) if w else '-'), 'frac=%.3f'%d['roofline']['frac'], 'check_ok=%s err_g=%.1e'%(c.get('ok'), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run() { name=$1; shift; timeout 300 python bench.py "$@" $B --steps 30 --warmup 5 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; summ $name gpurun_out/${TAG}_$name.json; }
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
run c4emulateworld8 --config c4 --emulate-world 8
run c4 --config c4
run c3 --config c3
run c2 --config c2
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 2>&1 | grep -v "slowest warp" | head -12
timeout 300 python tools/trace_rows.py --config c3 2>&1 | grep -v "slowest warp" | head -9
