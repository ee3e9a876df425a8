#!/bin/bash
# round 2, call S: parity suite; chain slot kernel with the global-term sums shared by the block's warps (c5 at 1,024 and 128 chains)
mkdir -p gpurun_out
TAG=${TAG:-r2s}
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
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
run c5 --config c5
run c5emulateworld8 --config c5 --emulate-world 8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_c5_w8.csv python bench.py --config c5 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c5_w8.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c5_w8.csv | grep skk
