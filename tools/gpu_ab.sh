#!/bin/bash
# A/B session on the GPU box.  Results -> gpurun_out/
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="--steps 100 --warmup 10 --no-cpu-baseline"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'row_kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'frac=%.3f'%d['roofline']['frac'], 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 300 python bench.py --config c4 --rows 20000000 --check --steps 5 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "check"
for r in 100000000 12500000; do
  timeout 200 python bench.py --config c4 --rows $r $B > gpurun_out/ab_c4_new_$r.json 2>gpurun_out/ab_c4_new_$r.err; summ c4_new_$r gpurun_out/ab_c4_new_$r.json
  SKK_B200_LIB=$PWD/sakkara_b200/lib/variants/libskk_b200_head.so timeout 200 python bench.py --config c4 --rows $r $B > gpurun_out/ab_c4_head_$r.json 2>gpurun_out/ab_c4_head_$r.err; summ c4_head_$r gpurun_out/ab_c4_head_$r.json
done
timeout 200 python bench.py --config c4 $B > gpurun_out/ab_c4_new2.json 2>/dev/null; summ c4_new_again gpurun_out/ab_c4_new2.json
