#!/bin/bash
# Lean A/B: sharded pipeline on one GPU (self exchange) against the oracle; predicated table load variant.
mkdir -p gpurun_out
B="--steps 100 --warmup 10 --no-cpu-baseline"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'row_kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'frac=%.3f'%d['roofline']['frac'], 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'launches', d['gpu_launches'], d['config']['collective'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
SKK_FORCE_PEER=1 timeout 200 python bench.py --config c4 --rows 20000000 --check --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_peer_self.err >gpurun_out/ab_peer_self.json; grep check gpurun_out/ab_peer_self.err; summ c4_20M_peer_self gpurun_out/ab_peer_self.json
V=$PWD/sakkara_b200/lib/variants/libskk_b200_pred.so
SKK_B200_LIB=$V timeout 200 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "crossed and float32" 2>&1 | tail -1
SKK_B200_LIB=$V timeout 200 python bench.py --config c4 --rows 20000000 --check --steps 5 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep check
timeout 200 python bench.py --config c4 $B > gpurun_out/ab_c4_main.json 2>/dev/null; summ c4_main gpurun_out/ab_c4_main.json
SKK_B200_LIB=$V timeout 200 python bench.py --config c4 $B > gpurun_out/ab_c4_pred.json 2>/dev/null; summ c4_pred gpurun_out/ab_c4_pred.json
