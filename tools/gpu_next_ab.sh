#!/bin/bash
# Same-box A/B of the variants built by tools/build_next_variants.sh (run under gpurun).
mkdir -p gpurun_out
B="--steps 100 --warmup 10 --no-cpu-baseline"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'frac=%.3f'%d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
V=$PWD/sakkara_b200/lib/variants
timeout 200 python bench.py --config c4 $B > gpurun_out/next_c4_main.json 2>/dev/null; summ c4_main gpurun_out/next_c4_main.json
for v in predval r19 r21 r19predval; do
  SKK_B200_LIB=$V/libskk_b200_$v.so timeout 200 python bench.py --config c4 --rows 20000000 --check --steps 3 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep check
  SKK_B200_LIB=$V/libskk_b200_$v.so timeout 200 python bench.py --config c4 $B > gpurun_out/next_c4_$v.json 2>/dev/null; summ c4_$v gpurun_out/next_c4_$v.json
done
timeout 200 python bench.py --config c5 $B > gpurun_out/next_c5_main.json 2>/dev/null; summ c5_main gpurun_out/next_c5_main.json
SKK_B200_LIB=$V/libskk_b200_chainhoist.so timeout 200 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "chain_per_lane and float64" 2>&1 | tail -1
SKK_B200_LIB=$V/libskk_b200_chainhoist.so timeout 200 python bench.py --config c5 $B > gpurun_out/next_c5_hoist.json 2>/dev/null; summ c5_chainhoist gpurun_out/next_c5_hoist.json
# the sharded pipeline on one GPU (self exchange): separate kernels vs the finish-push kernel (SKK_SHARDED_FINISH=1)
for flag in 0 1; do
  SKK_FORCE_PEER=1 SKK_SHARDED_FINISH=$flag timeout 200 python bench.py --config c4 --rows 20000000 --check --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/next_peer_$flag.err >gpurun_out/next_peer_$flag.json
  grep check gpurun_out/next_peer_$flag.err; summ c4_20M_peer_self_finish$flag gpurun_out/next_peer_$flag.json
done
for flag in 0 1; do
  SKK_FORCE_PEER=1 SKK_SHARDED_FINISH=$flag timeout 200 python bench.py --config c2 --rows 1000000 --check --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/next_peer_c2_$flag.err >gpurun_out/next_peer_c2_$flag.json
  grep check gpurun_out/next_peer_c2_$flag.err | head -2; summ c2_peer_self_finish$flag gpurun_out/next_peer_c2_$flag.json
done
