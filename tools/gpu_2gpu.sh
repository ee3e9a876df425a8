#!/bin/bash
# 2-GPU sanity: sharded parity tests and the row-sharded bench with the oracle check
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --rows 20000000 --check > gpurun_out/g2_check.json 2> gpurun_out/g2_check.err; grep check gpurun_out/g2_check.err | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/g2_c4.json 2> gpurun_out/g2_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g2_c4.json').read().strip().splitlines()[-1])
print('c4 2gpu step_us=%.1f'%(1e3*d['ms_per_step']), 'row_kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), d['config']['collective'], 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']))
PY
