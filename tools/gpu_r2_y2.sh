#!/bin/bash
# round 2, call Y (two GPUs): final build -- sharded parity tests and the c4 line at two ranks
mkdir -p gpurun_out
TAG=${TAG:-r2y2}
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --config c4 --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y2_c4.json').read().strip().splitlines()[-1])
c=d.get('check') or {}
print('c4 n_gpus=%d'%d['n_gpus'], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), d['config'].get('collective'), 'check_ok=%s err_l=%.1e err_g=%.1e'%(c.get('ok'), c.get('err_logp', float('nan')), c.get('err_grad', float('nan'))))
PY
