#!/bin/bash
# round 2: two real GPUs -- sharded parity tests (peer exchange and NCCL), c4 / c3 row-sharded bench lines with the oracle check
mkdir -p gpurun_out
TAG=${TAG:-r2g2}
N=${N:-2}
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/${TAG}_pytest_multi.log
port=29520
for c in c4 c3 c5; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err
  python - gpurun_out/${TAG}_$c.json $c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    print(sys.argv[2], 'n_gpus=%d'%d['n_gpus'], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), d['config'].get('collective'), 'check_ok=%s err_l=%.1e err_g=%.1e'%(c.get('ok'), c.get('err_logp', float('nan')), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
