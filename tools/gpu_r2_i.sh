#!/bin/bash
# round 2, call I (two GPUs): parity (member-wise components, new prior families), exchange with all ranks' lines polled at once
mkdir -p gpurun_out
TAG=${TAG:-r2i}
N=${N:-2}
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
port=29560
for c in c4 c3; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err
  python - gpurun_out/${TAG}_$c.json $c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    w=d.get('warm_l2',{}).get('ms_per_step')
    print(sys.argv[2], 'n_gpus=%d'%d['n_gpus'], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'warm_us=%s'%(('%.1f'%(1e3*w)) if w else '-'), d['config'].get('collective'), 'check_ok=%s err_l=%.1e err_g=%.1e'%(c.get('ok'), c.get('err_logp', float('nan')), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
