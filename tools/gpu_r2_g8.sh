#!/bin/bash
# round 2: eight real GPUs -- c4 row-sharded at 8 and 4 ranks, c5 chain-sharded at 8, all with the oracle check; sharded parity tests at 4 ranks
mkdir -p gpurun_out
TAG=${TAG:-r2g8}
port=29540
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    print(sys.argv[2], 'n_gpus=%d'%d['n_gpus'], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), d['config'].get('collective'), 'check_ok=%s err_l=%.1e err_g=%.1e'%(c.get('ok'), c.get('err_logp', float('nan')), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run() { n=$1; c=$2; name=$3; shift 3; port=$((port+1))
  timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $c $COLL --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  line gpurun_out/${TAG}_$name.json $name; }
run 8 c4 c4_n8 A=1
run 8 c4 c4_n8_nopdl SKK_NO_PDL=1
COLL="--collective nccl" run 8 c4 c4_n8_nccl A=1
run 4 c4 c4_n4 A=1
run 8 c5 c5_n8 A=1
run 8 c3 c3_n8 A=1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/${TAG}_pytest_multi.log
