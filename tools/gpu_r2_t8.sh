#!/bin/bash
# round 2, final multi-GPU call on an 8-GPU box: c4 row-sharded at 8 / 4 / 2 ranks, c5 chain-sharded at 8, c3 at 8, all with
# the oracle check; sharded parity tests (peer exchange and NCCL)
mkdir -p gpurun_out
TAG=${TAG:-r2t8}
port=29640
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    w=d.get('warm_l2',{}).get('ms_per_step')
    print(sys.argv[2], 'n_gpus=%d'%d['n_gpus'], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'warm_us=%s'%(('%.1f'%(1e3*w)) if w else '-'), d['config'].get('collective'), 'check_ok=%s err_l=%.1e err_g=%.1e'%(c.get('ok'), c.get('err_logp', float('nan')), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run() { n=$1; c=$2; name=$3; port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  line gpurun_out/${TAG}_$name.json $name; }
timeout 300 python -m pytest tests/test_gpu_shapes.py -m gpu -q -k chain 2>&1 | tail -2
run 8 c4 c4_n8
run 8 c4 c4_n8_again
run 4 c4 c4_n4
run 2 c4 c4_n2
timeout 300 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_c4_n1.json 2> gpurun_out/${TAG}_c4_n1.err; line gpurun_out/${TAG}_c4_n1.json c4_n1
run 8 c5 c5_n8
timeout 300 python bench.py --config c5 --steps 30 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_c5_n1.json 2> gpurun_out/${TAG}_c5_n1.err; line gpurun_out/${TAG}_c5_n1.json c5_n1
run 8 c3 c3_n8
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/${TAG}_pytest_multi.log
