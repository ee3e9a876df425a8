#!/bin/bash
# A/B session on the GPU box: tests, fused vs separate finish kernels, shard-size sweep, launch lists.
# Results -> gpurun_out/
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
for c in c4 c3 c2; do
  timeout 200 python bench.py --config $c $B > gpurun_out/ab_${c}_fused.json 2>gpurun_out/ab_${c}_fused.err; summ ${c}_fused gpurun_out/ab_${c}_fused.json
  SKK_NO_FUSED_FINISH=1 timeout 200 python bench.py --config $c $B > gpurun_out/ab_${c}_sep.json 2>gpurun_out/ab_${c}_sep.err; summ ${c}_separate gpurun_out/ab_${c}_sep.json
done
for r in 12500000 25000000 50000000; do
  timeout 200 python bench.py --config c4 --rows $r $B > gpurun_out/ab_c4_rows$r.json 2>gpurun_out/ab_c4_rows$r.err; summ c4_rows$r gpurun_out/ab_c4_rows$r.json
done
for c in c3 c4 c2; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ab_launches_$c.csv python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_ncu_$c.log 2>&1
  python - gpurun_out/ab_launches_$c.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
for r in rows[1:][-6:]:
    print('  ', r[ik][:70], r[iv], r[iu])
PY
done
