#!/bin/bash
# round 2, call C: LL exchange, PDL trigger modes, arena, R=19 variant, large GPU tests
mkdir -p gpurun_out
TAG=${TAG:-r2c}
B="--no-cpu-baseline --no-sub"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    c=d.get('check') or {}
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'launches/step=%.1f'%(d['gpu_launches']/d['steps']), 'frac=%.3f'%d['roofline']['frac'], 'check_ok=%s err_g=%.1e'%(c.get('ok'), c.get('err_grad', float('nan'))))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run() { name=$1; shift; timeout 300 python bench.py "$@" $B --steps 30 --warmup 5 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; summ $name gpurun_out/${TAG}_$name.json; }
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
V=$PWD/sakkara_b200/lib/variants
for c in c4 c3 c2; do
  run ${c}_trig2 --config $c
  SKK_PDL_TRIGGER=1 run ${c}_trig1 --config $c
  SKK_NO_PDL=1 run ${c}_nopdl --config $c
  SKK_B200_LIB=$V/libskk_b200_r19.so run ${c}_r19 --config $c
done
run c4_w8_trig2 --config c4 --emulate-world 8
SKK_PDL_TRIGGER=1 run c4_w8_trig1 --config c4 --emulate-world 8
SKK_NO_PDL=1 run c4_w8_nopdl --config c4 --emulate-world 8
SKK_SHARDED_FINISH_SPLIT=1 run c4_w8_split --config c4 --emulate-world 8
SKK_SHARDED_FINISH=0 run c4_w8_sep --config c4 --emulate-world 8
SKK_B200_LIB=$V/libskk_b200_r19.so run c4_w8_r19 --config c4 --emulate-world 8
run c3_w2 --config c3 --emulate-world 2
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 2>&1 | tee gpurun_out/${TAG}_trace_c4_w8.txt
timeout 300 python tools/trace_rows.py --config c3 2>&1 | tee gpurun_out/${TAG}_trace_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4_w8.csv python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c4_w8.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c4_w8.csv | tee gpurun_out/${TAG}_launches_c4_w8.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --config c3 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c3.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c3.csv | tee gpurun_out/${TAG}_launches_c3.txt
