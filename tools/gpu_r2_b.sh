#!/bin/bash
# round 2, call B: single-thread fences, PDL finish, flattened prologue -- tests, benches, traces
mkdir -p gpurun_out
TAG=${TAG:-r2b}
B="--no-cpu-baseline --no-sub"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'launches/step=%.1f'%(d['gpu_launches']/d['steps']), 'frac=%.3f'%d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
for c in c4 c3 c2; do
  timeout 300 python bench.py --config $c $B --steps 30 --warmup 5 > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err; grep check gpurun_out/${TAG}_$c.err | head -2; summ $c gpurun_out/${TAG}_$c.json
  SKK_NO_PDL=1 timeout 300 python bench.py --config $c $B --steps 30 --warmup 5 > gpurun_out/${TAG}_${c}_nopdl.json 2> gpurun_out/${TAG}_${c}_nopdl.err; summ ${c}_nopdl gpurun_out/${TAG}_${c}_nopdl.json
done
timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/${TAG}_c4_w8.json 2> gpurun_out/${TAG}_c4_w8.err; grep check gpurun_out/${TAG}_c4_w8.err; summ c4_w8_onelaunch gpurun_out/${TAG}_c4_w8.json
SKK_NO_PDL=1 timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/${TAG}_c4_w8_nopdl.json 2> gpurun_out/${TAG}_c4_w8_nopdl.err; summ c4_w8_onelaunch_nopdl gpurun_out/${TAG}_c4_w8_nopdl.json
SKK_SHARDED_FINISH_SPLIT=1 timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/${TAG}_c4_w8_split.json 2> gpurun_out/${TAG}_c4_w8_split.err; grep check gpurun_out/${TAG}_c4_w8_split.err; summ c4_w8_push+final gpurun_out/${TAG}_c4_w8_split.json
SKK_SHARDED_FINISH=0 timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/${TAG}_c4_w8_sep.json 2> gpurun_out/${TAG}_c4_w8_sep.err; grep check gpurun_out/${TAG}_c4_w8_sep.err; summ c4_w8_separate gpurun_out/${TAG}_c4_w8_sep.json
timeout 300 python bench.py --config c3 --emulate-world 2 $B --steps 50 --warmup 5 > gpurun_out/${TAG}_c3_w2.json 2> gpurun_out/${TAG}_c3_w2.err; grep check gpurun_out/${TAG}_c3_w2.err; summ c3_w2_onelaunch gpurun_out/${TAG}_c3_w2.json
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 2>&1 | tee gpurun_out/${TAG}_trace_c4_w8.txt
timeout 300 python tools/trace_rows.py --config c3 2>&1 | tee gpurun_out/${TAG}_trace_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4_w8.csv python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c4_w8.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c4_w8.csv | tee gpurun_out/${TAG}_launches_c4_w8.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --config c3 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c3.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c3.csv | tee gpurun_out/${TAG}_launches_c3.txt
# the default bench line with all sub-records (what the driver runs), and the reference arm
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err ) 2>&1 | tail -3
python - <<'PY'
import json,os
tag=os.environ.get('TAG','r2b')
d=json.loads(open(f'gpurun_out/{tag}_default.json').read().strip().splitlines()[-1])
print('default c4 step_us=%.1f e2e_us=%.1f frac=%.3f check=%s'%(1e3*d['ms_per_step'],1e3*d['e2e']['ms_per_step'],d['roofline']['frac'],d['check']))
for c,r in d.get('configs',{}).items():
    if 'error' in r: print(c, r); continue
    print(c,'step_us=%.1f e2e_us=%.1f frac=%.3f warm=%s check_ok=%s'%(1e3*r['ms_per_step'],1e3*r['e2e']['ms_per_step'],r['roofline']['frac'],r.get('warm_l2',{}).get('ms_per_step'),r['check'] and r['check']['ok']), r['roofline'].get('alu_peaks'))
print('cpu', d['cpu_baseline'])
PY
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err ) 2>&1 | tail -3; cat gpurun_out/${TAG}_reference.json | cut -c1-600
