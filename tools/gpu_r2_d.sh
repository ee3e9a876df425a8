#!/bin/bash
# round 2, call D (first call of the resumed session): where the build stands.  GPU tests, the driver's default
# line, the emulated shard of 8, timelines, launch lists, two full captures, sanitizer runs.
mkdir -p gpurun_out
TAG=${TAG:-r2d}
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
for c in c4 c3 c2; do
  run ${c} --config $c
done
SKK_NO_PDL=1 run c3_nopdl --config c3
run c4_w8 --config c4 --emulate-world 8
SKK_NO_PDL=1 run c4_w8_nopdl --config c4 --emulate-world 8
SKK_SHARDED_FINISH_SPLIT=1 run c4_w8_split --config c4 --emulate-world 8
SKK_SHARDED_FINISH=0 run c4_w8_sep --config c4 --emulate-world 8
run c3_w2 --config c3 --emulate-world 2
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 2>&1 | tee gpurun_out/${TAG}_trace_c4_w8.txt
timeout 300 python tools/trace_rows.py --config c3 2>&1 | tee gpurun_out/${TAG}_trace_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4_w8.csv python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c4_w8.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c4_w8.csv | tee gpurun_out/${TAG}_launches_c4_w8.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --config c3 --no-cpu-baseline --no-sub --no-check --steps 3 --warmup 1 > gpurun_out/${TAG}_ncu_c3.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_c3.csv | tee gpurun_out/${TAG}_launches_c3.txt
# full captures of the row kernel: the shard of 8 and c3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skk_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_c4_w8_row -f python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full_c4_w8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skk_finish -s 3 -c 1 -o gpurun_out/${TAG}_prof_c4_w8_finish -f python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --no-sub --no-check --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full_c4_w8_finish.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skk_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_c3_row -f python bench.py --config c3 --no-cpu-baseline --no-sub --no-check --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
# sanitizer
( time timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/${TAG}_memcheck.txt
( time SKK_SANITIZE_ROWS=60000 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/${TAG}_racecheck.txt
# the default bench line with all sub-records (what the driver runs), and the reference arm
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err ) 2>&1 | tail -3
python - <<'PY'
import json,os
tag=os.environ.get('TAG','r2d')
d=json.loads(open(f'gpurun_out/{tag}_default.json').read().strip().splitlines()[-1])
print('default c4 step_us=%.1f e2e_us=%.1f frac=%.3f check=%s'%(1e3*d['ms_per_step'],1e3*d['e2e']['ms_per_step'],d['roofline']['frac'],d['check']))
for c,r in d.get('configs',{}).items():
    if 'error' in r: print(c, r); continue
    print(c,'step_us=%.1f e2e_us=%.1f frac=%.3f warm=%s check_ok=%s'%(1e3*r['ms_per_step'],1e3*r['e2e']['ms_per_step'],r['roofline']['frac'],r.get('warm_l2',{}).get('ms_per_step'),r['check'] and r['check']['ok']), r['roofline'].get('alu_peaks'))
print('cpu', d['cpu_baseline'])
PY
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err ) 2>&1 | tail -3; cat gpurun_out/${TAG}_reference.json | cut -c1-600
