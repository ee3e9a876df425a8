#!/bin/bash
# round 2, call A: GPU tests, the fused finish on c2/c3/c4, the one-launch exchange on the emulated shard
# of rank 0 of 8, per-warp timelines and launch lists.
mkdir -p gpurun_out
B="--no-cpu-baseline --check"
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'step_us=%.1f'%(1e3*d['ms_per_step']), 'kernel_us=%.1f'%(1e3*d['roofline']['kernel_ms']), 'e2e_us=%.1f'%(1e3*d['e2e']['ms_per_step']), 'launches/step=%.1f'%(d['gpu_launches']/d['steps']), 'frac=%.3f'%d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r2a_pytest.log
for c in c4 c3 c2; do
  timeout 300 python bench.py --config $c $B --steps 30 --warmup 5 > gpurun_out/r2a_$c.json 2> gpurun_out/r2a_$c.err; grep check gpurun_out/r2a_$c.err | head -2; summ $c gpurun_out/r2a_$c.json
done
timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/r2a_c4_w8.json 2> gpurun_out/r2a_c4_w8.err; grep check gpurun_out/r2a_c4_w8.err; summ c4_w8_onelaunch gpurun_out/r2a_c4_w8.json
SKK_SHARDED_FINISH_SPLIT=1 timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/r2a_c4_w8_split.json 2> gpurun_out/r2a_c4_w8_split.err; grep check gpurun_out/r2a_c4_w8_split.err; summ c4_w8_push+final gpurun_out/r2a_c4_w8_split.json
SKK_SHARDED_FINISH=0 timeout 300 python bench.py --config c4 --emulate-world 8 $B --steps 50 --warmup 5 > gpurun_out/r2a_c4_w8_sep.json 2> gpurun_out/r2a_c4_w8_sep.err; grep check gpurun_out/r2a_c4_w8_sep.err; summ c4_w8_separate gpurun_out/r2a_c4_w8_sep.json
timeout 300 python bench.py --config c3 --emulate-world 2 $B --steps 50 --warmup 5 > gpurun_out/r2a_c3_w2.json 2> gpurun_out/r2a_c3_w2.err; grep check gpurun_out/r2a_c3_w2.err; summ c3_w2_onelaunch gpurun_out/r2a_c3_w2.json
timeout 300 python tools/trace_rows.py --config c4 --emulate-world 8 --out gpurun_out/r2a_trace_c4_w8.json 2>&1 | tee gpurun_out/r2a_trace_c4_w8.txt
timeout 300 python tools/trace_rows.py --config c4 2>&1 | tee gpurun_out/r2a_trace_c4.txt
timeout 300 python tools/trace_rows.py --config c3 --out gpurun_out/r2a_trace_c3.json 2>&1 | tee gpurun_out/r2a_trace_c3.txt
timeout 300 python tools/trace_rows.py --config c2 2>&1 | tee gpurun_out/r2a_trace_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_c4_w8.csv python bench.py --config c4 --emulate-world 8 --no-cpu-baseline --steps 3 --warmup 1 > gpurun_out/r2a_ncu_c4_w8.log 2>&1
python tools/launch_list.py gpurun_out/r2a_launches_c4_w8.csv | tee gpurun_out/r2a_launches_c4_w8.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_c3.csv python bench.py --config c3 --no-cpu-baseline --steps 3 --warmup 1 > gpurun_out/r2a_ncu_c3.log 2>&1
python tools/launch_list.py gpurun_out/r2a_launches_c3.csv | tee gpurun_out/r2a_launches_c3.txt
