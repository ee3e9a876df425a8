#!/bin/bash
# End-of-round measurement on one B200: smoke, GPU tests, the default bench line, the other configs,
# the CPU arm, the ncu launch list of the default command and one full capture of the row kernel.
mkdir -p gpurun_out
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4)
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
for c in c2 c3 c5; do timeout 200 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2>/dev/null; cat gpurun_out/bench_$c.json; done
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_default.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skk_row_kernel -s 3 -c 1 -o gpurun_out/prof_c4_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skk_finish_kernel -s 3 -c 1 -o gpurun_out/prof_c4_finish_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c4_finish.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
