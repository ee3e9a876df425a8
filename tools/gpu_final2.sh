#!/bin/bash
# Last check of the final build: GPU tests, the default bench line, its ncu launch list.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-1500
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
tail -2 gpurun_out/launches_default.csv | cut -c1-300
