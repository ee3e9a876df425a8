#!/bin/bash
# round 2, call L: ncu of the row kernel with tile-transposed staging (c4 at 100M rows, c3), launch list of the default bench
mkdir -p gpurun_out
TAG=${TAG:-r2l}
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:skk_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_c4_row -f python bench.py --config c4 --no-cpu-baseline --no-sub --no-check --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skk_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_c3_row -f python bench.py --config c3 --no-cpu-baseline --no-sub --no-check --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/${TAG}_ncu_default.log 2>&1
ls -la gpurun_out/*.ncu-rep
