#!/bin/bash
# round 2, call M: timelines with the prologue / epilogue broken down (c4 shard of 8, c3, c2), parity suite
mkdir -p gpurun_out
TAG=${TAG:-r2m}
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
for c in "c4 --emulate-world 8" c3 c2 c4; do
  n=$(echo $c | tr -d ' -' )
  timeout 300 python tools/trace_rows.py --config $c 2>&1 | tee gpurun_out/${TAG}_trace_$n.txt
done
