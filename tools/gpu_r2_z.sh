#!/bin/bash
# Validation of the round's last build (skk_leapfrog + many-chain NUTS, Gamma-log family): GPU tests, smoke, default bench.
mkdir -p gpurun_out
( time timeout 125 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/z_pytest.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/z_smoke.log
timeout 70 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/z_bench_default.json
