#!/bin/bash
# round 2, call V: parity suite with the row-dependent likelihood sigma (one plan per sigma)
mkdir -p gpurun_out
TAG=${TAG:-r2v}
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/${TAG}_pytest.log
