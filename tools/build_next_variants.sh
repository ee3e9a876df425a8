#!/bin/bash
# Build the default-off experiment variants as small side libraries under
# sakkara_b200/lib/variants/ (only the f32 Poisson row kernels / the f64 chain kernels are real in them).
set -e
cd "$(dirname "$0")/.."
SKK_VARIANT_ONLY=f32_poisson python tools/build_variant.py predval -DSKK_PRED_VALUE_LOAD
SKK_VARIANT_ONLY=f32_poisson python tools/build_variant.py r19 -DSKK_ROWS_PER_LANE=19
SKK_VARIANT_ONLY=f32_poisson python tools/build_variant.py r21 -DSKK_ROWS_PER_LANE=21
SKK_VARIANT_ONLY=f32_poisson python tools/build_variant.py r19predval -DSKK_ROWS_PER_LANE=19 -DSKK_PRED_VALUE_LOAD
SKK_VARIANT_ONLY=chain_f64,f64_normal python tools/build_variant.py chainhoist -DSKK_CHAIN_HOIST
ls -la sakkara_b200/lib/variants
