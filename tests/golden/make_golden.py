"""
Generates the golden fixtures of tests/golden/*.npz by running the REFERENCE's own
``sakkara.model.build()`` (from /root/reference, through the recording fake PyMC of
``oracle/ref_harness.py``) on the models of tests/model_zoo.py and lowering the recorded graph.
The fixtures pin: theta layout (free-variable creation order, names, shapes, dims, transforms),
every gather index the reference computes (int64, bit-exact) and the data columns.

    python tests/golden/make_golden.py         # only works where /root/reference exists
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'tests'), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from model_zoo import zoo  # noqa: E402
from oracle.ref_harness import reference_plan  # noqa: E402
from plan_io import plan_to_arrays  # noqa: E402


# random models of tests/test_composition_cpu.py (its generator runs on either implementation's modules)
RANDOM_SEEDS = (0, 1, 3, 5, 6, 10, 14, 16, 19, 22, 28, 35)


def random_models():
    import warnings
    from test_composition_cpu import random_model
    for seed in RANDOM_SEEDS:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            frame, _, description = random_model(seed)
            plan = reference_plan(lambda api, pm, pt: random_model(seed, api, pm, pt)[1], frame)
        path = os.path.join(HERE, f'random_{seed:02d}.npz')
        np.savez_compressed(path, **plan_to_arrays(plan))
        print(f'random {seed} ({description}): P={plan.n_params} N={plan.n_rows} -> {os.path.relpath(path, ROOT)}')


def main():
    random_models()
    for name, (df, make) in zoo().items():
        plan = reference_plan(make, df)
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **plan_to_arrays(plan))
        print(f'{name}: P={plan.n_params} N={plan.n_rows} -> {os.path.relpath(path, ROOT)}')


if __name__ == '__main__':
    main()
