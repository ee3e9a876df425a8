"""Same-box A/B of the fused integrator kernel (skk_leapfrog) against the torch statement of the same update, and the
cost of a NUTS leaf, on a 100k-row / 100-group regression (written at the end of round 2; the round's GPU budget ended before it ran).

    python tools/sampler_ab.py        # on a B200 box
"""
import sys
import time

import numpy as np
import pandas as pd

import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from helpers import build_model  # noqa: E402
from model_zoo import readme_likelihood  # noqa: E402
from sakkara_b200.sampling import sample_hmc, sample_nuts  # noqa: E402

rng = np.random.default_rng(0)
n, g = 100_000, 100
code = rng.integers(0, g, n)
k = rng.normal(0.7, 0.5, g)
x = rng.standard_normal(n)
df = pd.DataFrame({'x': x, 'y': k[code] * x + 0.3 + 0.5 * rng.standard_normal(n), 'group': code})
model = build_model((df, readme_likelihood(df)))
sample_hmc(model, chains=64, draws=5, tune=5, n_leapfrog=4)          # library load, first plan
for chains in (64, 1024):
    for fused in (True, False, True, False):
        t = time.time()
        sample_hmc(model, chains=chains, draws=40, tune=40, n_leapfrog=10, seed=1, fused_leapfrog=fused)
        dt = time.time() - t
        print(f'hmc chains={chains} fused={fused}: {dt:.3f} s for 800 leapfrog steps incl. plan creation = {dt / 800 * 1e6:.0f} us per step')
t = time.time()
res = sample_nuts(model, chains=256, draws=40, tune=80, seed=1)
dt = time.time() - t
print(f'nuts chains=256: {dt:.3f} s for 120 transitions, mean tree {res.stats["n_steps"].mean():.1f} leaves (draws), '
      f'accept {res.stats["accept"].mean():.3f}, divergences {int(res.stats["diverging"].sum())}')
