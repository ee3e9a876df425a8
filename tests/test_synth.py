"""The direct plan builders of sakkara_b200.synth produce what build() lowers from the DSL models."""
import pandas as pd

from golden.plan_io import assert_same, plan_to_arrays
from helpers import build_model
from model_zoo import readme_likelihood
from sakkara_b200 import synth


def _dsl(df, make):
    return plan_to_arrays(build_model((df, make)).plan)


def test_linreg_plan_equals_dsl():
    cols, _ = synth.linreg_columns(500, 13, seed=7)
    df = pd.DataFrame(cols)
    assert_same(_dsl(df, readme_likelihood(df)), plan_to_arrays(synth.linreg_plan(cols)))


def test_logistic_plan_equals_dsl():
    cols, _ = synth.logistic_columns(800, 5, 4, seed=7)
    df = pd.DataFrame(cols)

    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        slope = DC(pm.Normal, name='slope', group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=1.0))
        icpt = DC(pm.Normal, name='icpt', group='g2',
                  mu=DC(pm.Normal, group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=2.0)),
                  sigma=DC(pm.HalfNormal, sigma=1.0))
        return api.Likelihood(pm.Bernoulli, logit_p=slope * data['x'] + icpt, observed=data['y'])
    assert_same(_dsl(df, make), plan_to_arrays(synth.logistic_plan(cols)))


def test_poisson_plan_equals_dsl():
    cols, _ = synth.poisson_columns(900, 7, 5, seed=7)
    df = pd.DataFrame(cols)

    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        beta = DC(pm.Normal, name='beta')
        a = DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, sigma=1.0))
        b = DC(pm.Normal, name='b', group='B', sigma=DC(pm.HalfNormal, sigma=1.0))
        return api.Likelihood(pm.Poisson, mu=api.f_(pt.exp)(beta * data['x'] + a + b), observed=data['y'])
    assert_same(_dsl(df, make), plan_to_arrays(synth.poisson_plan(cols)))
