"""
Model definitions shared by the tests, written once against the public DSL so that the same
function builds the model with ``sakkara_b200.model`` and -- through ``oracle/ref_harness.py`` --
with the reference's ``sakkara.model``.  Each entry: ``name -> (make_df(), make_likelihood(api, pm, pt))``.
"""
import numpy as np
import pandas as pd


# ---------------------------------------------------------------------------------------------
# C1: README hierarchical linear regression (README.md:18-34; coefficient renamed 'k', quirk D.1)
# ---------------------------------------------------------------------------------------------
def readme_df(n_per_group: int = 5, seed: int = 0) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(2 * n_per_group)
    group = np.repeat([0, 1], n_per_group)
    k = np.array([1.0, -1.0])
    y = x * k[group] + 1e-3 * rng.standard_normal(2 * n_per_group)
    return pd.DataFrame({'x': x, 'y': y, 'group': group})


def readme_likelihood(df):
    def make(api, pm, pt):
        data = api.data_components(df)
        coeff = api.DistributionComponent(pm.Normal, name='k', group='group',
                                          mu=api.DistributionComponent(pm.Normal),
                                          sigma=api.DistributionComponent(pm.Exponential, lam=1))
        intercept = api.DistributionComponent(pm.Normal, name='intercept')
        return api.Likelihood(pm.Normal, name='est', mu=coeff * data['x'] + intercept,
                              sigma=api.DistributionComponent(pm.Exponential, lam=1), observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# linear regression with shuffled, string-labelled groups (first-appearance order matters)
# ---------------------------------------------------------------------------------------------
def shuffled_linreg_df(n: int = 200, groups: int = 7, seed: int = 1) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    labels = np.array([f'g{j}' for j in rng.permutation(groups)])
    code = rng.integers(0, groups, n)
    x = rng.standard_normal(n)
    k = rng.normal(0.7, 0.5, groups)
    y = k[code] * x + 0.3 + 0.5 * rng.standard_normal(n)
    return pd.DataFrame({'x': x, 'y': y, 'group': labels[code]})


# ---------------------------------------------------------------------------------------------
# C3-shaped: nested two-level logistic regression, written like tests/workflows/test_linreg.py:32-57
# ---------------------------------------------------------------------------------------------
def nested_df(n: int = 400, g1: int = 4, per_g1: int = 3, seed: int = 2) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    g2 = rng.integers(0, g1 * per_g1, n)
    x = rng.standard_normal(n)
    slope = rng.normal(0.5, 0.3, g1)
    icpt = rng.normal(0.0, 0.5, g1 * per_g1)
    eta = slope[g2 // per_g1] * x + icpt[g2]
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.int64)
    return pd.DataFrame({'x': x, 'y': y, 'g1': [f'a{v}' for v in g2 // per_g1], 'g2': [f'b{v}' for v in g2]})


def nested_logistic_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        slope = DC(pm.Normal, name='slope', group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=1.0))
        icpt = DC(pm.Normal, name='icpt', group='g2',
                  mu=DC(pm.Normal, group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=2.0)),
                  sigma=DC(pm.HalfNormal, sigma=1.0))
        return api.Likelihood(pm.Bernoulli, logit_p=slope * data['x'] + icpt, observed=data['y'])
    return make


def nested_normal_likelihood(df, sigma=0.7):
    """the test_linreg.py model itself: Normal likelihood with a constant sigma"""
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        coeff = DC(pm.Normal, name='coeff', group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=0.5))
        icpt = DC(pm.Normal, name='intercept', group='g2',
                  mu=DC(pm.Normal, group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=1.5)),
                  sigma=DC(pm.HalfNormal, sigma=0.8))
        return api.Likelihood(pm.Normal, mu=coeff * data['x'] + icpt, sigma=sigma, observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# C4-shaped: Poisson-log GLM with crossed random effects, both operand orders (quirk D.3)
# ---------------------------------------------------------------------------------------------
def crossed_df(n: int = 600, a: int = 6, b: int = 4, seed: int = 3) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    A = rng.integers(0, a, n)
    B = rng.integers(0, b, n)
    x = rng.standard_normal(n)
    eta = 0.2 * x + rng.normal(0, 0.3, a)[A] + rng.normal(0, 0.3, b)[B]
    return pd.DataFrame({'x': x, 'A': A, 'B': B, 'y': rng.poisson(np.exp(eta))})


def crossed_poisson_likelihood(df, obs_term_first: bool = True):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        beta = DC(pm.Normal, name='beta')
        a = DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, sigma=1.0))
        b = DC(pm.Normal, name='b', group='B', sigma=DC(pm.HalfNormal, sigma=1.0))
        eta = beta * data['x'] + a + b if obs_term_first else a + b + beta * data['x']
        return api.Likelihood(pm.Poisson, mu=api.f_(pt.exp)(eta), observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# tuple groups: parameter defined on (time, sensor), from tests/components/test_distribution_component.py:56-93
# ---------------------------------------------------------------------------------------------
def tuple_group_df() -> pd.DataFrame:
    return pd.DataFrame({
        'building': np.repeat(list('ab'), 10),
        'sensor': np.repeat(list('stuv'), 5),
        'time': np.tile(np.arange(5), 4),
        'x': np.linspace(-1, 1, 20),
        'y': np.cos(np.arange(20.0)),
    })


def tuple_group_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        level = DC(pm.Normal, name='level', group=('time', 'sensor'),
                   mu=DC(pm.Normal, name='by_time', group='time', mu=DC(pm.Normal, name='top', sigma=3.0)),
                   sigma=DC(pm.HalfNormal, name='spread', group='building', sigma=1.0))
        return api.Likelihood(pm.Normal, mu=level - 0.5 * data['x'], sigma=DC(pm.Exponential, lam=2.0),
                              observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# a Normal prior whose mu is a positive variable (valid in the reference: any component may be a parameter)
# ---------------------------------------------------------------------------------------------
def positive_mu_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        k = DC(pm.Normal, name='k', group='group', mu=DC(pm.HalfNormal, name='level', sigma=1.0),
               sigma=DC(pm.Exponential, name='spread', lam=2.0))
        return api.Likelihood(pm.Normal, mu=k * data['x'], sigma=0.5, observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# the prior families with constant parameters (HalfCauchy, Gamma, InverseGamma, Cauchy, Laplace, LogNormal,
# HalfStudentT) as hyper-priors of a random-slope / random-intercept regression
# ---------------------------------------------------------------------------------------------
def wide_priors_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        k = DC(pm.Normal, name='k', group='group', mu=DC(pm.Cauchy, name='k_loc', alpha=0.1, beta=2.0),
               sigma=DC(pm.HalfCauchy, name='k_scale', beta=1.5))
        c = DC(pm.Normal, name='c', group='group', mu=DC(pm.LogNormal, name='c_level', mu=-0.5, sigma=0.8),
               sigma=DC(pm.Gamma, name='c_scale', alpha=2.5, beta=2.0))
        icpt = DC(pm.Normal, name='icpt', mu=DC(pm.Laplace, name='icpt_loc', mu=0.2, b=1.3),
                  sigma=DC(pm.InverseGamma, name='icpt_scale', alpha=3.0, beta=2.0))
        return api.Likelihood(pm.Normal, mu=k * data['x'] + c + icpt,
                              sigma=DC(pm.HalfStudentT, name='noise', nu=4.0, sigma=1.2), observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# GroupComponent (reference composable/group.py, tests/components/test_group_component.py): a coefficient and a
# prior scale given member by member -- different parameter blocks, a scalar variable and a constant
# ---------------------------------------------------------------------------------------------
def member_wise_likelihood(df):
    def make(api, pm, pt):
        DC, GC = api.DistributionComponent, api.GroupComponent
        data = api.data_components(df[['x', 'y']])
        slope = GC('sensor', 'slope', {
            's': DC(pm.Normal, name='slope_s', group='time', sigma=2.0),
            't': DC(pm.Normal, name='slope_t', group='time', mu=DC(pm.Normal, name='slope_t_level')),
            'u': DC(pm.Normal, name='slope_u'),
            'v': 0.75})
        spread = GC('sensor', 'spread', {'s': 0.5, 't': DC(pm.HalfNormal, name='spread_t', sigma=1.0),
                                         'u': DC(pm.Exponential, name='spread_u', lam=2.0), 'v': 1.5})
        level = DC(pm.Normal, name='level', group='sensor', mu=0.3, sigma=spread)
        return api.Likelihood(pm.Normal, mu=slope * data['x'] + level, sigma=DC(pm.HalfNormal, name='noise', sigma=1.0),
                              observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# known measurement errors: sigma of the Normal likelihood is a data column (one value per row)
# ---------------------------------------------------------------------------------------------
def row_sigma_df(n: int = 150, groups: int = 5, seed: int = 8) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    g = rng.integers(0, groups, n)
    x = rng.normal(size=n)
    s = rng.uniform(0.2, 1.5, n)
    return pd.DataFrame({'x': x, 's': s, 'group': [f'g{v}' for v in g],
                         'y': x * rng.normal(0.7, 0.4, groups)[g] + 0.3 + s * rng.normal(size=n)})


def row_sigma_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y', 's']])
        k = DC(pm.Normal, name='k', group='group', mu=DC(pm.Normal, name='k_mean'), sigma=DC(pm.HalfNormal, name='k_sd'))
        return api.Likelihood(pm.Normal, mu=k * data['x'] + DC(pm.Normal, name='icpt'), sigma=data['s'],
                              observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# over-dispersed counts: NegativeBinomial(mu = exp(eta), alpha) on the crossed frame
# ---------------------------------------------------------------------------------------------
def crossed_negbinomial_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        a = DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, name='sigma_a', sigma=1.0))
        b = DC(pm.Normal, name='b', group='B', sigma=DC(pm.HalfCauchy, name='sigma_b', beta=1.0))
        eta = DC(pm.Normal, name='beta') * data['x'] + a + b
        return api.Likelihood(pm.NegativeBinomial, mu=api.f_(pt.exp)(eta), alpha=2.5, observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# positive continuous observations: the Gamma GLM with a log link, Gamma(alpha, beta = alpha / exp(eta)), in the three
# spellings the lowering accepts (alpha / exp(eta); alpha * exp(-eta); a constant other than alpha over exp(eta))
# ---------------------------------------------------------------------------------------------
def gamma_df(crossed: bool, seed: int = 11) -> pd.DataFrame:
    df = crossed_df(seed=seed) if crossed else shuffled_linreg_df(seed=seed)
    rng = np.random.default_rng(seed + 100)
    mean = np.exp(0.3 * df['x'].values)
    df['y'] = rng.gamma(3.0, mean / 3.0)
    return df


def crossed_gamma_likelihood(df):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        a = DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, name='sigma_a', sigma=1.0))
        b = DC(pm.Normal, name='b', group='B', sigma=DC(pm.HalfNormal, name='sigma_b', sigma=1.0))
        eta = DC(pm.Normal, name='beta') * data['x'] + a + b
        return api.Likelihood(pm.Gamma, alpha=3.0, beta=3.0 / api.f_(pt.exp)(eta), observed=data['y'])
    return make


def grouped_gamma_likelihood(df, spelling):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        k = DC(pm.Normal, name='k', group='group', sigma=DC(pm.HalfNormal, name='k_sd', sigma=1.0))
        eta = k * data['x'] + DC(pm.Normal, name='icpt')
        if spelling == 'times':
            beta = 2.0 * api.f_(pt.exp)(-eta)               # alpha * exp(-eta)
        else:
            beta = 0.5 / api.f_(pt.exp)(eta)                # mean alpha / beta = 4 exp(eta): a constant of the predictor
        return api.Likelihood(pm.Gamma, alpha=2.0, beta=beta, observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# sigma per group as a parameter: the measurement model of /root/reference/tests/workflows/test_state_space.py:64-82
# (a GroupComponent of a constant and a positive variable as sigma of a Normal likelihood), on a regression
# ---------------------------------------------------------------------------------------------
def group_sigma_df(n: int = 160, seed: int = 10) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    group = rng.integers(0, 3, n)
    x = rng.normal(size=n)
    sd = np.array([1e-1, 0.6, 1.4])[group]
    return pd.DataFrame({'x': x, 'group': group, 'y': 0.8 * x - 0.2 + sd * rng.normal(size=n)})


def group_sigma_likelihood(df):
    def make(api, pm, pt):
        DC, GC = api.DistributionComponent, api.GroupComponent
        data = api.data_components(df[['x', 'y']])
        R = GC(name='R', group='group', membercomponents={0: 1e-1, 1: DC(pm.Exponential, name='R_1', lam=10),
                                                           2: DC(pm.HalfNormal, name='R_2', sigma=2.0)})
        k = DC(pm.Normal, name='k', group='group', sigma=DC(pm.HalfNormal, name='k_sd', sigma=1.0))
        return api.Likelihood(pm.Normal, mu=k * data['x'] + DC(pm.Normal, name='icpt'), sigma=R, observed=data['y'])
    return make


# ---------------------------------------------------------------------------------------------
# robust regression: StudentT likelihood, sigma per group as a parameter / one parameter / a constant
# ---------------------------------------------------------------------------------------------
def studentt_likelihood(df, sigma_kind):
    def make(api, pm, pt):
        DC = api.DistributionComponent
        data = api.data_components(df[['x', 'y']])
        k = DC(pm.Normal, name='k', group='group', sigma=DC(pm.HalfNormal, name='k_sd', sigma=1.0))
        sigma = {'group': lambda: DC(pm.HalfNormal, name='noise', group='group', sigma=1.0),
                 'one': lambda: DC(pm.Exponential, name='noise', lam=2.0),
                 'const': lambda: 0.7}[sigma_kind]()
        return api.Likelihood(pm.StudentT, nu=4.0, mu=k * data['x'] + DC(pm.Normal, name='icpt'), sigma=sigma,
                              observed=data['y'])
    return make


def zoo():
    out = {}
    df = readme_df()
    out['readme'] = (df, readme_likelihood(df))
    df = shuffled_linreg_df()
    out['shuffled_linreg'] = (df, readme_likelihood(df))
    df = nested_df()
    out['nested_logistic'] = (df, nested_logistic_likelihood(df))
    out['nested_normal'] = (df, nested_normal_likelihood(df))
    df = crossed_df()
    out['crossed_poisson'] = (df, crossed_poisson_likelihood(df, True))
    out['crossed_poisson_dense'] = (df, crossed_poisson_likelihood(df, False))
    df = tuple_group_df()
    out['tuple_group'] = (df, tuple_group_likelihood(df))
    df = shuffled_linreg_df(seed=5)
    out['positive_mu'] = (df, positive_mu_likelihood(df))
    df = shuffled_linreg_df(seed=6)
    out['wide_priors'] = (df, wide_priors_likelihood(df))
    df = tuple_group_df()
    out['member_wise'] = (df, member_wise_likelihood(df))
    df = row_sigma_df()
    out['row_sigma'] = (df, row_sigma_likelihood(df))
    df = crossed_df(seed=9)
    out['crossed_negbinomial'] = (df, crossed_negbinomial_likelihood(df))
    df = gamma_df(True)
    out['crossed_gamma'] = (df, crossed_gamma_likelihood(df))
    df = gamma_df(False)
    out['grouped_gamma_times'] = (df, grouped_gamma_likelihood(df, 'times'))
    out['grouped_gamma_shifted'] = (df, grouped_gamma_likelihood(df, 'shifted'))
    df = group_sigma_df()
    out['group_sigma'] = (df, group_sigma_likelihood(df))
    out['studentt_group_sigma'] = (df, studentt_likelihood(df, 'group'))
    out['studentt_one_sigma'] = (df, studentt_likelihood(df, 'one'))
    out['studentt_const_sigma'] = (df, studentt_likelihood(df, 'const'))
    return out
