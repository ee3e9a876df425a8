"""
The CPU oracle (oracle/logp_numpy.py, oracle/logp_c.c) against independent evidence -- the
floating-point values of this path are unpinned by the reference, so the oracle is
anchored on: torch autograd over the recorded graph with torch.distributions densities,
scipy.stats log-densities, central finite differences, and hand-computed values on the README model.
"""
import numpy as np
import pytest
from scipy import stats

from helpers import build_model, theta_points
from model_zoo import zoo
from oracle.autograd_check import logp_dlogp_autograd
from oracle.logp_c import COracle
from oracle.logp_numpy import logp_dlogp, parity_error

ZOO = zoo()


@pytest.mark.parametrize('name', sorted(ZOO))
def test_numpy_oracle_equals_autograd_on_recorded_graph(name):
    model = build_model(ZOO[name])
    for theta in theta_points(model.plan, 4, seed=1):
        logp, grad = logp_dlogp_autograd(model.recorded, theta)
        err_l, err_g = parity_error(logp, grad, model.plan, theta)
        assert err_l < 1e-13 and err_g < 1e-13


@pytest.mark.parametrize('dtype,tol', [('float64', 1e-13), ('float32', 2e-6)])
@pytest.mark.parametrize('name', sorted(ZOO))
def test_c_oracle_equals_numpy_oracle(name, dtype, tol):
    model = build_model(ZOO[name])
    if model.plan.likelihood.sigma_rows_element is not None or model.plan.likelihood.family == 'StudentT':
        pytest.skip('the C oracle (bench baseline) covers the four bench families with one sigma; row-dependent sigmas and StudentT '
                    'are covered by the NumPy and autograd oracles')
    f = COracle(model.plan, dtype, threads=3)
    for theta in theta_points(model.plan, 3, seed=2):
        theta = theta.astype(dtype)
        logp, grad = f(theta)
        err_l, err_g = parity_error(logp, grad, model.plan, theta.astype(np.float64))
        assert err_l < tol and err_g < tol


def test_readme_model_against_scipy():
    """logp of the README model written out with scipy.stats, term by term."""
    df, _ = ZOO['readme']
    model = build_model(ZOO['readme'])
    theta = np.array([0.3, -0.2, 0.9, -1.1, 0.05, -0.4])     # [mu_k, log s_k, k0, k1, intercept, log s_est]
    mu_k, s_k, k, icpt, s_est = theta[0], np.exp(theta[1]), theta[2:4], theta[4], np.exp(theta[5])
    expected = stats.norm.logpdf(mu_k, 0, 1)
    expected += stats.expon.logpdf(s_k, scale=1.0) + theta[1]                     # Exponential(lam=1) + log-Jacobian
    expected += stats.norm.logpdf(k, mu_k, s_k).sum()
    expected += stats.norm.logpdf(icpt, 0, 1)
    expected += stats.expon.logpdf(s_est, scale=1.0) + theta[5]
    eta = k[df['group'].values] * df['x'].values + icpt
    expected += stats.norm.logpdf(df['y'].values, eta, s_est).sum()
    logp, _ = logp_dlogp(model.plan, theta)
    assert abs(logp - expected) < 1e-12 * abs(expected)


def test_halfnormal_bernoulli_poisson_against_scipy():
    model = build_model(ZOO['nested_logistic'])
    plan = model.plan
    theta = theta_points(plan, 2, seed=9)[1]
    sl = plan.block_slices()
    v = {name: theta[s] for name, s in sl.items()}
    par = plan.blocks[7].params['mu'].index
    expected = stats.norm.logpdf(v['mu_slope'], 0, 1).sum()
    expected += (stats.halfnorm.logpdf(np.exp(v['sigma_slope']), scale=1.0) + v['sigma_slope']).sum()
    expected += stats.norm.logpdf(v['slope'], v['mu_slope'], np.exp(v['sigma_slope'])).sum()
    expected += stats.norm.logpdf(v['mu_mu_icpt'], 0, 1).sum()
    expected += (stats.halfnorm.logpdf(np.exp(v['sigma_mu_icpt']), scale=2.0) + v['sigma_mu_icpt']).sum()
    expected += stats.norm.logpdf(v['mu_icpt'], v['mu_mu_icpt'], np.exp(v['sigma_mu_icpt'])).sum()
    expected += (stats.halfnorm.logpdf(np.exp(v['sigma_icpt']), scale=1.0) + v['sigma_icpt']).sum()
    expected += stats.norm.logpdf(v['icpt'], v['mu_icpt'][par], np.exp(v['sigma_icpt'])).sum()
    eta = v['slope'][plan.terms[0].index] * plan.terms[0].x + v['icpt'][plan.terms[1].index]
    p = 1 / (1 + np.exp(-eta))
    expected += stats.bernoulli.logpmf(plan.likelihood.y.astype(int), p).sum()
    assert abs(logp_dlogp(plan, theta)[0] - expected) < 1e-11 * abs(expected)

    model = build_model(ZOO['crossed_poisson'])
    plan = model.plan
    theta = theta_points(plan, 2, seed=9)[1]
    eta = sum(theta[plan.blocks[t.block].offset + t.index] * (1.0 if t.x is None else t.x) for t in plan.terms)
    lik = stats.poisson.logpmf(plan.likelihood.y.astype(int), np.exp(eta)).sum()
    from oracle.logp_c import priors_only
    assert abs(logp_dlogp(plan, theta)[0] - (logp_dlogp(priors_only(plan), theta)[0] + lik)) < 1e-10 * abs(lik)


@pytest.mark.parametrize('name', sorted(ZOO))
def test_gradient_against_finite_differences(name):
    plan = build_model(ZOO[name]).plan
    theta = theta_points(plan, 2, seed=4, scale=0.3)[1]
    _, grad = logp_dlogp(plan, theta)
    h = 1e-6
    for j in range(plan.n_params):
        e = np.zeros_like(theta)
        e[j] = h
        fd = (logp_dlogp(plan, theta + e)[0] - logp_dlogp(plan, theta - e)[0]) / (2 * h)
        assert abs(fd - grad[j]) <= 1e-5 * max(1.0, abs(grad[j])), (j, fd, grad[j])


def test_float32_mode_accumulates_sums_in_float64():
    plan = build_model(ZOO['shuffled_linreg']).plan
    theta = theta_points(plan, 2, seed=1)[1].astype(np.float32)
    logp, grad = logp_dlogp(plan, theta, np.float32)
    assert logp.dtype == np.float32 and grad.dtype == np.float32
    err_l, err_g = parity_error(logp, grad, plan, theta.astype(np.float64))
    assert err_l < 1e-6 and err_g < 1e-5


def test_error_metric_uses_the_l1_mass_of_the_reduction():
    plan = build_model(ZOO['readme']).plan
    theta = np.zeros(plan.n_params)
    logp, grad, mass_l, mass_g = logp_dlogp(plan, theta, np.float64, True)
    assert (mass_g >= np.abs(grad) - 1e-12).all() and mass_l >= abs(logp) - 1e-12


def test_round2_prior_families_against_scipy():
    """HalfCauchy, Gamma, InverseGamma, Cauchy, Laplace, LogNormal, HalfStudentT: PyMC's parametrisation written out
    with scipy.stats on the back-transformed value, plus the log-Jacobian of the LogTransform for the positive ones."""
    model = build_model(ZOO['wide_priors'])
    plan = model.plan
    theta = theta_points(plan, 2, seed=11)[1]
    v = {name: theta[s] for name, s in plan.block_slices().items()}
    e = np.exp
    expected = stats.cauchy.logpdf(v['k_loc'], 0.1, 2.0).sum()
    expected += (stats.halfcauchy.logpdf(e(v['k_scale']), scale=1.5) + v['k_scale']).sum()
    expected += stats.norm.logpdf(v['k'], v['k_loc'], e(v['k_scale'])).sum()
    expected += (stats.lognorm.logpdf(e(v['c_level']), s=0.8, scale=e(-0.5)) + v['c_level']).sum()
    expected += (stats.gamma.logpdf(e(v['c_scale']), a=2.5, scale=1 / 2.0) + v['c_scale']).sum()
    expected += stats.norm.logpdf(v['c'], e(v['c_level']), e(v['c_scale'])).sum()
    expected += stats.laplace.logpdf(v['icpt_loc'], 0.2, 1.3).sum()
    expected += (stats.invgamma.logpdf(e(v['icpt_scale']), a=3.0, scale=2.0) + v['icpt_scale']).sum()
    expected += stats.norm.logpdf(v['icpt'], v['icpt_loc'], e(v['icpt_scale'])).sum()
    # a half Student-t is the positive half of a Student-t with twice the density
    expected += (stats.t.logpdf(e(v['noise']), df=4.0, scale=1.2) + np.log(2.0) + v['noise']).sum()
    eta = sum(theta[plan.term_base(t) + t.index] * (1.0 if t.x is None else t.x) for t in plan.terms)
    expected += stats.norm.logpdf(plan.likelihood.y, eta, e(v['noise'])).sum()
    assert abs(logp_dlogp(plan, theta)[0] - expected) < 1e-11 * abs(expected)


def test_negative_binomial_against_scipy():
    plan = build_model(ZOO['crossed_negbinomial']).plan
    theta = theta_points(plan, 2, seed=12)[1]
    eta = sum(theta[plan.term_base(t) + t.index] * (1.0 if t.x is None else t.x) for t in plan.terms)
    alpha, mu = plan.likelihood.sigma_const, np.exp(eta)
    lik = stats.nbinom.logpmf(plan.likelihood.y.astype(int), alpha, alpha / (alpha + mu)).sum()
    from oracle.logp_c import priors_only
    assert abs(logp_dlogp(plan, theta)[0] - (logp_dlogp(priors_only(plan), theta)[0] + lik)) < 1e-11 * abs(lik)


@pytest.mark.parametrize('name,alpha,scale', [('crossed_gamma', 3.0, 1.0), ('grouped_gamma_times', 2.0, 1.0),
                                              ('grouped_gamma_shifted', 2.0, 4.0)])
def test_gamma_glm_against_scipy(name, alpha, scale):
    """Gamma(alpha, beta = alpha / mu): scipy's density with shape alpha and scale mu / alpha, mu = scale * exp(eta) with
    eta written out from the frame and the blocks of theta (not from the plan's terms)."""
    model = build_model(ZOO[name])
    plan, df = model.plan, ZOO[name][0]
    assert plan.likelihood.family == 'Gamma' and plan.likelihood.sigma_const == alpha
    theta = theta_points(plan, 2, seed=14)[1]
    sl = plan.block_slices()
    if name == 'crossed_gamma':
        ia = [list(plan.coords['A']).index(v) for v in df['A']]
        ib = [list(plan.coords['B']).index(v) for v in df['B']]
        eta = theta[sl['beta']][0] * df['x'].values + theta[sl['a']][ia] + theta[sl['b']][ib]
    else:
        ig = [list(plan.coords['group']).index(v) for v in df['group']]
        eta = theta[sl['k']][ig] * df['x'].values + theta[sl['icpt']][0]
    mu = scale * np.exp(eta)
    lik = stats.gamma.logpdf(df['y'].values, alpha, scale=mu / alpha).sum()
    from oracle.logp_c import priors_only
    assert abs(logp_dlogp(plan, theta)[0] - (logp_dlogp(priors_only(plan), theta)[0] + lik)) < 1e-11 * abs(lik)


@pytest.mark.parametrize('name', ['studentt_group_sigma', 'studentt_one_sigma', 'studentt_const_sigma', 'group_sigma'])
def test_location_scale_likelihoods_against_scipy(name):
    """StudentT / Normal with sigma per group, one sigma parameter or a constant: scipy.stats on the constrained scale."""
    model = build_model(ZOO[name])
    plan = model.plan
    theta = theta_points(plan, 2, seed=13)[1]
    sl = plan.block_slices()
    eta = sum(theta[plan.term_base(t) + t.index] * (1.0 if t.x is None else t.x) for t in plan.terms)
    df = ZOO[name][0]
    group = df['group'].values
    if name == 'studentt_const_sigma':
        sigma = np.full(len(df), 0.7)
        y, eta = df['y'].values, eta * 0.7            # (the plan's rows are standardised: undo it)
    elif name == 'studentt_one_sigma':
        sigma, y = np.full(len(df), np.exp(theta[sl['noise']][0])), df['y'].values
    elif name == 'studentt_group_sigma':
        members = list(plan.coords['group'])          # element k of the block belongs to the k-th member (first appearance)
        sigma, y = np.exp(theta[sl['noise']])[[members.index(g) for g in group]], df['y'].values
    else:
        sigma = np.where(group == 0, 1e-1, np.where(group == 1, np.exp(theta[sl['R_1']][0]), np.exp(theta[sl['R_2']][0])))
        y = df['y'].values
    lik = (stats.norm.logpdf(y, eta, sigma) if name == 'group_sigma' else stats.t.logpdf(y, 4.0, loc=eta, scale=sigma)).sum()
    prior = logp_dlogp(plan.take_rows([]), theta)[0]
    assert abs(logp_dlogp(plan, theta)[0] - (prior + lik)) < 1e-11 * max(1.0, abs(lik))
