"""
Many-chain HMC on top of the batch evaluator (first cut).

Stock ``pm.sample`` runs one process per chain and evaluates one theta per call, so it can never use
the batch dimension of ``skk_eval``.  This driver keeps the positions, momenta and gradients of all
chains on the device (torch tensors) and evaluates logp+dlogp for every chain with ONE call per
leapfrog step (``SKK_EVAL_DEVICE_IO``: no PCIe traffic inside the trajectory).  With >= 32 chains and a
model without a crossed factor the chain-per-lane kernels are used.

It is a plain fixed-length HMC with a diagonal mass matrix and dual-averaging step-size adaptation
during tuning -- enough to exercise the batch path end to end; it is not a NUTS replacement.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np


@dataclass
class HMCResult:
    draws: np.ndarray            # [chains, draws, P] on the unconstrained scale
    logp: np.ndarray             # [chains, draws]
    accept_rate: np.ndarray      # [chains]
    step_size: float
    value_var_names: list

    def constrained(self, plan) -> Dict[str, np.ndarray]:
        """Posterior draws per block on the constrained scale (exp for log-transformed blocks)."""
        out = {}
        for b in plan.blocks:
            v = self.draws[:, :, b.offset:b.offset + b.size]
            out[b.name] = np.exp(v) if b.transform == 'log' else v
        return out


def hmc_chains(evaluate, theta, draws: int, tune: int, n_leapfrog: int, step_size: float, target_accept: float,
               generator=None):
    """
    The sampler proper, on whatever device ``theta`` lives: fixed-length HMC for all chains at once,
    one ``evaluate(q, logp_out, grad_out)`` call per leapfrog step for the whole batch.

    :param evaluate: fills ``logp_out [chains]`` and ``grad_out [chains, P]`` for the positions ``q [chains, P]``
        (on the GPU: ``skk_eval`` with device pointers, enqueued on the current stream).
    :param theta: start positions ``[chains, P]`` (modified in place).
    :returns: ``(kept [chains, draws, P], kept_logp [chains, draws], accepted [chains], final step size)``
    """
    import torch

    chains, P = theta.shape
    dev, tdt = theta.device, theta.dtype
    logp = torch.empty(chains, dtype=tdt, device=dev)
    grad = torch.empty(chains, P, dtype=tdt, device=dev)
    prop = torch.empty_like(theta)
    prop_logp = torch.empty_like(logp)
    prop_grad = torch.empty_like(grad)

    # dual averaging (Hoffman & Gelman 2014), one step size shared by all chains
    log_eps, log_eps_bar, h_bar = float(np.log(step_size)), 0.0, 0.0
    mu, gamma, t0, kappa = float(np.log(10 * step_size)), 0.05, 10.0, 0.75
    inv_mass = torch.ones(P, dtype=tdt, device=dev)
    kept = torch.empty(chains, draws, P, dtype=tdt, device=dev)
    kept_logp = torch.empty(chains, draws, dtype=tdt, device=dev)
    accepted = torch.zeros(chains, dtype=torch.float64, device=dev)
    window = []

    evaluate(theta, logp, grad)
    for it in range(tune + draws):
        eps = float(np.exp(log_eps if it < tune else log_eps_bar))
        mom = torch.randn(chains, P, dtype=tdt, device=dev, generator=generator) / inv_mass.sqrt()
        h0 = -logp + 0.5 * (mom * mom * inv_mass).sum(1)
        prop.copy_(theta)
        prop_grad.copy_(grad)
        p_half = mom + 0.5 * eps * prop_grad
        for step in range(n_leapfrog):
            prop.add_(eps * inv_mass * p_half)
            evaluate(prop, prop_logp, prop_grad)
            p_half = p_half + (eps if step < n_leapfrog - 1 else 0.5 * eps) * prop_grad
        h1 = -prop_logp + 0.5 * (p_half * p_half * inv_mass).sum(1)
        log_alpha = torch.nan_to_num(h0 - h1, nan=-float('inf'))
        take = torch.rand(chains, dtype=tdt, device=dev, generator=generator).log() < log_alpha
        theta[take] = prop[take]
        logp[take] = prop_logp[take]
        grad[take] = prop_grad[take]
        if it < tune:
            alpha = float(log_alpha.clamp(max=0).exp().mean())
            m = it + 1
            h_bar = (1 - 1 / (m + t0)) * h_bar + (target_accept - alpha) / (m + t0)
            log_eps = mu - np.sqrt(m) / gamma * h_bar
            log_eps_bar = m ** -kappa * log_eps + (1 - m ** -kappa) * log_eps_bar
            if it >= tune // 4:
                window.append(theta.clone())
            if it == (3 * tune) // 4 and len(window) > 10:      # one mass-matrix update from the warm-up draws
                var = torch.stack(window).reshape(-1, P).var(0)
                inv_mass = var.clamp(min=1e-8)
                window = []
                log_eps, log_eps_bar, h_bar, mu = float(np.log(eps)), 0.0, 0.0, float(np.log(10 * eps))
        else:
            k = it - tune
            kept[:, k] = theta
            kept_logp[:, k] = logp
            accepted += take.double()
    return kept, kept_logp, accepted, float(np.exp(log_eps_bar))


def sample_hmc(model, chains: int = 64, draws: int = 500, tune: int = 500, n_leapfrog: int = 16,
               step_size: float = 0.01, target_accept: float = 0.8, dtype: str = 'float64', seed: int = 0,
               device: Optional[int] = None, init_scale: float = 0.1) -> HMCResult:
    """
    :param model: a built :class:`sakkara_b200.model.B200Model` (or anything with ``plan`` and
        ``logp_dlogp_function``).
    """
    import torch

    f = model.logp_dlogp_function(dtype=dtype, max_batch=chains, device=device)
    skk = f.skk
    P = f.size
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    stream = torch.cuda.ExternalStream(skk.stream, device=dev)

    def evaluate(q, lp, g):
        skk.eval_device(q.data_ptr(), chains, lp.data_ptr(), g.data_ptr(), sync=False)

    # everything -- the start positions included -- is created and consumed on the plan's stream: the library
    # launches there, and torch's caching allocator then knows which stream uses the buffers
    stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(stream):
        theta = init_scale * torch.randn(chains, P, dtype=tdt, device=dev, generator=gen)
        kept, kept_logp, accepted, eps = hmc_chains(evaluate, theta, draws, tune, n_leapfrog, step_size,
                                                    target_accept, generator=gen)
    torch.cuda.synchronize()
    result = HMCResult(draws=kept.cpu().numpy(), logp=kept_logp.cpu().numpy(),
                       accept_rate=(accepted / max(draws, 1)).cpu().numpy(), step_size=eps,
                       value_var_names=[b.value_name for b in model.plan.blocks])
    f.close()
    return result
