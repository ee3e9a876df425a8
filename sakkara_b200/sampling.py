"""
Many-chain samplers on top of the batch evaluator.

Stock ``pm.sample`` runs one process per chain and evaluates one theta per call, so it can never use
the batch dimension of ``skk_eval``.  The drivers here keep the positions, momenta and gradients of all
chains on the device (torch tensors) and evaluate logp+dlogp for every chain with ONE call per
leapfrog step (``SKK_EVAL_DEVICE_IO``: no PCIe traffic inside a trajectory).  With >= 32 chains and a
model without a crossed factor the chain-per-lane kernels are used.

* :func:`nuts_chains` / :func:`sample_nuts` -- the No-U-Turn sampler (multinomial variant, the algorithm behind
  ``pm.sample``'s default step method) for all chains in lock-step: every chain doubles its own tree in its own
  direction and stops on its own U-turn or divergence, but the leapfrog steps of one tree level are taken
  together, one batched evaluation each; chains that are done are masked.  Step size by dual averaging on the mean
  acceptance statistic of all chains, diagonal mass matrix from Stan's expanding warm-up windows with the draws of
  all chains pooled.
* :func:`hmc_chains` / :func:`sample_hmc` -- fixed-length HMC with the same adaptation, the cheapest driver of the
  batch path.

The arithmetic between two evaluations (half kick, drift) is one fused launch of the library
(``skk_leapfrog``, csrc/skk_aux_kernels.cuh) on the plan's stream when the chains live on the GPU; the same update
in torch operations serves CPU tensors, which is how the loops are tested without a GPU (tests/test_sampling_cpu.py).
"""
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np


@dataclass
class HMCResult:
    draws: np.ndarray            # [chains, draws, P] on the unconstrained scale
    logp: np.ndarray             # [chains, draws]
    accept_rate: np.ndarray      # [chains]
    step_size: float
    value_var_names: list
    stats: Dict[str, np.ndarray] = field(default_factory=dict)   # NUTS: depth, n_steps, diverging, accept, energy

    def constrained(self, plan) -> Dict[str, np.ndarray]:
        """Posterior draws per block on the constrained scale (exp for log-transformed blocks)."""
        out = {}
        for b in plan.blocks:
            v = self.draws[:, :, b.offset:b.offset + b.size]
            out[b.name] = np.exp(v) if b.transform == 'log' else v
        return out

    def summary(self, plan) -> Dict[str, Dict[str, np.ndarray]]:
        """Per block: posterior mean and sd on the constrained scale, bulk ESS and split R-hat per element."""
        out = {}
        for name, v in self.constrained(plan).items():
            flat = v.reshape(v.shape[0], v.shape[1], -1)
            out[name] = {'mean': flat.mean((0, 1)), 'sd': flat.std((0, 1), ddof=1),
                         'ess_bulk': np.array([ess_bulk(flat[:, :, k]) for k in range(flat.shape[2])]),
                         'r_hat': np.array([split_rhat(flat[:, :, k]) for k in range(flat.shape[2])])}
        return out


# ---------------------------------------------------------------------------------------------------------------
# the integrator step shared by both samplers


def kick_drift_torch(q, r, g, eps, inv_mass, a: float, b: float) -> None:
    """``r += a * eps * g``, then ``q += b * eps * inv_mass * r`` (in place; ``eps [chains, 1]``, signed per chain)."""
    r.add_(eps * g, alpha=a)
    if b != 0.0:
        q.add_(eps * inv_mass * r, alpha=b)


class ChainPool:
    """
    What the chains of one process share with the chains of the other processes of a chain-sharded run (one process
    per GPU, every GPU holds all rows and its share of the chains): nothing on the evaluation path,
    only the adaptation statistics -- the mean acceptance behind the common step size and the moment sums behind
    the common mass matrix -- so that all ranks adapt as one ensemble.  The default is a single process.
    """

    def mean(self, value):
        """Mean over all chains of all ranks of a per-rank mean (0-d tensor) -> float."""
        return float(value)

    def sum_(self, tensor, count: int) -> int:
        """In-place sum of ``tensor`` over the ranks; returns the summed ``count``."""
        return count


class DistributedChainPool(ChainPool):
    """``torch.distributed`` all-reduces (NCCL on the GPUs, gloo in the CPU tests); equal chain counts per rank."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)

    def mean(self, value):
        v = value.detach().clone().reshape(1)
        self.dist.all_reduce(v, group=self.group)
        return float(v) / self.world

    def sum_(self, tensor, count: int) -> int:
        self.dist.all_reduce(tensor, group=self.group)
        return count * self.world


class _DualAveraging:
    """Step-size adaptation of Hoffman & Gelman (2014), one step size for all chains."""

    def __init__(self, eps: float, target: float, gamma: float = 0.05, t0: float = 10.0, kappa: float = 0.75):
        self.target, self.gamma, self.t0, self.kappa = target, gamma, t0, kappa
        self.restart(eps)

    def restart(self, eps: float) -> None:
        self.eps0 = float(eps)
        self.mu = float(np.log(10.0 * eps))
        self.log_eps, self.log_eps_bar, self.h_bar, self.m = float(np.log(eps)), 0.0, 0.0, 0

    def update(self, accept: float) -> None:
        self.m += 1
        m = self.m
        self.h_bar = (1 - 1 / (m + self.t0)) * self.h_bar + (self.target - accept) / (m + self.t0)
        self.log_eps = self.mu - np.sqrt(m) / self.gamma * self.h_bar
        w = m ** -self.kappa
        self.log_eps_bar = w * self.log_eps + (1 - w) * self.log_eps_bar

    @property
    def eps(self) -> float:
        return float(np.exp(self.log_eps)) if self.m else self.eps0

    @property
    def eps_final(self) -> float:
        return float(np.exp(self.log_eps_bar)) if self.m else self.eps0


def warmup_windows(tune: int):
    """
    Stan's schedule for the mass matrix: a fast initial buffer, windows ``[start, end)`` of doubling length, a fast
    terminal buffer (75 / 25, 50, ... / 50 iterations for ``tune >= 150``; 15 % / 75 % / 10 % below that).
    """
    if tune < 20:
        return []
    init, term, base = 75, 50, 25
    if init + base + term > tune:
        init, term = int(0.15 * tune), int(0.1 * tune)
        base = tune - init - term
    last = tune - term
    windows, start, size = [], init, base
    while start < last:
        end = start + size
        if end + 2 * size > last:       # the next window would not fit: this one takes the rest
            end = last
        windows.append((start, end))
        start, size = end, 2 * size
    return windows


class _MassAdapt:
    """Pooled (all chains) running variance of theta inside a warm-up window, regularised like Stan's."""

    def __init__(self, P, dev, pool: Optional['ChainPool'] = None):
        import torch
        self.pool = pool or ChainPool()
        self.sums = torch.zeros(2, P, dtype=torch.float64, device=dev)      # sum of theta, sum of theta^2
        self.s1, self.s2 = self.sums[0], self.sums[1]
        self.n = 0

    def add(self, theta) -> None:
        t = theta.double()
        self.s1 += t.sum(0)
        self.s2 += (t * t).sum(0)
        self.n += theta.shape[0]

    def variance(self, dtype):
        n = self.pool.sum_(self.sums, self.n)
        var = (self.s2 - self.s1 * self.s1 / n) / (n - 1)
        var = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0))
        self.sums.zero_()
        self.n = 0
        return var.clamp(min=1e-10).to(dtype)


def _kinetic(r, inv_mass):
    return 0.5 * (r * r * inv_mass).sum(1)


def _reasonable_step_size(evaluate, kd, theta, logp, grad, inv_mass, eps, work, generator, pool):
    """Doubles / halves the step until the mean acceptance of one leapfrog step crosses 0.8 (Hoffman & Gelman, alg. 4)."""
    import torch
    q, r, g, lp, eps_c = work
    direction = 0
    for _ in range(40):
        r.copy_(torch.randn(theta.shape, dtype=theta.dtype, device=theta.device, generator=generator) / inv_mass.sqrt())
        h0 = -logp + _kinetic(r, inv_mass)
        q.copy_(theta)
        g.copy_(grad)
        eps_c.fill_(eps)
        kd(q, r, g, eps_c, inv_mass, 0.5, 1.0)
        evaluate(q, lp, g)
        kd(q, r, g, eps_c, inv_mass, 0.5, 0.0)
        delta = torch.nan_to_num(-lp + _kinetic(r, inv_mass) - h0, nan=float('inf'))
        accept = pool.mean((-delta).clamp(max=0).exp().mean())
        d = 1 if accept > 0.8 else -1
        if direction == 0:
            direction = d
        elif d != direction:
            break
        eps *= 2.0 ** direction
    return eps


# ---------------------------------------------------------------------------------------------------------------
# fixed-length HMC


def hmc_chains(evaluate, theta, draws: int, tune: int, n_leapfrog: int, step_size: float, target_accept: float,
               generator=None, kick_drift=None, pool: Optional[ChainPool] = None):
    """
    Fixed-length HMC for all chains at once on whatever device ``theta`` lives, one
    ``evaluate(q, logp_out, grad_out)`` call per leapfrog step for the whole batch.

    :param evaluate: fills ``logp_out [chains]`` and ``grad_out [chains, P]`` for the positions ``q [chains, P]``
        (on the GPU: ``skk_eval`` with device pointers, enqueued on the current stream).
    :param theta: start positions ``[chains, P]`` (modified in place).
    :param kick_drift: the integrator update between two evaluations (default :func:`kick_drift_torch`; on the GPU
        the library's fused kernel).
    :returns: ``(kept [chains, draws, P], kept_logp [chains, draws], accepted [chains], final step size)``
    """
    import torch

    kd = kick_drift or kick_drift_torch
    pool = pool or ChainPool()
    chains, P = theta.shape
    dev, tdt = theta.device, theta.dtype
    logp = torch.empty(chains, dtype=tdt, device=dev)
    grad = torch.empty(chains, P, dtype=tdt, device=dev)
    prop = torch.empty_like(theta)
    prop_logp = torch.empty_like(logp)
    prop_grad = torch.empty_like(grad)
    mom = torch.empty_like(theta)
    eps_c = torch.empty(chains, 1, dtype=tdt, device=dev)

    adapt = _DualAveraging(step_size, target_accept)
    inv_mass = torch.ones(P, dtype=tdt, device=dev)
    kept = torch.empty(chains, draws, P, dtype=tdt, device=dev)
    kept_logp = torch.empty(chains, draws, dtype=tdt, device=dev)
    accepted = torch.zeros(chains, dtype=torch.float64, device=dev)
    mass = _MassAdapt(P, dev, pool)
    windows = warmup_windows(tune)
    in_window = {it for a, b in windows for it in range(a, b)}
    window_end = {b - 1 for _, b in windows}

    evaluate(theta, logp, grad)
    for it in range(tune + draws):
        eps = adapt.eps if it < tune else adapt.eps_final
        eps_c.fill_(eps)
        mom.copy_(torch.randn(chains, P, dtype=tdt, device=dev, generator=generator) / inv_mass.sqrt())
        h0 = -logp + _kinetic(mom, inv_mass)
        prop.copy_(theta)
        prop_grad.copy_(grad)
        kd(prop, mom, prop_grad, eps_c, inv_mass, 0.5, 1.0)
        for step in range(n_leapfrog):
            evaluate(prop, prop_logp, prop_grad)
            if step < n_leapfrog - 1:
                kd(prop, mom, prop_grad, eps_c, inv_mass, 1.0, 1.0)
            else:
                kd(prop, mom, prop_grad, eps_c, inv_mass, 0.5, 0.0)
        h1 = -prop_logp + _kinetic(mom, inv_mass)
        log_alpha = torch.nan_to_num(h0 - h1, nan=-float('inf'))
        take = torch.rand(chains, dtype=tdt, device=dev, generator=generator).log() < log_alpha
        theta.copy_(torch.where(take[:, None], prop, theta))
        logp.copy_(torch.where(take, prop_logp, logp))
        grad.copy_(torch.where(take[:, None], prop_grad, grad))
        if it < tune:
            adapt.update(pool.mean(log_alpha.clamp(max=0).exp().mean()))
            if it in in_window:
                mass.add(theta)
            if it in window_end:
                inv_mass = mass.variance(tdt)
                adapt.restart(adapt.eps)
        else:
            k = it - tune
            kept[:, k] = theta
            kept_logp[:, k] = logp
            accepted += take.double()
    return kept, kept_logp, accepted, adapt.eps_final


# ---------------------------------------------------------------------------------------------------------------
# NUTS


def _ckpt_range(leaf: int):
    """
    Which momentum checkpoints the leaf ``leaf`` (0-based position inside a subtree built left to right) has to be
    tested against: the balanced sub-subtrees that END at this leaf start at checkpoints ``idx_min..idx_max``; an
    even leaf starts a new one at ``idx_max`` instead.  ``idx_max`` = number of set bits of ``leaf >> 1``,
    ``idx_max - idx_min + 1`` = number of trailing set bits of ``leaf``.
    """
    idx_max = bin(leaf >> 1).count('1')
    trailing = 0
    while (leaf >> trailing) & 1:
        trailing += 1
    return idx_max - trailing + 1, idx_max


def _turning(ra, rb, rsum, inv_mass):
    """Generalised U-turn criterion (Betancourt 2017) between the two ends of a trajectory with momentum sum ``rsum``."""
    rho = rsum - 0.5 * (ra + rb)
    return ((ra * inv_mass * rho).sum(1) <= 0) | ((rb * inv_mass * rho).sum(1) <= 0)


def nuts_chains(evaluate, theta, draws: int, tune: int, max_depth: int = 8, step_size: float = 0.1,
                target_accept: float = 0.8, generator=None, kick_drift=None, max_delta_energy: float = 1000.0,
                check_every: int = 4, pool: Optional[ChainPool] = None):
    """
    The No-U-Turn sampler with multinomial sampling of the trajectory (Betancourt 2017; what ``pm.sample`` runs per
    chain) for all chains in lock-step, one ``evaluate`` call per leapfrog step for the whole batch.

    Per transition every chain draws its momentum and then doubles its trajectory up to ``max_depth`` times, each
    time in its own random direction.  A new subtree of ``2**depth`` leaves is built leaf by leaf; the U-turn tests
    inside it use the checkpoint scheme of iterative NUTS (momenta of the sub-subtrees still open, ``max_depth``
    slots), whose indices depend on the leaf number only -- the same for every chain.  A chain whose subtree turns
    or diverges, or whose whole trajectory turns after the merge, is finished and masked from then on; the level
    ends early when no chain is building any more (checked every ``check_every`` leaves: one host read).

    :returns: ``(kept [chains, draws, P], kept_logp [chains, draws], stats, final step size)`` with ``stats`` =
        per draw ``depth``, ``n_steps``, ``diverging``, ``accept`` (mean acceptance statistic of the tree) and
        ``energy``, each ``[chains, draws]``.
    """
    import torch

    kd = kick_drift or kick_drift_torch
    pool = pool or ChainPool()
    chains, P = theta.shape
    dev, tdt = theta.device, theta.dtype

    def mat():
        return torch.empty(chains, P, dtype=tdt, device=dev)

    def vec(dtype=tdt):
        return torch.empty(chains, dtype=dtype, device=dev)

    def uniform():
        return torch.rand(chains, dtype=tdt, device=dev, generator=generator)

    def put(dst, mask, src):
        dst.copy_(torch.where(mask[:, None] if dst.dim() == 2 else mask, src, dst))

    logp, grad = vec(), mat()
    q, r, g, lp, eps_c = mat(), mat(), mat(), vec(), torch.empty(chains, 1, dtype=tdt, device=dev)
    q_left, r_left, g_left, q_right, r_right, g_right = mat(), mat(), mat(), mat(), mat(), mat()
    q_prop, g_prop, lp_prop, r_sum = mat(), mat(), vec(), mat()
    s_q, s_g, s_lp, s_rsum = mat(), mat(), vec(), mat()
    r_ckpt = torch.empty(max(max_depth, 1), chains, P, dtype=tdt, device=dev)
    rsum_ckpt = torch.empty_like(r_ckpt)
    neg_inf = torch.full((chains,), -float('inf'), dtype=tdt, device=dev)
    false = torch.zeros(chains, dtype=torch.bool, device=dev)

    inv_mass = torch.ones(P, dtype=tdt, device=dev)
    kept = torch.empty(chains, draws, P, dtype=tdt, device=dev)
    kept_logp = torch.empty(chains, draws, dtype=tdt, device=dev)
    stats = {'depth': torch.zeros(chains, draws, dtype=torch.int32, device=dev),
             'n_steps': torch.zeros(chains, draws, dtype=torch.int32, device=dev),
             'diverging': torch.zeros(chains, draws, dtype=torch.bool, device=dev),
             'accept': torch.zeros(chains, draws, dtype=tdt, device=dev),
             'energy': torch.zeros(chains, draws, dtype=tdt, device=dev)}
    mass = _MassAdapt(P, dev, pool)
    windows = warmup_windows(tune)
    in_window = {it for a, b in windows for it in range(a, b)}
    window_end = {b - 1 for _, b in windows}

    evaluate(theta, logp, grad)
    if tune > 0:
        step_size = _reasonable_step_size(evaluate, kd, theta, logp, grad, inv_mass, step_size,
                                          (q, r, g, lp, eps_c), generator, pool)
    adapt = _DualAveraging(step_size, target_accept)

    for it in range(tune + draws):
        eps = adapt.eps if it < tune else adapt.eps_final
        r0 = torch.randn(chains, P, dtype=tdt, device=dev, generator=generator) / inv_mass.sqrt()
        h0 = -logp + _kinetic(r0, inv_mass)
        for dst, src in ((q_left, theta), (q_right, theta), (q_prop, theta), (r_left, r0), (r_right, r0), (r_sum, r0),
                         (g_left, grad), (g_right, grad), (g_prop, grad), (lp_prop, logp)):
            dst.copy_(src)
        weight = torch.zeros(chains, dtype=tdt, device=dev)       # log of the trajectory's total weight
        active = torch.ones(chains, dtype=torch.bool, device=dev)
        diverged = false.clone()
        depth = torch.zeros(chains, dtype=torch.int32, device=dev)
        n_steps = torch.zeros(chains, dtype=torch.int32, device=dev)
        sum_accept = torch.zeros(chains, dtype=tdt, device=dev)

        for level in range(max_depth):
            if level > 0 and not bool(active.any()):
                break
            go_right = uniform() < 0.5
            eps_c.fill_(eps)
            eps_c.mul_(torch.where(go_right, 1.0, -1.0)[:, None])
            gr = go_right[:, None]
            q.copy_(torch.where(gr, q_right, q_left))
            r.copy_(torch.where(gr, r_right, r_left))
            g.copy_(torch.where(gr, g_right, g_left))
            s_weight, s_turn, s_div = neg_inf.clone(), false.clone(), false.clone()
            s_rsum.zero_()
            building = active.clone()
            n_leaves = 1 << level
            for leaf in range(n_leaves):
                kd(q, r, g, eps_c, inv_mass, 0.5, 1.0)
                evaluate(q, lp, g)
                kd(q, r, g, eps_c, inv_mass, 0.5, 0.0)
                energy = -lp + _kinetic(r, inv_mass)
                delta = torch.nan_to_num(energy - h0, nan=float('inf'))
                s_div |= building & (delta > max_delta_energy)
                # multinomial sampling inside the subtree: the leaf replaces the candidate with probability w / W
                new_weight = torch.logaddexp(s_weight, -delta)
                take = building & (uniform() < (-delta - new_weight).exp())
                put(s_q, take, q)
                put(s_g, take, g)
                put(s_lp, take, lp)
                s_weight = torch.where(building, new_weight, s_weight)
                put(s_rsum, building, s_rsum + r)
                sum_accept += torch.where(building, (-delta).clamp(max=0).exp(), torch.zeros_like(delta))
                n_steps += building.int()
                idx_min, idx_max = _ckpt_range(leaf)
                if leaf % 2 == 0:
                    r_ckpt[idx_max].copy_(r)
                    rsum_ckpt[idx_max].copy_(s_rsum)
                else:
                    turn = false
                    for i in range(idx_max, idx_min - 1, -1):
                        turn = turn | _turning(r_ckpt[i], r, s_rsum - rsum_ckpt[i] + r_ckpt[i], inv_mass)
                    s_turn |= building & turn
                building = building & ~s_turn & ~s_div
                if (leaf + 1) % check_every == 0 and leaf + 1 < n_leaves and not bool(building.any()):
                    break
            # merge: chains whose subtree is complete and valid
            ok = building
            take = ok & (uniform() < (s_weight - weight).exp())       # biased towards the new half
            put(q_prop, take, s_q)
            put(g_prop, take, s_g)
            put(lp_prop, take, s_lp)
            right, left = ok & go_right, ok & ~go_right
            put(q_right, right, q)
            put(r_right, right, r)
            put(g_right, right, g)
            put(q_left, left, q)
            put(r_left, left, r)
            put(g_left, left, g)
            weight = torch.where(ok, torch.logaddexp(weight, s_weight), weight)
            put(r_sum, ok, r_sum + s_rsum)
            depth += active.int()
            diverged |= active & s_div
            active = ok & ~_turning(r_left, r_right, r_sum, inv_mass)

        theta.copy_(q_prop)
        logp.copy_(lp_prop)
        grad.copy_(g_prop)
        accept = sum_accept / n_steps.clamp(min=1).to(tdt)
        if it < tune:
            adapt.update(pool.mean(accept.mean()))
            if it in in_window:
                mass.add(theta)
            if it in window_end:
                inv_mass = mass.variance(tdt)
                adapt.restart(_reasonable_step_size(evaluate, kd, theta, logp, grad, inv_mass, adapt.eps,
                                                    (q, r, g, lp, eps_c), generator, pool))
        else:
            k = it - tune
            kept[:, k] = theta
            kept_logp[:, k] = logp
            stats['depth'][:, k] = depth
            stats['n_steps'][:, k] = n_steps
            stats['diverging'][:, k] = diverged
            stats['accept'][:, k] = accept
            stats['energy'][:, k] = h0          # the transition's energy level (Hamiltonian at its start)
    return kept, kept_logp, stats, adapt.eps_final


# ---------------------------------------------------------------------------------------------------------------
# drivers on the CUDA evaluator


def chain_pool(chains: int):
    """
    ``(pool, chains of this process, rank)``: under ``torch.distributed`` with more than one rank (``torchrun``, one
    process per GPU) the chains are split evenly over the ranks -- every GPU holds all rows, no collective on the
    evaluation path (chains) -- and the ranks adapt step size and mass matrix as one ensemble.
    """
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            world = dist.get_world_size()
            if chains % world:
                raise ValueError(f'{chains} chains do not split evenly over {world} ranks')
            return DistributedChainPool(), chains // world, dist.get_rank()
    except ImportError:      # pragma: no cover
        pass
    return ChainPool(), chains, 0


def _host_setup(f, dtype, seed):
    """
    A model that is not one device plan -- a sum of plans added on the host (``PartitionedValueGrad``,
    ``SplitSigmaValueGrad``): the chains live in host memory and every leapfrog step is one batched call of the
    evaluator with host buffers (the rows are still read by the kernels; theta and the gradients cross PCIe).
    """
    import torch
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    gen = torch.Generator()
    gen.manual_seed(seed)

    def evaluate(q, lp, g):
        logp, grad = f(q.numpy())
        lp.copy_(torch.from_numpy(np.ascontiguousarray(logp)))
        g.copy_(torch.from_numpy(np.ascontiguousarray(grad)))

    return f, None, torch.device('cpu'), tdt, gen, None, evaluate, None


class _OnStream:
    """The plan's stream as torch's current one (device-resident chains); nothing for host-resident chains."""

    def __init__(self, stream, dev):
        self.stream, self.dev, self.ctx = stream, dev, None

    def __enter__(self):
        if self.stream is not None:
            import torch
            # everything -- the start positions included -- is created and consumed on the plan's stream: the library
            # launches there, and torch's caching allocator then knows which stream uses the buffers
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            import torch
            self.ctx.__exit__(*exc)
            torch.cuda.synchronize()
        return False


def _device_setup(model, chains, dtype, device, seed):
    import torch
    f = model.logp_dlogp_function(dtype=dtype, max_batch=chains, device=device)
    if not hasattr(f, 'skk'):
        return _host_setup(f, dtype, seed)
    skk = f.skk
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    stream = torch.cuda.ExternalStream(skk.stream, device=dev)

    def evaluate(q, lp, g):
        skk.eval_device(q.data_ptr(), chains, lp.data_ptr(), g.data_ptr(), sync=False)

    def kick_drift(q, r, g, eps, inv_mass, a, b):
        skk.leapfrog(q.data_ptr(), r.data_ptr(), g.data_ptr(), eps.data_ptr(), inv_mass.data_ptr(), chains, a, b)

    return f, skk, dev, tdt, gen, stream, evaluate, kick_drift


def sample_hmc(model, chains: int = 64, draws: int = 500, tune: int = 500, n_leapfrog: int = 16,
               step_size: float = 0.01, target_accept: float = 0.8, dtype: str = 'float64', seed: int = 0,
               device: Optional[int] = None, init_scale: float = 0.1, fused_leapfrog: bool = True) -> HMCResult:
    """
    :param model: a built :class:`sakkara_b200.model.B200Model` (or anything with ``plan`` and
        ``logp_dlogp_function``).
    :param fused_leapfrog: ``False`` takes the integrator update in torch operations (the A/B of the fused kernel).
    """
    import torch

    pool, chains, rank = chain_pool(chains)
    f, skk, dev, tdt, gen, stream, evaluate, kick_drift = _device_setup(model, chains, dtype, device, seed + 7919 * rank)
    with _OnStream(stream, dev):
        theta = init_scale * torch.randn(chains, f.size, dtype=tdt, device=dev, generator=gen)
        kept, kept_logp, accepted, eps = hmc_chains(evaluate, theta, draws, tune, n_leapfrog, step_size,
                                                    target_accept, generator=gen, pool=pool,
                                                    kick_drift=kick_drift if fused_leapfrog else None)
    result = HMCResult(draws=kept.cpu().numpy(), logp=kept_logp.cpu().numpy(),
                       accept_rate=(accepted / max(draws, 1)).cpu().numpy(), step_size=eps,
                       value_var_names=[b.value_name for b in model.plan.blocks])
    f.close()
    return result


def sample_nuts(model, chains: int = 64, draws: int = 500, tune: int = 500, max_depth: int = 8,
                step_size: float = 0.1, target_accept: float = 0.8, dtype: str = 'float64', seed: int = 0,
                device: Optional[int] = None, init_scale: float = 0.1, fused_leapfrog: bool = True) -> HMCResult:
    """
    NUTS for ``chains`` chains at once on the batch evaluator (``pm.sample``'s arguments where they apply:
    ``draws``, ``tune``, ``chains``, ``target_accept``; ``max_depth`` = ``max_treedepth``).  ``result.stats`` holds the
    sampler statistics per draw (``depth``, ``n_steps``, ``diverging``, ``accept``, ``energy``).
    """
    import torch

    pool, chains, rank = chain_pool(chains)
    f, skk, dev, tdt, gen, stream, evaluate, kick_drift = _device_setup(model, chains, dtype, device, seed + 7919 * rank)
    with _OnStream(stream, dev):
        theta = init_scale * torch.randn(chains, f.size, dtype=tdt, device=dev, generator=gen)
        kept, kept_logp, stats, eps = nuts_chains(evaluate, theta, draws, tune, max_depth=max_depth,
                                                  step_size=step_size, target_accept=target_accept, generator=gen,
                                                  pool=pool, kick_drift=kick_drift if fused_leapfrog else None)
    stats = {k: v.cpu().numpy() for k, v in stats.items()}
    result = HMCResult(draws=kept.cpu().numpy(), logp=kept_logp.cpu().numpy(),
                       accept_rate=stats['accept'].mean(1), step_size=eps,
                       value_var_names=[b.value_name for b in model.plan.blocks], stats=stats)
    f.close()
    return result


# ---------------------------------------------------------------------------------------------------------------
# convergence diagnostics (Vehtari, Gelman, Simpson, Carpenter, Buerkner 2021), NumPy


def _split(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    half = x.shape[1] // 2
    return np.concatenate([x[:, :half], x[:, x.shape[1] - half:]], axis=0)


def _rank_normalise(x: np.ndarray) -> np.ndarray:
    from scipy.stats import norm, rankdata
    ranks = rankdata(x, method='average').reshape(x.shape)
    return norm.ppf((ranks - 0.375) / (x.size + 0.25))


def _rhat(x: np.ndarray) -> float:
    n = x.shape[1]
    within = x.var(axis=1, ddof=1).mean()
    between = n * x.mean(axis=1).var(ddof=1)
    if within == 0:
        return float('nan')
    return float(np.sqrt(((n - 1) / n * within + between / n) / within))


def split_rhat(x: np.ndarray) -> float:
    """Rank-normalised split R-hat of ``x [chains, draws]`` (the maximum of the bulk and the folded version)."""
    s = _split(x)
    folded = np.abs(s - np.median(s))
    return max(_rhat(_rank_normalise(s)), _rhat(_rank_normalise(folded)))


def _ess(x: np.ndarray) -> float:
    m, n = x.shape
    if n < 4:
        return float('nan')
    size = 1 << int(np.ceil(np.log2(2 * n)))
    centred = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(centred, n=size, axis=1)
    acov = np.fft.irfft(f * np.conj(f), n=size, axis=1)[:, :n] / n
    chain_var = acov[:, 0] * n / (n - 1.0)
    mean_var = chain_var.mean()
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    if var_plus == 0:
        return float('nan')
    rho = np.zeros(n)
    rho[0] = 1.0
    mean_acov = acov.mean(axis=0)
    rho_even, rho_odd = 1.0, 1.0 - (mean_var - mean_acov[1]) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n - 4 and rho_even + rho_odd > 0:         # Geyer's initial positive sequence
        rho_even = 1.0 - (mean_var - mean_acov[t + 1]) / var_plus
        rho_odd = 1.0 - (mean_var - mean_acov[t + 2]) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1], rho[t + 2] = rho_even, rho_odd
        t += 2
    max_t = t
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 3:                               # ... made monotone
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    tau = -1.0 + 2.0 * rho[:max_t].sum() + rho[max_t + 1]
    tau = max(tau, 1.0 / np.log10(m * n))
    return float(m * n / tau)


def ess_bulk(x: np.ndarray) -> float:
    """Bulk effective sample size of ``x [chains, draws]`` (rank-normalised, split chains)."""
    return _ess(_rank_normalise(_split(x)))
