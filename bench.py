#!/usr/bin/env python
"""
bench.py -- logp+dlogp throughput of the B200 evaluator on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl b200|reference]

One JSON line on stdout (rank 0).  A *step* is one logp+dlogp evaluation of the whole batch of
parameter vectors over all rows of the workload.

  value      row-evaluations / s (= rows_total * batch / step time), inputs resident in HBM; step time
             from CUDA events on the library's stream, L2 flushed between steps, max over ranks
  e2e        the same metric through the public call with HOST buffers (theta H2D, logp+dlogp D2H
             and the stream synchronisation inside the timed region)
  check      one evaluation of the build that is timed, against the CPU oracle (oracle/logp_c.c in fp64
             over this rank's rows, summed over the ranks with torch.distributed, + the priors) -- in
             every line, for N = 1 and for sharded runs
  roofline   stored bytes the row kernel reads per launch / its event-timed duration, against the
             measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the fused OpenMP C restatement (oracle/logp_c.c) on the same rows, all host cores
  configs    (default workload only) the other BASELINE configurations as sub-records with the same keys:
             c2 and c3 on one GPU (c3 is the north_star's single-GPU roofline case), c5 at every N
             (chains sharded, no collective)

Default workload c4: Poisson-log GLM with crossed random effects, 100M rows, 10k x 1k groups, fp32 --
the configuration BASELINE.json quotes at 1/2/4/8 GPUs; rows are sharded contiguously by the sorted
primary factor and every evaluation ends in one all-reduce of [dlogp, logp] (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: rows, dtype, batch (total), description
    'c2': dict(rows=1_000_000, dtype='float64', batch=4, shard='rows', canonical=20,
               workload='hierarchical linear regression: 1M rows, 1,000 groups, fp64, 4 chains'),
    'c3': dict(rows=10_000_000, dtype='float32', batch=1, shard='rows', canonical=16,
               workload='hierarchical logistic regression (Bernoulli-logit): 10M rows, nested 100x100 groups, fp32'),
    'c4': dict(rows=100_000_000, dtype='float32', batch=1, shard='rows', canonical=16,
               workload='Poisson-log GLM with crossed random effects: 100M rows, 10k+1k groups, fp32, row-sharded'),
    'c5': dict(rows=1_000_000, dtype='float64', batch=1024, shard='chains', canonical=20,
               workload='batched multi-chain evaluation: 1,024 chains x 1M rows x 1k groups, fp64, chains sharded'),
}
RTOL = {'float64': 1e-10, 'float32': 1e-4}     # BASELINE.json north_star


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_plan(config: str, rows: int, rank: int, world: int, seed: int = 1):
    """
    This rank's rows of the synthetic workload as a ModelPlan.  Group members carry their global
    numbers, so every rank of a sharded run has the same theta layout (P is the same everywhere) and
    the all-reduced gradient means the same thing on all ranks.
    """
    from sakkara_b200 import synth
    if config == 'c5':                      # every rank holds all rows; the chains are sharded
        cols, _ = synth.linreg_columns(rows, 1000, seed=seed)
        return synth.linreg_plan(cols, groups=1000)
    share = rows // world + (1 if rank < rows % world else 0)
    if config == 'c2':
        cols, _ = synth.linreg_columns(rows, 1000, seed=seed)
        lo = sum(rows // world + (1 if r < rows % world else 0) for r in range(rank))
        cols = {k: v[lo:lo + share] for k, v in cols.items()}
        return synth.linreg_plan(cols, groups=1000)
    if config == 'c3':
        g1, per = 100, 100
        span = (g1 * per) // world
        cols, _ = synth.logistic_columns(share, g1, per, seed=seed, block=rank,
                                         g2_range=(rank * span, (rank + 1) * span if rank < world - 1 else g1 * per))
        return synth.logistic_plan(cols, groups=(g1, per))
    if config == 'c4':
        a, b = 10_000, 1_000
        span = a // world
        parts = []
        chunk = 12_500_000
        for start in range(0, share, chunk):
            c, _ = synth.poisson_columns(min(chunk, share - start), a, b, seed=seed, block=rank * 1000 + start // chunk,
                                         a_range=(rank * span, (rank + 1) * span if rank < world - 1 else a))
            parts.append(c)
        cols = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        return synth.poisson_plan(cols, groups=(a, b))
    raise ValueError(config)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.device}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def profiled_traffic(config: str, rows_per_gpu: int):
    """
    DRAM bytes of the row kernel per launch from the newest committed `ncu --set full` capture of this shape
    (profiles/*traffic.json; cannot be measured inside a bench run), with the capture it comes from -- or None.
    """
    best = None
    pdir = os.path.join(ROOT, 'profiles')
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith('traffic.json'):
            with open(os.path.join(pdir, name)) as f:
                table = json.load(f)
            entry = table.get(f'{config}:{rows_per_gpu}')
            if entry:
                best = (entry['read'] + entry['write'], f'profiles/{name}: {entry.get("source", "ncu --set full")}')
    return best


def cpu_baseline(plan, config: str, dtype: str, seconds: float = 10.0):
    """Fused OpenMP C restatement on the rows of this run, all host cores, for about `seconds`."""
    from oracle.logp_c import COracle
    from sakkara_b200 import synth
    theta = synth.start_point(plan, 1, dtype=dtype)[0]
    cores = os.cpu_count() or 1
    f = COracle(plan, dtype=dtype, threads=cores)
    f(theta)
    n, t0 = 0, time.perf_counter()
    while True:
        f(theta)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 5000:
            break
    return {'value': plan.n_rows * n / dt, 'unit': 'row-evals/s', 'cores': cores, 'kind': 'port',
            'sample': f'all {plan.n_rows} rows of the {config} workload, {n} evaluations of oracle/logp_c.c (fused OpenMP '
                      f'loop, {dtype}; a port -- PyMC/PyTensor are not installable) in {dt:.1f} s'}


def run_reference(args):
    """CPU arm: the oracle port on the host cores (PyMC/PyTensor are not installable here), on the GPU arm's rows."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    from oracle.logp_c import COracle
    from sakkara_b200 import synth
    rows = args.rows or cfg['rows']
    if args.ref_rows:
        rows = min(rows, args.ref_rows)
    plan = make_plan(args.config, rows, 0, 1)
    batch = cfg['batch']
    theta = synth.start_point(plan, batch, dtype=cfg['dtype'])
    cores = os.cpu_count() or 1
    f = COracle(plan, dtype=cfg['dtype'], threads=cores)
    # bound the chains per step so that the whole run stays within a few minutes (c5: 1,024 chains)
    t0 = time.perf_counter()
    f(theta[0])
    one = time.perf_counter() - t0
    chains = max(1, min(batch, int(120.0 / max(one, 1e-6) / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        for th in theta[:chains]:
            f(th)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for th in theta[:chains]:
            f(th)
    dt = time.perf_counter() - t0
    value = rows * chains * args.steps / dt
    sample = (f'{rows} rows x {chains} of {batch} parameter vectors of the {args.config} workload per step, '
              f'oracle/logp_c.c fused OpenMP loop, {cfg["dtype"]}')
    print(json.dumps({
        'impl': 'reference', 'metric': 'logp_dlogp_row_evals_per_s', 'value': value, 'unit': 'row-evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64' if cfg['dtype'] == 'float64' else 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'rows_total': rows, 'rows_per_step': rows, 'batch': chains},
        'cpu_baseline': {'value': value, 'unit': 'row-evals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'row-evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def oracle_check(f, plan, theta, dtype, rank, world, dev, max_chains=4):
    """
    One evaluation of the library as it is about to be timed, against the CPU oracle: the likelihood of this
    rank's rows by oracle/logp_c.c in fp64 (summed over the ranks of a row-sharded run), plus the priors by
    oracle/logp_numpy.py.  Error metric of SURVEY.md 8d for logp (relative to max(|ref|, 1)); gradient entries
    relative to max(|ref_j|, 1e-3 max|ref|).  Returns the `check` record of the JSON line.
    """
    import ctypes
    import torch
    import torch.distributed as dist
    from oracle.logp_c import COracle, priors_only
    from oracle.logp_numpy import logp_dlogp
    P = plan.n_params
    lp, g = f(theta)
    lp, g = np.atleast_1d(lp), np.atleast_2d(g)
    shard = COracle(plan, dtype='float64', threads=max(1, (os.cpu_count() or 1) // world))
    err_l = err_g = 0.0
    n = min(max_chains, theta.shape[0])
    for b in range(n):
        th64 = theta[b].astype(np.float64)
        grad = np.zeros(P)
        d = shard.desc
        if plan.likelihood.family == 'Normal':
            d.sigma = float(np.exp(th64[d.sigma_theta_index])) if d.sigma_theta_index >= 0 \
                else plan.likelihood.sigma_const
        lik = shard.fn(ctypes.byref(d), th64.ctypes.data, grad.ctypes.data, shard.threads) + shard.const
        vec = torch.from_numpy(np.concatenate([grad, [lik]])).to(dev)
        if world > 1:
            dist.all_reduce(vec)
        vec = vec.cpu().numpy()
        pl_, pg_ = logp_dlogp(priors_only(plan), th64)
        ref_l, ref_g = vec[-1] + pl_, vec[:-1] + pg_
        err_l = max(err_l, abs(float(lp[b]) - ref_l) / max(abs(ref_l), 1.0))
        scale = np.maximum(np.abs(ref_g), np.abs(ref_g).max() * 1e-3 + 1e-300)
        err_g = max(err_g, float(np.max(np.abs(np.asarray(g[b], dtype=np.float64) - ref_g) / scale)))
    tol = RTOL[dtype]
    errs = torch.tensor([err_l, err_g], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)      # every rank holds the full result: the worst one counts
    err_l, err_g = (float(v) for v in errs.cpu())
    ok = bool(np.isfinite(err_l) and np.isfinite(err_g) and err_l <= tol and err_g <= tol)
    if rank == 0:
        log(f'[check] logp={float(lp[0]):.6e} rel.err logp={err_l:.2e} grad={err_g:.2e} tol={tol:g} '
            f'{"ok" if ok else "PARITY FAILURE"} (vs oracle/logp_c.c fp64 summed over {world} rank(s), {n} chain(s))')
    return {'err_logp': err_l, 'err_grad': err_g, 'tol': tol, 'ok': ok, 'rows': int(plan.n_rows), 'ranks': world,
            'chains': n, 'oracle': 'oracle/logp_c.c (fp64 likelihood) + oracle/logp_numpy.py (priors)'}


def run_config(config, args, rank, world, local, steps, warmup, rows_override=0, with_cpu=True):
    """Plan, parity check and all timings of one configuration; returns the record (complete on rank 0)."""
    import torch
    import torch.distributed as dist
    from sakkara_b200 import synth
    from sakkara_b200.distributed import connect
    from sakkara_b200.valuegrad import ValueGradFunction

    dev = torch.device('cuda', local)
    cfg = CONFIGS[config]
    rows_total = rows_override or cfg['rows']
    dtype, batch = cfg['dtype'], cfg['batch']
    chains_sharded = cfg['shard'] == 'chains'
    batch_total = batch
    if chains_sharded:
        batch = batch_total // world + (1 if rank < batch_total % world else 0)
    emulate = args.emulate_world if config == args.config else 1
    t0 = time.time()
    if emulate > 1 and chains_sharded:
        if world != 1:
            raise SystemExit('--emulate-world runs on one GPU')
        batch = batch_total // emulate                   # rank 0's chains of a W-rank run; nothing to exchange
        plan = make_plan(config, rows_total, 0, 1)
    elif emulate > 1:
        if world != 1:
            raise SystemExit('--emulate-world runs on one GPU')
        os.environ['SKK_FORCE_PEER'] = '1'
        plan = make_plan(config, rows_total, 0, emulate)
    else:
        if args.emulate_world > 1:
            os.environ.pop('SKK_FORCE_PEER', None)       # (set by the emulated main record, not for the sub-records)
        plan = make_plan(config, rows_total, rank, world)
    t_gen = time.time() - t0
    f = ValueGradFunction(plan, dtype=dtype, max_batch=batch, device=local,
                          n_rows_total=plan.n_rows if chains_sharded else rows_total)
    t0 = time.time()
    kp = f.kernel_plan
    t_plan = time.time() - t0
    t0 = time.time()
    skk = f.skk
    t_upload = time.time() - t0
    st0 = skk.stats()
    if rank == 0:
        log(f'[bench] {config}: rows/rank={plan.n_rows} P={plan.n_params} gen {t_gen:.1f}s plan {t_plan:.1f}s '
            f'upload {t_upload:.1f}s grid={st0["grid"]}x{st0["block"]} smem={st0["smem_bytes"]} '
            f'bytes/row={st0["stored_bytes_per_row"]}')
        log(kp.describe())

    collective = 'none (chains are independent)' if chains_sharded else connect(skk, rank, world, args.collective)

    theta = synth.start_point(plan, batch, seed=rank if chains_sharded else 0, dtype=dtype)
    P = plan.n_params
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    th_dev = torch.from_numpy(theta).to(dev)
    logp_dev = torch.empty(batch, dtype=tdt, device=dev)
    grad_dev = torch.empty(batch, P, dtype=tdt, device=dev)
    stream = torch.cuda.ExternalStream(skk.stream, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    check = None
    if not args.no_check:
        check = oracle_check(f, plan, theta, dtype, rank, 1 if chains_sharded else world, dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_steps(n, timed, cold=True):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        with torch.cuda.stream(stream):
            for i in range(n):
                if cold:
                    flush.zero_()                               # evict the row data from L2
                starts[i].record(stream)
                skk.eval_device(th_dev.data_ptr(), batch, logp_dev.data_ptr(), grad_dev.data_ptr(), sync=False)
                ends[i].record(stream)
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in zip(starts, ends)] if timed else None

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps -------------------------
    sampler = ClockSampler(local)       # samples every 20 ms from the warm-up to the end of all timed loops
    sampler.start()
    device_steps(warmup, False)
    launches0 = skk.stats()['launches']
    barrier()
    wall0 = time.perf_counter()
    step_ms = device_steps(steps, True)
    barrier()
    wall = time.perf_counter() - wall0
    launches = skk.stats()['launches'] - launches0
    ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / steps
    work = rows_total * (batch_total if chains_sharded else batch)      # row-evaluations per step, all ranks
    value = work / (ms_per_step * 1e-3)
    # the same with the data left in L2 (what back-to-back leapfrog steps of a sampler see when the rows fit)
    warm_ms = None
    if kp.n_rows * st0['stored_bytes_per_row'] < 100e6:
        device_steps(2, False, cold=False)
        barrier()
        w = torch.tensor([float(np.mean(device_steps(steps, True, cold=False)))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        warm_ms = float(w.item())

    # ---- the row kernel alone (events inside the library, around the launch) -------------------------
    kern = []
    with torch.cuda.stream(stream):
        for _ in range(max(3, min(steps, 10))):
            flush.zero_()
            torch.cuda.synchronize()
            skk.eval_device(th_dev.data_ptr(), batch, logp_dev.data_ptr(), grad_dev.data_ptr(), sync=True,
                            time_kernels=True)
            kern.append(skk.stats()['last_row_kernel_ms'])
    kern_ms = float(np.mean(kern))
    esize = 8 if dtype == 'float64' else 4
    chain_path = bool(skk.stats()['row_kernel_launches'] and batch >= 32 and kp.n_secondary == 0)
    bt = batch if chain_path else (4 if batch >= 2 else 1)
    launches_per_eval = 1 if chain_path else ((batch + 3) // 4 if batch >= 2 else 1)
    alg_bytes = plan.n_rows * st0['stored_bytes_per_row'] + bt * (2 * P * esize + esize)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    canonical = cfg['canonical'] * plan.n_rows / (kern_ms * 1e-3) / 1e9
    traffic = profiled_traffic(config, plan.n_rows)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic[0] if traffic else None,
                'traffic_source': traffic[1] if traffic else 'no ncu capture of this shape (cannot be measured inside a bench run)',
                'kernel': 'skk_row_kernel', 'kernel_ms': kern_ms, 'bytes_per_launch': alg_bytes,
                'launches_per_step': launches_per_eval, 'peak_source': peak_src,
                'canonical': {'bytes_per_row': cfg['canonical'], 'achieved': canonical, 'frac': canonical / peak,
                              'note': 'SURVEY 8d canonical layout (floatX y, int32 codes); the plan stores fewer bytes'}}

    if chain_path:
        # K2 is ALU-bound: 11 flop per row-evaluation for the Normal/identity model, six fp64
        # instructions per row-evaluation in the SASS; the denominator is the measured fp64 FMA rate of this GPU
        from sakkara_b200.runtime import alu_peak
        alu = alu_peak(local)
        key = 'dfma_per_s' if dtype == 'float64' else 'ffma_per_s'
        instr_rate = 6 * plan.n_rows * batch / (kern_ms * 1e-3)
        roofline = {'bound': 'alu', 'achieved': 11 * plan.n_rows * batch / (kern_ms * 1e-3) / 1e12,
                    'peak': 2 * alu[key] / 1e12, 'unit': 'TFLOP/s', 'frac': instr_rate / alu[key], 'traffic': None,
                    'kernel': 'skk_chain_kernel', 'kernel_ms': kern_ms, 'launches_per_step': 1,
                    'peak_source': f'measured in this run: skk_alu_peak ({key} x 2 flop)',
                    'note': 'frac = fp64 pipe instructions issued (6 per row-evaluation: 2 DADD + 3 DFMA + 1 DADD) / measured '
                            'DFMA instruction rate; achieved counts the 11 algorithmic flop per row-evaluation',
                    'alu_peaks': alu, 'hbm': {'achieved': achieved, 'peak': peak, 'frac': achieved / peak}}

    # ---- end to end through the public call with host buffers ---------------------------------------------
    for _ in range(warmup):
        f(theta)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        f(theta)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    e2e_value = work * steps / float(t_e2e.item())
    e2e = {'value': e2e_value, 'unit': 'row-evals/s', 'h2d_bytes_per_step': batch * P * esize,
           'd2h_bytes_per_step': batch * (P + 1) * esize, 'ms_per_step': 1e3 * float(t_e2e.item()) / steps}

    out = {
        'metric': 'logp_dlogp_row_evals_per_s', 'value': value, 'unit': 'row-evals/s', 'n_gpus': world,
        'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64' if dtype == 'float64' else 'f32',
        'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'rows_total': rows_total, 'rows_per_gpu': plan.n_rows,
                   'n_params': P, 'batch': batch_total if chains_sharded else batch, 'batch_per_gpu': batch,
                   'l2': 'flushed between steps (256 MiB memset)',
                   'parallelism': (f'chains x{world}' if chains_sharded else f'rows x{world}') if world > 1 else 'single GPU',
                   'collective': collective, 'grid': st0['grid'], 'block': st0['block'], 'smem_bytes': st0['smem_bytes']},
        'evals_per_s': (batch_total if chains_sharded else batch) / (ms_per_step * 1e-3), 'clocks': clocks, 'e2e': e2e,
        'gpu_launches': int(launches), 'check': check, 'roofline': roofline, 'wall_s_timed_region': wall,
    }
    if emulate > 1:
        out['config']['emulated_world'] = emulate
        out['config']['note'] = ('one GPU runs the shard of rank 0 through the sharded pipeline (self exchange); '
                                 'value is not a job throughput')
    if warm_ms is not None:
        out['warm_l2'] = {'ms_per_step': warm_ms, 'value': work / (warm_ms * 1e-3),
                          'note': 'no L2 flush between steps: the row data stays in the 126 MB L2'}
    if rank == 0 and world == 1 and with_cpu and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(plan, config, dtype, seconds=args.cpu_seconds)
    else:
        out['cpu_baseline'] = None
    f.close()
    del flush, th_dev, grad_dev, logp_dev
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--config', default=os.environ.get('SKK_BENCH_CONFIG', 'c4'), choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rows', type=int, default=0, help='override the number of rows (testing)')
    ap.add_argument('--ref-rows', type=int, default=0, help='reference arm: cap on the rows per step (0: the full workload)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--collective', default='auto', choices=['auto', 'peer', 'nccl'])
    ap.add_argument('--check', action='store_true', help='(kept for compatibility: the oracle check now always runs)')
    ap.add_argument('--no-check', action='store_true', help='skip the oracle check (profiling runs)')
    ap.add_argument('--no-sub', action='store_true', help='default workload only, without the c2 / c3 / c5 sub-records')
    ap.add_argument('--emulate-world', type=int, default=1,
                    help='one GPU runs the shard of rank 0 of W ranks through the sharded pipeline (self exchange)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    out = run_config(args.config, args, rank, world, local, args.steps, args.warmup, rows_override=args.rows)
    # the other BASELINE configurations, as sub-records of the default line
    if args.config == 'c4' and not args.no_sub and not args.rows and args.emulate_world == 1:
        subs = ['c3', 'c2', 'c5'] if world == 1 else ['c5']
        out['configs'] = {}
        for c in subs:
            try:
                rec = run_config(c, args, rank, world, local, args.steps, args.warmup, with_cpu=False)
                for k in ('metric', 'unit', 'higher_is_better', 'vs_baseline', 'data', 'steps', 'warmup', 'cpu_baseline',
                          'wall_s_timed_region', 'n_gpus'):
                    rec.pop(k, None)
                out['configs'][c] = rec
            except Exception as err:      # a sub-record must not take the headline down with it
                out['configs'][c] = {'error': f'{type(err).__name__}: {err}'}
                log(f'[bench] sub-record {c} failed: {err!r}')
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
