"""
Row-sharded evaluation over several GPUs (one process per GPU, NCCL all-reduce inside
libskk_b200.so): every rank must return the full (logp, dlogp), equal to the oracle.
Skipped unless at least two CUDA devices are visible.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, dtype, collective, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from helpers import assert_parity
        from sakkara_b200 import synth
        from sakkara_b200.distributed import sharded_valuegrad
        results = []
        for plan in (synth.poisson_plan(synth.poisson_columns(200_000, 400, 30, seed=5)[0]),
                     synth.linreg_plan(synth.linreg_columns(150_000, 300, seed=5)[0]),
                     synth.logistic_plan(synth.logistic_columns(120_000, 20, 30, seed=5)[0])):
            f = sharded_valuegrad(plan, dtype=dtype, max_batch=4, device=rank, collective=collective)
            assert f.kernel_plan.n_rows < plan.n_rows
            theta = synth.start_point(plan, 3, seed=1, scale=0.3, dtype=dtype)
            logp, grad = f(theta)
            for b in range(3):
                results.append(assert_parity(logp[b], grad[b], plan, theta[b], dtype))
            for _ in range(3):                  # repeated calls reuse both parities of the peer buffers
                again = f(theta)
                assert np.array_equal(again[1], grad) and np.array_equal(again[0], logp)
            f.close()
        out[rank] = max(max(e) for e in results)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('collective', ['peer', 'nccl'])
@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_row_sharded_parity(dtype, collective):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs at least two GPUs')
    with mp.Manager() as manager:
        out = manager.dict()
        mp.spawn(_worker, args=(world, 29650 + os.getpid() % 300, dtype, collective, out), nprocs=world, join=True)
        assert sorted(out.keys()) == list(range(world))
