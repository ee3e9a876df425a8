"""
Multi-GPU plumbing (one process per GPU, ``torch.distributed``): row sharding and the exchange of
the NCCL unique id.  The data path itself -- one all-reduce(sum) of ``[dlogp, logp]`` per
evaluation -- runs inside libskk_b200.so on the plan's stream (``skk_comm_init``); nothing here
touches the per-evaluation path.  The reference has no counterpart.
"""
import os
from typing import Optional, Tuple

import numpy as np


def rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of a rank: sizes differ by at most one."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sorted_shard(plan, rank: int, world: int) -> np.ndarray:
    """
    Rows of this rank when the rows are sorted by the finest grouping level first and then cut
    into contiguous ranges: groups stay (almost) whole on one rank, so every rank
    sees long segments.  Returns row numbers of the unsharded plan, ascending within the shard.
    """
    from sakkara_b200.plan.kernel_plan import make_kernel_plan
    if world == 1:
        return np.arange(plan.n_rows)
    full = make_kernel_plan(plan, dtype='float64', structure_only=True)
    lo, hi = shard_bounds(plan.n_rows, rank, world)
    perm = np.arange(plan.n_rows) if full.perm is None else full.perm
    return np.sort(perm[lo:hi])


def exchange_unique_id(skk_plan, group=None) -> bytes:
    """Rank 0 creates the NCCL unique id; everybody receives it through torch.distributed."""
    import torch.distributed as dist
    box = [skk_plan.unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def _peer_access_everywhere(world: int, group=None) -> bool:
    """True when every rank sees all GPUs of the box and can map every peer's memory (NVLink / P2P)."""
    import torch
    import torch.distributed as dist
    ok = 1
    try:
        mine = torch.cuda.current_device()
        if torch.cuda.device_count() < world:
            ok = 0                              # ranks are isolated with CUDA_VISIBLE_DEVICES: cannot probe
        else:
            for other in range(world):
                if other != mine and not torch.cuda.can_device_access_peer(mine, other):
                    ok = 0
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32,
                        device='cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


def connect(skk_plan, rank: int, world: int, collective: str = 'auto', group=None) -> str:
    """
    Enable the per-evaluation all-reduce on a plan.  ``'peer'``: one-shot exchange over NVLink peer
    memory written by our own kernels (ranks of one box, <= 8); ``'nccl'``: ncclAllReduce;
    ``'auto'``: peer when possible.  Returns the collective in use.
    """
    if world == 1:
        if os.environ.get('SKK_FORCE_PEER'):       # measure the sharded pipeline on one GPU (self exchange)
            skk_plan.peer_init([skk_plan.peer_alloc(1)], 0, 1)
            return 'peer (self)'
        return 'none'
    import torch.distributed as dist
    if collective == 'auto':
        collective = 'peer' if world <= 8 and _peer_access_everywhere(world, group) else 'nccl'
    if collective == 'peer':
        # the slots of the exchange are laid out from (n_params, max_batch): every rank must agree, or a
        # rank would write into the wrong place of its peers' buffers
        shapes = [None] * world
        dist.all_gather_object(shapes, (int(skk_plan.n_params), int(skk_plan.max_batch), str(skk_plan.dtype)), group=group)
        if len(set(shapes)) != 1:
            raise ValueError(f'ranks disagree on (n_params, max_batch, dtype): {shapes}; every rank must build its '
                             f'plan from the same model (same theta layout) and the same max_batch')
        handles = [None] * world
        dist.all_gather_object(handles, skk_plan.peer_alloc(world), group=group)
        skk_plan.peer_init(handles, rank, world)
        dist.barrier(group=group)           # every rank has mapped every buffer before the first push
    elif collective == 'nccl':
        skk_plan.comm_init(exchange_unique_id(skk_plan, group), rank, world)
    else:
        raise ValueError(collective)
    return collective


def sharded_valuegrad(plan, dtype='float64', max_batch: int = 1, device: Optional[int] = None, group=None,
                      collective: str = 'auto'):
    """
    ValueGradFunction over this rank's rows with the all-reduce enabled: every rank passes the same
    (full) ``plan`` and receives the full ``(logp, dlogp)`` from each call.
    """
    import torch.distributed as dist
    from sakkara_b200.valuegrad import ValueGradFunction
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = sorted_shard(plan, rank, world)
    f = ValueGradFunction(plan, dtype=dtype, max_batch=max_batch, device=device, rows=rows, n_rows_total=plan.n_rows)
    connect(f.skk, rank, world, collective, group)
    return f
