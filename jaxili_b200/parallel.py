"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` for the exchange.

Training shards the global batch by contiguous row blocks; every rank computes
``sum_local d(-log p)/dp * (1/B_global)`` with the fused kernel, then ONE all-reduce (sum) of the flat
gradient buffer (P floats + the loss in the padding slot) makes the gradient global; the global-norm
clip and the Adam update run redundantly on every rank on the reduced gradient, so replicas stay
bit-identical.  log_prob evaluation and sampling shard by rows / observations with no collective.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n_rows, rank=None, world=None):
    """Contiguous row block [start, stop) of rank ``rank``; blocks differ by at most one row."""
    if rank is None or world is None:
        rank, world = world_info()
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def inv_global_batch(local_rows):
    """1 / B_global for equal-sized local batches (the trainer's drop_last batches)."""
    _, world = world_info()
    return 1.0 / (local_rows * world)


def allreduce_grad_(gbuf):
    """Sum the flat gradient(+loss) buffer over ranks, in place; no-op for a single process."""
    _, world = world_info()
    if world > 1:
        dist.all_reduce(gbuf, op=dist.ReduceOp.SUM)
    return gbuf


def dp_step(compute_local, apply_update, gbuf):
    """One data-parallel optimisation step: local fused fwd+bwd -> all-reduce -> replicated update."""
    compute_local(gbuf)
    allreduce_grad_(gbuf)
    apply_update(gbuf)
    return gbuf


def philox_row_offset(rank_start_row):
    """Sampling shards by rows: Philox subsequence = global row index, so samples do not depend on
    the number of GPUs."""
    return int(rank_start_row)


# gradients up to this size use the one-shot peer-memory all-reduce (latency-bound regime);
# larger ones (the wide config's 15 MB) stay on NCCL, whose ring/NVLS algorithms are bandwidth-optimal
PEER_ALLREDUCE_MAX_BYTES = 1 << 20


def setup_peer_allreduce(engine, step):
    """Map every rank's symmetric gradient buffer into every other rank (CUDA IPC over NVLink) and
    bind the device step counter ``step`` as the exchange epoch.  Returns True when the fused peer
    all-reduce is available for this engine.  Collective: every rank must call it."""
    rank, world = world_info()
    if world == 1 or world > 8 or engine.n_params * 4 > PEER_ALLREDUCE_MAX_BYTES:
        return False
    if os.environ.get("MAFB200_NO_PEER"):        # force the NCCL path (A/B measurements)
        return False
    if getattr(engine, "_dp_failed", False):
        return False
    if not getattr(engine, "_dp_ready", False):
        # mapping the peers' buffers needs CUDA IPC + peer access (NVLink / NVSwitch boxes); if any rank cannot
        # map them, every rank falls back to the NCCL all-reduce together
        ok = 1
        try:
            handle = engine.dp_init(rank, world)
        except Exception:
            handle, ok = b"\0" * 64, 0
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        if ok:
            try:
                engine.dp_connect(handles)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=engine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            engine._dp_failed = True
            return False
        engine._dp_ready = True
        engine._dp_step = None
    if engine._dp_step is None or engine._dp_step.data_ptr() != step.data_ptr():
        # a new step counter (new optimizer): epochs restart, forget the old ones on every rank
        torch.cuda.synchronize()
        dist.barrier()
        engine.dp_reset()
        engine.dp_bind_step(step)
        dist.barrier()
    return True


def reset_peer_epochs(engine):
    """Collective: forget the exchange epochs seen so far (the device step counter was rewound)."""
    torch.cuda.synchronize()
    dist.barrier()
    engine.dp_reset()
    dist.barrier()
