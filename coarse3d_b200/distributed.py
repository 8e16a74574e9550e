"""Multi-GPU plumbing of the hot path: scans shard across ranks, and the only
exchange is one all-reduce of the packed prototype sums and counts.

The reference's only hot-path collective is `dist.all_reduce(protos / world)`
after per-rank EMAs (salsanext_proto.py:397-400).  Here ranks sum the
[K*D sums | K counts] payload BEFORE the EMA, so every rank applies the identical
update and the result equals a single-process run on the concatenated batch.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def world(group=None):
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_scans(n_scans, rank, world_size):
    """Contiguous block of scan indices owned by `rank` (what DistributedSampler
    gives the reference, trainer.py:300-306).  Sizes differ by at most one."""
    base, rem = divmod(n_scans, world_size)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def allreduce_packed(packed: torch.Tensor, group=None, async_op=False):
    """Sum the packed [K*D | K] payload over ranks, in place.  Counts are small
    integers in float32 and sum exactly; the feature sums are reduced in the
    collective's order, identically on every rank."""
    if world(group)[1] == 1:
        return None
    return dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def prototype_update(embedding, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b, momentum,
                     ignore_label=0, gumbel=None, assign_mode=None, seed=None, max_rows=None,
                     want_target=False, group=None, workspace=None, packed=None, out=None):
    """accumulate -> all-reduce -> apply.  Returns (new prototypes, EmaAccum)."""
    from . import ops
    acc = ops.proto_ema_accumulate(embedding, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                                   ignore_label=ignore_label, gumbel=gumbel, assign_mode=assign_mode,
                                   seed=seed, max_rows=max_rows, want_target=want_target,
                                   workspace=workspace, packed=packed)
    allreduce_packed(acc.packed, group)
    new = ops.proto_ema_apply(prototypes, acc.packed, momentum, ignore_label, out=out)
    return new, acc
