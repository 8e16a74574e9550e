"""Multi-GPU plumbing of the hot path: scans shard across ranks, and the only
exchange is one all-reduce of the packed prototype sums and counts.

Two synchronisation rules (`sync`):

* "sum" (default, BASELINE.json north_star): ranks sum the [K*D sums | K counts] payload
  BEFORE the EMA, so every rank applies one identical update and the banks stay bit-identical.
  Guarantee, exactly: the result equals `apply(sum_r accumulate(shard_r))`, i.e. one EMA from
  the summed per-rank sums.  The per-rank sums themselves are rank-local: the Sinkhorn
  assignment normalises over the rows of a class ON THAT RANK (the reference's
  `distributed_sinkhorn` communicates nothing either, sinkhorn.py:5-33), so the result is NOT
  that of a single process running the concatenated batch.
* "average" (the reference's rule, salsanext_proto.py:397-400): every rank applies its own EMA
  from its local sums, then the banks are averaged, `all_reduce(protos / world)`, and not
  re-normalised.

On one rank both coincide with the reference.  Works with any torch.distributed backend
(NCCL on the GPUs, gloo in CPU tests).
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
                     want_target=False, group=None, workspace=None, packed=None, out=None, sync="sum"):
    """accumulate -> all-reduce -> apply ("sum"), or accumulate -> apply -> average ("average").
    Returns (new prototypes, EmaAccum)."""
    from . import ops
    acc = ops.proto_ema_accumulate(embedding, label, prototypes, ln_d_w, ln_d_b, ln_c_w, ln_c_b,
                                   ignore_label=ignore_label, gumbel=gumbel, assign_mode=assign_mode,
                                   seed=seed, max_rows=max_rows, want_target=want_target,
                                   workspace=workspace, packed=packed)
    return finish_update(prototypes, acc, momentum, ignore_label, group, out, sync), acc


def finish_update(prototypes, acc, momentum, ignore_label=0, group=None, out=None, sync="sum"):
    """The part of the update after the local accumulation (see the module docstring)."""
    from . import ops
    if sync == "sum":
        allreduce_packed(acc.packed, group)
        return ops.proto_ema_apply(prototypes, acc.packed, momentum, ignore_label, out=out)
    if sync != "average":
        raise ValueError("sync must be 'sum' or 'average'")
    new = ops.proto_ema_apply(prototypes, acc.packed, momentum, ignore_label, out=out)
    n = world(group)[1]
    if n > 1:                                   # salsanext_proto.py:397-400
        new.div_(n)
        dist.all_reduce(new, op=dist.ReduceOp.SUM, group=group)
    return new
