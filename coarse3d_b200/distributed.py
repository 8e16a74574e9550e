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


class PeerExchange:
    """The all-reduce of the packed prototype sums fused with the EMA into ONE kernel over peer
    memory (c3d_proto_ema_apply_peers; NVLink / NVSwitch P2P loads and stores, no collective call
    on the step's critical path).  Set-up (once, collective over `group`): every rank allocates
    an exchange buffer, the CUDA IPC handles travel through torch.distributed, every rank maps
    the others' buffers.  Afterwards `apply()` is device-side only and can be captured in a CUDA
    graph.  Ranks must live on one node; all ranks must call `apply` the same number of times."""

    def __init__(self, n_classes, sub_protos, dim, device, group=None, timeout_s=20.0):
        import ctypes
        from ._lib import check, lib
        self.lib, self.check, self.group = lib, check, group
        self.rank, self.world = world(group)
        self.shape = (n_classes, sub_protos, dim)
        self.timeout_s = float(timeout_s)
        self.device = torch.device(device)
        self._own, self._mapped = None, []
        ok, why = 1, ""
        with torch.cuda.device(self.device):
            nbytes = lib.c3d_peer_exchange_bytes(n_classes, sub_protos, dim, self.world)
            if nbytes == 0:
                raise ValueError("peer exchange: bad shape or more than 8 ranks")
            handle = ctypes.create_string_buffer(64)
            try:
                own = ctypes.c_void_p()
                check(lib.c3d_peer_alloc(nbytes, ctypes.byref(own)))
                self._own = own
                check(lib.c3d_peer_export(own, handle))
            except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC unavailable in this container)
                ok, why = 0, repr(e)
            # every rank takes part in every collective below, whatever happened locally
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, (ok, handle.raw), group=group)
            else:
                handles[0] = (ok, handle.raw)
            ptrs = (ctypes.c_void_p * self.world)()
            if ok and all(h[0] for h in handles):
                try:
                    for r, (_, h) in enumerate(handles):
                        if r == self.rank:
                            ptrs[r] = self._own.value
                            continue
                        q = ctypes.c_void_p()
                        check(lib.c3d_peer_import(ctypes.create_string_buffer(h, 64), ctypes.byref(q)))
                        self._mapped.append(q)
                        ptrs[r] = q.value
                except Exception as e:  # noqa: BLE001
                    ok, why = 0, repr(e)
            else:
                ok = 0
            self._ptrs = ptrs
            self.state = torch.zeros(lib.c3d_peer_state_bytes(n_classes, sub_protos) // 4, dtype=torch.int32,
                                     device=self.device)
        if self.world > 1:
            # one decision for all ranks; also: every buffer is mapped everywhere before the first use
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = int(flag.item())
        self.available, self.why_not = bool(ok), why
        if not ok:
            self._release()

    def apply(self, prototypes, packed, momentum, ignore_label=0, out=None, normalised_out=None,
              seed_counters=None):
        """sum `packed` over the ranks (left in `packed`) and apply the EMA: the multi-GPU
        c3d_proto_ema_apply.  Bit-identical banks on every rank."""
        from .ops import _need_cuda, _p, _stream
        if not self.available:
            raise RuntimeError("peer exchange unavailable on this machine: %s" % self.why_not)
        _need_cuda(prototypes=prototypes, packed=packed)
        C, M, D = prototypes.shape
        if (C, M, D) != self.shape or packed.numel() != C * M * D + C * M or packed.dtype != torch.float32:
            raise ValueError("bank / payload shape differs from the exchange buffer's")
        if out is None:
            out = torch.empty_like(prototypes)
        self.check(self.lib.c3d_proto_ema_apply_peers(
            _p(prototypes), _p(packed), self._ptrs, self.rank, self.world, _p(self.state), C, M, D,
            int(ignore_label), float(momentum), _p(out), _p(normalised_out), _p(seed_counters),
            self.timeout_s, _stream()))
        return out

    def errors(self):
        """Bit r set: rank r's payload did not arrive within the timeout in some call (host sync)."""
        return int(self.state[3].item())

    def _release(self):
        with torch.cuda.device(self.device):
            for q in self._mapped:
                self.lib.c3d_peer_close(q)
            if self._own is not None:
                self.lib.c3d_peer_free(self._own)
        self._own, self._mapped = None, []

    def close(self):
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)   # nobody still reads a buffer that is about to go away
        self._release()


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
