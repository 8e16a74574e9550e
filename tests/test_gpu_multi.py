"""GPU, world size 2 over NCCL: the multi-GPU rule of the prototype update.

What is guaranteed (coarse3d_b200/distributed.py), and asserted here ON HARDWARE:
  * sync="sum": after the update every rank holds the bit-identical bank, and that bank equals
    a single process that accumulates each rank's shard separately, adds the packed payloads and
    applies ONE EMA (the Sinkhorn assignment stays rank-local, as in the reference);
  * sync="average" (the reference's rule, salsanext_proto.py:397-400): the mean of the per-rank
    post-EMA banks;
  * the step pipeline with the all-reduce captured in its CUDA graph keeps the banks identical.
Needs two CUDA devices (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(rank, B=2, D=32, H=16, W=128, C=7, M=4):
    g = torch.Generator().manual_seed(100 + rank)
    centers = torch.nn.functional.normalize(torch.randn(C, D, generator=torch.Generator().manual_seed(5)), dim=-1)
    label = torch.randint(1, C, (B, H, W), generator=g)
    emb = torch.nn.functional.normalize(
        torch.randn(B, D, H, W, generator=g) * 0.7 + 1.5 * centers[label].permute(0, 3, 1, 2), dim=1)
    label = label * (torch.rand(B, H, W, generator=g) < 0.05)
    return emb.contiguous(), label


def _bank0(C=7, M=4, D=32):
    g = torch.Generator().manual_seed(9)
    centers = torch.nn.functional.normalize(torch.randn(C, D, generator=torch.Generator().manual_seed(5)), dim=-1)
    return (torch.randn(C, M, D, generator=g) * 0.02 + centers[:, None, :] * 0.05).contiguous()


def _worker(rank, world, port, result):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from coarse3d_b200 import distributed, ops, synth
        from coarse3d_b200.pipeline import HotPathStep
        C, M, D = 7, 4, 32
        ln = [torch.ones(D, device=dev), torch.zeros(D, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev)]
        emb, label = (t.to(dev) for t in _problem(rank))
        p0 = _bank0().to(dev)
        for sync in ("sum", "average"):
            new, _ = distributed.prototype_update(emb, label, p0, *ln, 0.9, assign_mode=ops.ASSIGN_ARGMAX, sync=sync)
            banks = [torch.empty_like(new) for _ in range(world)]
            dist.all_gather(banks, new)
            assert all(torch.equal(b, banks[0]) for b in banks), "banks differ across ranks (%s)" % sync
            # single-process statement of the rule, on this rank's GPU
            accs = []
            for r in range(world):
                e, l = (t.to(dev) for t in _problem(r))
                accs.append(ops.proto_ema_accumulate(e, l, p0, *ln, assign_mode=ops.ASSIGN_ARGMAX))
            if sync == "sum":
                packed = accs[0].packed.clone()
                for a in accs[1:]:
                    packed += a.packed
                want = ops.proto_ema_apply(p0, packed, 0.9)
                assert torch.equal(new, want), "sum-before-EMA != apply(sum of per-shard payloads)"
                # ... and it is NOT the update of one process on the concatenated batch in general
            else:
                want = sum(ops.proto_ema_apply(p0, a.packed, 0.9) / world for a in accs)
                assert (new - want).abs().max() <= 1e-7, "average mode != mean of per-rank post-EMA banks"
        # the fused form: all-reduce + EMA as ONE kernel over peer memory (CUDA IPC, NVLink P2P)
        px = distributed.PeerExchange(C, M, D, dev)
        cur = p0.clone()
        bank_n = torch.empty_like(cur)
        want = p0.clone()
        for it in range(5):                       # both buffer halves, several times
            acc = ops.proto_ema_accumulate(emb, label, cur, *ln, assign_mode=ops.ASSIGN_ARGMAX)
            mine = acc.packed.clone()
            px.apply(cur, acc.packed, 0.9, out=cur, normalised_out=bank_n)
            # reference: NCCL all-reduce of the same payload (2 ranks: a + b in either order) + apply
            dist.all_reduce(mine)
            assert torch.equal(acc.packed, mine), "peer sum != all-reduce sum (step %d)" % it
            want = ops.proto_ema_apply(want, mine, 0.9)
            assert torch.equal(cur, want), "peer-memory EMA != all-reduce + apply (step %d)" % it
            assert torch.equal(bank_n, ops.bank_normalise(cur))
        assert px.errors() == 0
        banks = [torch.empty_like(cur) for _ in range(world)]
        dist.all_gather(banks, cur)
        assert all(torch.equal(b, banks[0]) for b in banks), "peer-memory banks differ across ranks"
        px.close()
        # the step pipeline, exchange inside the captured graph (peer memory by default)
        step = HotPathStep(synth.NUSCENES, 2, dim=32, sub_protos=4, num_anchor=16, n_sets=2,
                           seed0=500 + 10000 * rank, device=dev)
        for i in range(2):
            step.run(i, seed=i)
        graphed = step.capture()
        for i in range(4):
            step.step(i)
        torch.cuda.synchronize(dev)
        banks = [torch.empty_like(step.protos) for _ in range(world)]
        dist.all_gather(banks, step.protos)
        assert all(torch.equal(b, banks[0]) for b in banks), "pipeline banks differ across ranks"
        losses = [torch.zeros((), device=dev) for _ in range(world)]
        dist.all_gather(losses, step.loss)
        assert not torch.equal(losses[0], losses[1])          # ranks really saw different scans
        assert step.peer is not None and step.peer.errors() == 0
        if rank == 0:
            result["ok"] = True
            result["graphed"] = bool(graphed)
        step.graphs = None
        step.peer.close()
        dist.barrier()
        torch.cuda.synchronize(dev)
    finally:
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def test_two_ranks_nccl_bank_rule():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        result = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, result)) for r in range(2)]
        for p in procs:
            p.start()
        import time
        deadline = time.time() + 150
        while time.time() < deadline and any(p.is_alive() for p in procs):
            if any(p.exitcode not in (None, 0) for p in procs):
                break                       # a rank failed: its peer may be stuck in a collective
            time.sleep(0.5)
        codes = [p.exitcode for p in procs]
        for p in procs:
            if p.is_alive():
                p.terminate()
        assert all(c == 0 for c in codes), codes
        assert result.get("ok")
