"""CPU, world_size 2, gloo: the host side of the multi-GPU path -- scan sharding
and the single all-reduce of the packed prototype payload.  The
per-rank payloads come from the oracle (no GPU here); what is checked is that
sharding + `allreduce_packed` + one EMA gives every rank the same bank, equal to
the EMA of the summed payloads."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coarse3d_b200 import distributed
from oracle import proto_ema as oema


def _problem(B, D, H, W, C, M, seed):
    g = torch.Generator().manual_seed(seed)
    centers = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=-1)
    label = torch.randint(1, C, (B, H, W), generator=g)
    emb = torch.nn.functional.normalize(
        torch.randn(B, D, H, W, generator=g) * 0.7 + 1.5 * centers[label].permute(0, 3, 1, 2), dim=1)
    protos0 = torch.randn(C, M, D, generator=g) * 0.02 + centers[:, None, :] * 0.05
    label = label * (torch.rand(B, H, W, generator=g) < 0.1)
    ln = [torch.ones(D), torch.zeros(D), torch.ones(C), torch.zeros(C)]
    return emb, label, protos0, ln


def _rank_payload(emb, label, protos0, ln, C, idx):
    _, sums, counts, _ = oema.prototype_learning(emb[idx], label[idx], protos0, *ln, C, 0, 0.9,
                                                 labelled_only=True)
    return torch.cat([sums.reshape(-1), counts.reshape(-1)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, D, H, W, C, M = 5, 16, 4, 32, 6, 3
        emb, label, protos0, ln = _problem(B, D, H, W, C, M, 1)
        idx = list(distributed.shard_scans(B, rank, world))
        packed = _rank_payload(emb, label, protos0, ln, C, idx)
        assert distributed.world() == (rank, world)
        distributed.allreduce_packed(packed)
        K = C * M
        new = oema.ema_from_sums(oema.l2_normalize(protos0), packed[:K * D].view(C, M, D),
                                 packed[K * D:].view(C, M), 0.9)
        gathered = [torch.empty_like(new) for _ in range(world)]
        dist.all_gather(gathered, new)
        if rank == 0:
            torch.save({"new": new, "same": all(torch.equal(g, new) for g in gathered),
                        "packed": packed}, out)
    finally:
        dist.destroy_process_group()


def test_shard_scans_partitions_evenly():
    for n, w in [(8, 2), (64, 8), (5, 2), (3, 4), (0, 2)]:
        parts = [list(distributed.shard_scans(n, r, w)) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert distributed.world() == (0, 1)
    t = torch.ones(4)
    assert distributed.allreduce_packed(t) is None and torch.equal(t, torch.ones(4))


@pytest.mark.timeout(120)
def test_two_rank_allreduce_then_single_ema(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["same"], "ranks ended with different banks"
    B, D, H, W, C, M = 5, 16, 4, 32, 6, 3
    emb, label, protos0, ln = _problem(B, D, H, W, C, M, 1)
    total = sum(_rank_payload(emb, label, protos0, ln, C, list(distributed.shard_scans(B, r, 2)))
                for r in range(2))
    assert torch.allclose(res["packed"], total, rtol=0, atol=1e-6)
    K = C * M
    want = oema.ema_from_sums(oema.l2_normalize(protos0), total[:K * D].view(C, M, D),
                              total[K * D:].view(C, M), 0.9)
    assert torch.allclose(res["new"], want, atol=1e-6)
    assert total[K * D:].sum() > 0
