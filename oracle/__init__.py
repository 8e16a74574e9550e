"""CPU oracle for the COARSE3D per-scan hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain numpy / torch-CPU,
the algorithms of the four reference operators that `coarse3d_b200` replaces
with hand-written CUDA:

    oracle.projection   <- pc_processor/dataset/preprocess/projection.py:4-115
    oracle.knn          <- pc_processor/postproc/knn.py:11-142
    oracle.proto_loss   <- pc_processor/loss/contrast_pixel_loss.py:8-195
    oracle.proto_ema    <- pc_processor/models/salsanext_proto.py:19-35,337-402,497-510
                           pc_processor/models/sinkhorn.py:5-33
    oracle.assemble     <- pc_processor/dataset/semantic_kitti/wss_sem_kitti_loader.py:124-172
                           tasks/weak_segmentation/trainer.py:600-608   (the projection's caller)
    oracle.unproject    <- tasks/weak_segmentation/trainer.py:714-724,
                           pc_processor/metrics/iou_eval.py:35-58
    oracle.entropy_select <- tasks/weak_segmentation/trainer.py:447-518
    oracle.lovasz       <- pc_processor/loss/lovasz_softmax.py:51-157

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU
baseline.  Nothing under `coarse3d_b200/` imports it; the product path has no
CPU fallback.

Pinning.  The reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, executed in the build container by
`tests/golden/make_golden.py` (which imports them from /root/reference) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.
Where the reference leaves behaviour undefined (unstable argsort on equal
depths, `torch.topk` tie choice, the RNG streams of `multinomial`,
`randperm`, `gumbel_softmax`, and numpy's machine-dependent SIMD float32
`arctan2`/`arcsin`) the oracle fixes a rule, stated in each module's header.
"""
