"""Oracle: feature assembly on top of the projection (TEST INFRASTRUCTURE).

Restates the caller side of `RangeProjection` in the reference's data path:
  * pc_processor/dataset/semantic_kitti/wss_sem_kitti_loader.py:124-132 -- projected
    eval / train label images from the per-point labels of the winning points;
  * wss_sem_kitti_loader.py:159-172 -- the (5,H,W) input `[range, x, y, z,
    intensity * (intensity != -1)]`;
  * tasks/weak_segmentation/trainer.py:600-608 -- `.long()` labels, `eval_mask =
    eval_label > 0`, `(feature - mean) / std * eval_mask`.
(The nuScenes loader builds the same tensors, wss_nuscenes_loader.py:126-171.)
All arithmetic is float32, one numpy/torch op per reference op.
"""
import numpy as np

from . import projection as oproj

F32 = np.float32


def assemble(points, fov, sem_label, weak_label, img_mean=None, img_std=None, depth=None):
    """One scan.  Returns dict: feature (5,H,W) f32, train_label / eval_label (H,W) i64,
    plus the projection outputs the trainer also consumes (proj_range, proj_idx,
    uproj_x_idx, uproj_y_idx, uproj_depth)."""
    o = oproj.project(points, fov, depth)
    idx = o["proj_idx"]
    valid = idx > -1
    eval_label = np.zeros(idx.shape, dtype=F32)                       # loader :124-127
    eval_label[valid] = np.asarray(sem_label)[idx[valid]]
    train_label = np.zeros(idx.shape, dtype=F32)                      # loader :129-132
    train_label[valid] = np.asarray(weak_label)[idx[valid]]
    pc = o["proj_pointcloud"]
    inten = pc[..., 3]
    inten = (inten != -1).astype(F32) * inten                         # loader :161-164
    feature = np.concatenate([o["proj_range"][None], pc[..., :3].transpose(2, 0, 1), inten[None]], 0)
    eval_l = eval_label.astype(np.int64)                              # trainer :600-601
    train_l = train_label.astype(np.int64)
    if img_mean is not None:
        mean = np.asarray(img_mean, dtype=F32)[:, None, None]
        std = np.asarray(img_std, dtype=F32)[:, None, None]
        mask = (eval_l > 0)                                           # trainer :603
        feature = ((feature - mean) / std * mask[None].astype(F32)).astype(F32)  # :604-608
    return dict(feature=feature.astype(F32), train_label=train_l, eval_label=eval_l,
                proj_range=o["proj_range"], proj_idx=idx, uproj_x_idx=o["uproj_x_idx"],
                uproj_y_idx=o["uproj_y_idx"], uproj_depth=o["uproj_depth"])
