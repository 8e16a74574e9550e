"""Oracle: un-projection gather + confusion matrix (TEST INFRASTRUCTURE).

Restates tasks/weak_segmentation/trainer.py:714-724 (`argmax_2d[ii, uproj_y_idx[ii],
uproj_x_idx[ii]]` per scan) and `IOUEval.addBatch`, pc_processor/metrics/iou_eval.py:35-58
(conf[pred, gt] += 1 via index_put with accumulate; rows = prediction, columns = target).
CSR batch instead of the loader's zero-padded (max_points) arrays.
"""
import torch


def unproject_confusion(argmax_2d, px, py, offsets, labels, nclasses):
    """argmax_2d (B,H,W) int; px, py, labels (sum N,) int; offsets (B+1,).
    Returns (unproj_argmax (sum N,) int64, conf_matrix (C,C) int64)."""
    conf = torch.zeros((nclasses, nclasses)).long()                    # iou_eval.py:29-31
    out = []
    for ii in range(argmax_2d.shape[0]):
        lo, hi = int(offsets[ii]), int(offsets[ii + 1])
        u = argmax_2d[ii, py[lo:hi].long(), px[lo:hi].long()]           # trainer.py:719
        out.append(u.long())
        if labels is not None:
            x_row, y_row = u.reshape(-1).long(), labels[lo:hi].reshape(-1).long()  # :44-45
            idxs = torch.stack([x_row, y_row], dim=0)                   # :48
            conf = conf.index_put_(tuple(idxs), torch.ones(idxs.shape[-1]).long(), accumulate=True)
    return torch.cat(out), conf
