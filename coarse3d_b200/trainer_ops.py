"""Reference-shaped replacements for the hot-path methods of
tasks/weak_segmentation/trainer.py `Trainer` (the loss's caller)."""
from . import ops


def entropy_based_selection(self, output, wss_mask, eval_mask, train_label, select_ratio):
    """Drop-in for `Trainer.entropy_based_selection` (trainer.py:447-518): same arguments,
    same `(pseudo_label int64 (B,H,W), new_wss_mask bool (B,H,W))` result, one batched call
    instead of the B x C Python loop.  Reads `self.settings.ignore_cls` / `n_classes` like
    the reference.  Draws come from the device Philox stream seeded by torch's global
    generator (the reference consumes the same generator through torch.multinomial)."""
    if output.shape[1] != self.settings.n_classes:
        raise ValueError("output has %d channels, settings.n_classes is %d"
                         % (output.shape[1], self.settings.n_classes))
    return ops.entropy_select_batch(output.contiguous(), wss_mask.bool().contiguous(),
                                    eval_mask.bool().contiguous(), train_label.long().contiguous(),
                                    float(select_ratio), ignore_cls=int(self.settings.ignore_cls))
