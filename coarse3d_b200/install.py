"""Swap the hot-path operators into the reference's `pc_processor` package."""
import importlib
import sys


def install(pc_processor=None):
    """Replace the reference's hot-path operators with the B200 ones.

    Call once, after `import pc_processor` (the reference package) and before
    the task script builds its datasets / criteria, e.g. at the top of
    tasks/weak_segmentation/main.py:

        import pc_processor, coarse3d_b200
        coarse3d_b200.install(pc_processor)

    Everything else in `pc_processor` is left untouched.
    """
    if pc_processor is None:
        pc_processor = sys.modules.get("pc_processor") or importlib.import_module("pc_processor")
    from .pc_processor.dataset.preprocess.projection import RangeProjection
    from .pc_processor.postproc.knn import KNN
    from .pc_processor.loss.contrast_pixel_loss import ContrastMEMLoss

    pc_processor.dataset.preprocess.projection.RangeProjection = RangeProjection
    pc_processor.dataset.preprocess.RangeProjection = RangeProjection
    pc_processor.postproc.knn.KNN = KNN
    pc_processor.postproc.KNN = KNN
    pc_processor.loss.contrast_pixel_loss.ContrastMEMLoss = ContrastMEMLoss
    pc_processor.loss.ContrastMEMLoss = ContrastMEMLoss
    return pc_processor
