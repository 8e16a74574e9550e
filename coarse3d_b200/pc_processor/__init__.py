"""Mirror of the reference's `pc_processor` operator surface for the hot path.

Only the four hot-path operators live here, at the same dotted paths and with
the same names, signatures and error behaviour as the reference:

    pc_processor.dataset.preprocess.projection.RangeProjection
    pc_processor.postproc.knn.KNN
    pc_processor.loss.contrast_pixel_loss.ContrastMEMLoss
    pc_processor.models.prototype.{momentum_update, prototype_learning, PrototypeBank}
        (`prototype_learning` is the method `install()` binds onto SalsaNextProto /
        RangeNetProto / SqueezeSegV3Proto; `PrototypeBank` carries it too)

`coarse3d_b200.install()` swaps them into the real `pc_processor` package so
that tasks/weak_segmentation runs unchanged.
"""
from . import dataset, loss, models, postproc  # noqa: F401
