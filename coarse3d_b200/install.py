"""Swap the hot-path operators into the reference's `pc_processor` package."""
import importlib
import sys


def install(pc_processor=None, trainer_cls=None):
    """Replace the reference's hot-path operators with the B200 ones.

    Call once, after `import pc_processor` (the reference package) and before
    the task script builds its datasets / criteria, e.g. at the top of
    tasks/weak_segmentation/main.py:

        import pc_processor, coarse3d_b200
        coarse3d_b200.install(pc_processor)

    Rebound: `RangeProjection`, `KNN`, `ContrastMEMLoss`, `Lovasz_softmax`, and the
    `prototype_learning` method + `momentum_update` helper of SalsaNextProto / RangeNetProto /
    SqueezeSegV3Proto.  Everything else in `pc_processor` is left untouched.  Pass the task's `Trainer` class
    (tasks/weak_segmentation/trainer.py:17) as `trainer_cls` to also replace its
    `entropy_based_selection` method (trainer.py:447-518) with the batched kernel.
    """
    if pc_processor is None:
        pc_processor = sys.modules.get("pc_processor") or importlib.import_module("pc_processor")
    from .pc_processor.dataset.preprocess.projection import RangeProjection
    from .pc_processor.postproc.knn import KNN
    from .pc_processor.loss.contrast_pixel_loss import ContrastMEMLoss
    from .pc_processor.loss.lovasz_softmax import Lovasz_softmax, lovasz_softmax

    pc_processor.dataset.preprocess.projection.RangeProjection = RangeProjection
    pc_processor.dataset.preprocess.RangeProjection = RangeProjection
    pc_processor.postproc.knn.KNN = KNN
    pc_processor.postproc.KNN = KNN
    pc_processor.loss.contrast_pixel_loss.ContrastMEMLoss = ContrastMEMLoss
    pc_processor.loss.ContrastMEMLoss = ContrastMEMLoss
    if hasattr(pc_processor.loss, "lovasz_softmax"):   # pc_processor/loss/__init__.py:2
        pc_processor.loss.lovasz_softmax.Lovasz_softmax = Lovasz_softmax
        pc_processor.loss.lovasz_softmax.lovasz_softmax = lovasz_softmax
    pc_processor.loss.Lovasz_softmax = Lovasz_softmax
    # a3: the EMA prototype update is a METHOD of the three *Proto models and a module-level
    # helper of their files (salsanext_proto.py:19-31,337-402; rangenet_proto.py:460-567;
    # squeezesegv3_Proto.py:253-351): rebind both, on every model file that is present.
    from .pc_processor.models.prototype import momentum_update, prototype_learning
    models = getattr(pc_processor, "models", None)
    for mod_name, cls_name in (("salsanext_proto", "SalsaNextProto"), ("rangenet_proto", "RangeNetProto"),
                               ("squeezesegv3_Proto", "SqueezeSegV3Proto")):
        mod = getattr(models, mod_name, None) if models is not None else None
        cls = getattr(mod, cls_name, None) if mod is not None else None
        if cls is None:
            continue
        cls.prototype_learning = prototype_learning
        mod.momentum_update = momentum_update
    if trainer_cls is not None:
        from .trainer_ops import entropy_based_selection
        trainer_cls.entropy_based_selection = entropy_based_selection
    return pc_processor
