"""Modules with the reference's names and call signatures for the hot path.

    import mdseg_b200.dropin as dropin
    dropin.install()            # before the trainer imports lib.loss.* / lib.class_remap

`install()` registers these modules in ``sys.modules`` under the reference's import paths, so
``from lib.loss.ohem_ce_loss import OhemCELoss`` (tools/train_amp.py:24, lib/loss/loss_cross_datasets.py:6) and
``from lib.loss.loss_cross_datasets import CrossDatasetsLoss, ..., CrossDatasetsCELoss_AdvGNN``
(tools/train_ltbgnn_all_datasets_snp.py:28) resolve to the B200 implementation while every other ``lib.*`` module
still comes from the reference checkout.  Each drop-in module is a SUPERSET of the module it shadows: the classes on
the hot path are native (libmdseg_b200.so), every other name is passed through from the reference's own file
(``_reference.py``).  Nothing here computes on the CPU: tensors must be CUDA tensors and the C-ABI library must be
present (ImportError otherwise).

native:       lib.loss.ohem_ce_loss        OhemCELoss, MdsOhemCELoss, MdsOhemNLLPlusLoss
              lib.class_remap              ClassRemap (every method), ClassRemapOneHotLabel.{SegRemapping,
                                           SingleSegRemappingOneHot}
              lib.loss.loss_cross_datasets CrossDatasetsCELoss, CrossDatasetsCELoss_CLIP, CrossDatasetsCELoss_GNN,
                                           CrossDatasetsCELoss_AdvGNN, LabelToOneHot, UnifyPrototypeFunction
passed through (reference code, using the native classes above where it calls them):
              CrossDatasetsLoss, CrossDatasetsCELoss_KMeans, CrossDatasetsCELoss_AdvGNN_ce, the contrast / prototype
              remaps of ClassRemapOneHotLabel, RecallCrossEntropy / FocalLoss / AdjNLLPlusLoss re-exports, ...
"""
import importlib
import sys
import types

# reference import path -> module of this package
MODULES = {
    "lib.loss.ohem_ce_loss": "ohem_ce_loss",
    "lib.class_remap": "class_remap",
    "lib.loss.loss_cross_datasets": "loss_cross_datasets",
}


def install(extra=(), only=None):
    """Alias the drop-in modules under the reference's import paths.  `only`: restrict to some reference paths;
    `extra`: additional (reference path, local module) pairs."""
    done = {}
    for ref_name, local in list(MODULES.items()) + list(extra):
        if only is not None and ref_name not in only:
            continue
        mod = importlib.import_module(f"{__name__}.{local}")
        sys.modules[ref_name] = mod
        done[ref_name] = mod
        # `import lib.loss.x as m` walks the parent packages: use the reference's when a checkout is importable,
        # empty stand-ins otherwise, and hang the drop-in on its parent like the import system would
        parts = ref_name.split(".")
        for k in range(1, len(parts)):
            parent = ".".join(parts[:k])
            if parent not in sys.modules:
                try:
                    importlib.import_module(parent)
                except ImportError:
                    stub = types.ModuleType(parent)
                    stub.__path__ = []
                    stub.__mdseg_stub__ = True
                    sys.modules[parent] = stub
                    if k > 1:
                        setattr(sys.modules[".".join(parts[:k - 1])], parts[k - 1], stub)
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return done


def uninstall():
    """Remove the aliases (and the privately loaded reference modules) again — for tests."""
    from . import _reference
    for ref_name in MODULES:
        mod = sys.modules.get(ref_name)
        if mod is not None and getattr(mod, "__name__", "").startswith(__name__):
            del sys.modules[ref_name]
    for name in [n for n in sys.modules if n.startswith(_reference.PRIVATE_PREFIX + ".")
                 or getattr(sys.modules[n], "__mdseg_stub__", False)]:
        del sys.modules[name]
    _reference._loaded.clear()
