"""Modules with the reference's names and call signatures for the hot path.

    import mdseg_b200.dropin as dropin
    dropin.install()            # before the trainer imports lib.loss.* / lib.class_remap

`install()` registers these modules in ``sys.modules`` under the reference's import paths, so
``from lib.loss.ohem_ce_loss import OhemCELoss`` (tools/train_amp.py:24, lib/loss/loss_cross_datasets.py:12)
resolves to the B200 implementation while every other ``lib.*`` module still comes from the
reference checkout.  Nothing here computes on the CPU: tensors must be CUDA tensors and the C-ABI
library must be present (ImportError otherwise).
"""
import importlib
import sys

# reference import path -> module of this package
MODULES = {
    "lib.loss.ohem_ce_loss": "ohem_ce_loss",
    "lib.class_remap": "class_remap",
}


def install(extra=()):
    """Alias the drop-in modules under the reference's import paths.  `extra` may add
    ("lib.loss.loss_cross_datasets", "loss_cross_datasets") to route the SEG-stage loss as well."""
    done = {}
    for ref_name, local in list(MODULES.items()) + list(extra):
        mod = importlib.import_module(f"{__name__}.{local}")
        sys.modules[ref_name] = mod
        done[ref_name] = mod
    return done
