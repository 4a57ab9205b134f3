"""mdseg_b200 — B200-native (sm_100a) per-pixel multi-dataset label-space hot path.

Layout
    csrc/      hand-written CUDA kernels + the C ABI (include/mdseg.h) -> libmdseg_b200.so
    native.py  ctypes binding of the C ABI (no torch)
    ops.py     torch-facing operators / autograd Functions over the C ABI
    dropin/    modules with the reference's import paths and call signatures
               (lib.loss.ohem_ce_loss, lib.class_remap, lib.loss.loss_cross_datasets, evaluate)

The directory name carries a hyphen (it is the repo's name); import it as
``mdseg_b200`` (alias package at the repo root).
"""
__version__ = "0.1.0"
