"""Pass-through to the reference checkout for the names a drop-in module does not implement itself.

The drop-in modules are installed under the reference's import paths (``lib.loss.ohem_ce_loss`` ...), so they must
be SUPERSETS of the modules they shadow: ``lib/loss/loss_cross_datasets.py:6`` imports three classes from
``lib.loss.ohem_ce_loss`` and the ``ltbgnn_*`` trainers import six loss classes from ``lib.loss.loss_cross_datasets``
(``tools/train_ltbgnn_all_datasets_snp.py:28``, ``tools/eval_snp.py:28``).  Names on the hot path are implemented on
libmdseg_b200.so; every other name is served from the reference's OWN source file, loaded under a private module name
(``_mdseg_reference.<path>``) so that it cannot shadow the drop-in.  The reference file's own ``from lib...`` imports
resolve through ``sys.modules`` and therefore pick up the drop-in classes.

The reference checkout is found on ``sys.path`` (the trainers run from its root with ``sys.path.insert(0, '.')``) or
under ``$MDSEG_REFERENCE_ROOT``.  Without a checkout the drop-in modules still work; only the pass-through names are
missing (AttributeError naming this mechanism).
"""
import importlib.util
import os
import sys
import types

PRIVATE_PREFIX = "_mdseg_reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = {}


def find_source(ref_name):
    """Path of the reference's source file of module `ref_name` ('lib.class_remap'), or None."""
    rel = ref_name.replace(".", os.sep) + ".py"
    roots = []
    env = os.environ.get("MDSEG_REFERENCE_ROOT")
    if env:
        roots.append(env)
    roots += [p if p else os.getcwd() for p in sys.path]
    for root in roots:
        cand = os.path.abspath(os.path.join(root, rel))
        if os.path.isfile(cand) and not cand.startswith(_HERE + os.sep):
            return cand
    return None


def load(ref_name):
    """The reference's own module `ref_name`, executed from its source file under a private name."""
    if ref_name in _loaded:
        return _loaded[ref_name]
    src = find_source(ref_name)
    if src is None:
        raise ImportError(f"no reference checkout with {ref_name.replace('.', '/')}.py on sys.path or under "
                          "$MDSEG_REFERENCE_ROOT")
    name = f"{PRIVATE_PREFIX}.{ref_name}"
    spec = importlib.util.spec_from_file_location(name, src)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    _loaded[ref_name] = mod
    return mod


def try_load(ref_name):
    try:
        return load(ref_name)
    except Exception:  # missing checkout, or the reference file's own imports are not installed here
        return None


def module_getattr(ref_name, native_names):
    """A module-level ``__getattr__`` (PEP 562) serving every name the drop-in lacks from the reference's file."""

    def __getattr__(name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        try:
            mod = load(ref_name)
        except ImportError as e:
            raise AttributeError(
                f"{ref_name}.{name}: not implemented natively by the B200 drop-in (native: {sorted(native_names)}) "
                f"and the reference pass-through is unavailable: {e}") from e
        try:
            return getattr(mod, name)
        except AttributeError:
            raise AttributeError(f"module {ref_name!r} has no attribute {name!r} (neither the B200 drop-in nor the "
                                 "reference's own file defines it)") from None

    return __getattr__


def graft_method(obj, ref_name, cls_name, name):
    """Bound method `name` of the reference's class `cls_name`, running on the drop-in instance `obj` (which keeps
    every attribute the reference's ``__init__`` sets).  AttributeError when the reference has no such method."""
    mod = try_load(ref_name)
    fn = getattr(getattr(mod, cls_name, None), name, None) if mod is not None else None
    if fn is None or not callable(fn):
        raise AttributeError(f"{type(obj).__name__!r} object has no attribute {name!r} (B200 drop-in; not found in "
                             f"the reference's {ref_name}.{cls_name} either)")
    return types.MethodType(fn, obj)
