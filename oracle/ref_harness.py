"""Import the UNMODIFIED reference sequence half as the oracle-of-the-oracle.

Works only where the reference checkout exists (the build container, /root/reference): used by
tests/golden/make_golden.py to generate committed fixtures and by the `not gpu` tests that pin
oracle/sequence.py against the real thing (skipped when the checkout is absent, e.g. on the GPU box).
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("LR_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models", "lipreader"))


class _RefModules:
    pass


_cached = None


def load():
    """Returns a namespace with the reference modules: data_loader, better_model, train_better_model,
    ctc_loss.  The reference's top-level package is called `src`, as is this repo's alias package, so
    the import is done with a temporarily swapped sys.path / sys.modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_mods:
        del sys.modules[k]
    os.environ.setdefault("LIP_READING_WS_PATH", "/tmp/lr_ws")
    sys.path = [_SHIMS, REFERENCE_ROOT] + [p for p in saved_path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import src.data.data_loader as dl
            import src.models.lipreader.better_model as bm
            import src.train.train_better_model as tb
            import src.train.ctc_loss as cl
        ns = _RefModules()
        ns.data_loader, ns.better_model, ns.train_better_model, ns.ctc_loss = dl, bm, tb, cl
    finally:
        ref_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path = saved_path
    _cached = ns
    return ns
