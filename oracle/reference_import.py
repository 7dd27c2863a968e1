"""Imports the UNMODIFIED reference modules from /root/reference (this container only).

Used by tests/golden/make_golden.py to generate fixtures and by tests that validate the oracle
restatement against the live reference when it is present.  The GPU box has no /root/reference:
nothing that runs there may call this module.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("NERF_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "barf"))


def load(variant: str = "barf"):
    """Returns a namespace with the reference modules of one variant directory imported flat,
    the way the reference itself imports them (CWD = variant directory)."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    vdir = os.path.join(REFERENCE_ROOT, variant)
    # the variant directories share module names: drop earlier flat imports first
    for name in list(sys.modules):
        mod = sys.modules[name]
        f = getattr(mod, "__file__", None) or ""
        if f.startswith(REFERENCE_ROOT):
            del sys.modules[name]
    for p in (vdir, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, _STUBS)
    sys.path.insert(0, vdir)

    class NS:
        pass

    ns = NS()
    names = {
        "barf": ["positional_encodings", "model_interpolation_architecture", "gaussian",
                 "model_camera_extrinsics", "magic", "model_interpolation",
                 "model_garf_radiance", "model_garf_proposal"],
        "garf": ["gaussian", "model_radiance", "model_proposal", "model_camera_extrinsics"],
        "sarf": ["activation"],
        "gaborf": ["gabor"],
    }[variant]
    for n in names:
        setattr(ns, n, importlib.import_module(n))
    return ns
