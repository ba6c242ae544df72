"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference package from /root/reference (this container
only; the path does not exist on the GPU box) so that `oracle/gen_golden.py`
can execute the reference's own hot-path code and record golden vectors.

The reference's package __init__ pulls in gymnasium / matplotlib / lz4 / wandb
(active_inference_diffusion/__init__.py:13-17, agents/base_agent.py:12,
utils/util.py:2, utils/buffers.py:9, utils/logger.py:5), none of which are on
the hot path and none of which are installed here.  We pre-seed `sys.modules`
with inert stand-ins for exactly those names; no reference source is modified
or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AID_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "active_inference_diffusion"))


class _Anything:
    """Base class stand-in: subclassable, callable, attribute-tolerant."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub_module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _module_getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything

    mod.__getattr__ = _module_getattr  # type: ignore[attr-defined]
    mod.__path__ = []  # behave like a package so `import x.y` resolves
    sys.modules[name] = mod
    return mod


def _install_stubs() -> None:
    def have(name: str) -> bool:
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not have("gymnasium"):
        spaces = _stub_module("gymnasium.spaces", Box=_Anything, Dict=_Anything, Space=_Anything)
        wrappers = _stub_module("gymnasium.wrappers", RecordVideo=_Anything,
                                FrameStackObservation=_Anything)
        _stub_module("gymnasium", Env=_Anything, Wrapper=_Anything,
                     ObservationWrapper=_Anything, ActionWrapper=_Anything,
                     RewardWrapper=_Anything, Space=_Anything, spaces=spaces,
                     wrappers=wrappers, make=_Anything())
    if not have("matplotlib"):
        pyplot = _stub_module("matplotlib.pyplot")
        _stub_module("matplotlib", pyplot=pyplot, use=lambda *a, **k: None)
    if not have("lz4"):
        frame = _stub_module("lz4.frame", compress=lambda b: b, decompress=lambda b: b)
        _stub_module("lz4", frame=frame)
    if not have("wandb"):
        _stub_module("wandb", init=lambda *a, **k: None, log=lambda *a, **k: None,
                     finish=lambda *a, **k: None)
    if not have("cloudpickle"):
        import pickle
        _stub_module("cloudpickle", dumps=pickle.dumps, loads=pickle.loads)
    if not have("mujoco"):
        _stub_module("mujoco")
    if not have("PIL"):
        image = _stub_module("PIL.Image")
        _stub_module("PIL", Image=image)
    if not have("cv2"):
        _stub_module("cv2")
    if not have("imageio"):
        _stub_module("imageio")


_ref = None


def import_reference():
    """Return the reference's top-level package (`active_inference_diffusion`)."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import torch  # noqa: F401  (must be fully imported before the stubs exist)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _ref = importlib.import_module("active_inference_diffusion")
    return _ref
