"""Import shim for the read-only reference checkout (container only, never on the GPU box).

The reference (nerfstudio fork) imports ~25 packages that are absent here and are not on the
arithmetic path (viewer, plotting, metrics, dataset devkits).  They are replaced by permissive
stub modules so that the reference's own fp32 torch path can be executed to produce golden vectors.
Used only by tests/golden/make_golden.py and tests that are skipped when /root/reference is missing.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NEURADAR_REFERENCE", "/root/reference")

_STUBS = (
    "viser viser.transforms nerfacc matplotlib matplotlib.pyplot matplotlib.cm plotly plotly.graph_objects "
    "plotly.express torchmetrics torchmetrics.functional torchmetrics.image torchmetrics.image.lpip pyquaternion "
    "open3d mediapy splines splines.quaternion gsplat timm pytorch_msssim zod comet_ml av vod pathos git "
    "sklearn sklearn.neighbors"
).split()


class _Any(types.ModuleType):
    """A module whose every attribute is again a stub that can be called or used as a base class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Any(f"{self.__name__}.{name}")
        child.__spec__ = importlib.machinery.ModuleSpec(child.__name__, None)
        setattr(self, name, child)
        sys.modules[child.__name__] = child
        return child

    def __call__(self, *args, **kwargs):
        return self

    def __mro_entries__(self, bases):
        return (object,)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nerfstudio"))


def install() -> None:
    """Make `import nerfstudio...` resolve to the reference checkout."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name in sys.modules:
            continue
        mod = _Any(name)
        mod.__path__ = []
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
