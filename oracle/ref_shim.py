"""Import shim for the UNMODIFIED reference modules (test / baseline infrastructure, never imported by the product).

Resolution order: oracle/_ref/ (materialised by oracle/build_ref.py; this is what exists on the GPU box), then the
read-only checkout (/root/reference or $NEURADAR_REFERENCE; container only).  The reference imports ~25 packages that are
absent here and are not on the arithmetic path (viewer, plotting, metrics, dataset devkits, nerfacc); they are replaced
by permissive stub modules so that its fp32 torch path can be executed.
"""
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, "_ref"), os.environ.get("NEURADAR_REFERENCE", "/root/reference")]

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

    def __getitem__(self, item):  # used as a generic in annotations
        return self


def reference_root():
    for root in _CANDIDATES:
        if root and os.path.isdir(os.path.join(root, "nerfstudio")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


def install() -> str:
    """Make `import nerfstudio...` resolve to the reference modules; returns the root that was used."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not found: run `python oracle/build_ref.py` where /root/reference exists")
    for name in _STUBS:
        if name in sys.modules:
            continue
        mod = _Any(name)
        mod.__path__ = []
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = mod
    import typing

    _install_nerfacc_dense(sys.modules["nerfacc"])
    sys.modules["git"].Optional = typing.Optional  # model_components/radar_utils.py:20 imports typing.Optional through GitPython
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def _install_nerfacc_dense(mod) -> None:
    """Give the `nerfacc` placeholder the three entry points the model calls (nerfacc==0.5.2 is a third-party package that
    is not vendored).  CPU tensors: the dense contract restated in torch - weights = alpha *
    exclusive_cumprod(1 - alpha), accumulate = sum_s w v - which is what makes the reference model a CPU oracle for the
    compositing tail; CUDA tensors: the product's drop-in module (neuradar_b200.nerfacc_compat), i.e. what a box without
    nerfacc would run."""
    import torch

    def _gpu():
        from neuradar_b200 import nerfacc_compat

        return nerfacc_compat

    def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
        if alphas.is_cuda:
            return _gpu().render_weight_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
        trans = torch.cumprod(torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas[..., :-1]], dim=-1), dim=-1)
        return alphas * trans, trans

    def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
        if sigmas.is_cuda:
            return _gpu().render_weight_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays, prefix_trans)
        ds = sigmas * (t_ends - t_starts)
        alphas = 1 - torch.exp(-ds)
        trans = torch.exp(-torch.cat([torch.zeros_like(ds[..., :1]), torch.cumsum(ds[..., :-1], dim=-1)], dim=-1))
        return alphas * trans, trans, alphas

    def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
        if weights.is_cuda:
            return _gpu().accumulate_along_rays(weights, values, ray_indices, n_rays)
        if values is None:
            return weights.sum(dim=-1, keepdim=True)
        return (weights[..., None] * values).sum(dim=-2)

    mod.render_weight_from_alpha = render_weight_from_alpha
    mod.render_weight_from_density = render_weight_from_density
    mod.accumulate_along_rays = accumulate_along_rays
