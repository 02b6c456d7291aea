"""neuradar_b200: the NeuRadar per-ray neural-field hot path as hand-written sm_100a CUDA kernels behind the
nerfstudio / neurad-studio API (Field.get_density / forward, RaySamples, ProposalNetworkSampler / PDFSampler,
compositing and renderers).  See DESIGN.md and include/neuradar_b200.h."""
from . import _lib  # noqa: F401
from .field_components import (  # noqa: F401
    MLP,
    ActorSettings,
    HashEncoding,
    NeuRADHashEncoding,
    NeuRADHashEncodingConfig,
    SHEncoding,
    SigmoidDensity,
    StaticSettings,
    trunc_exp,
)
from .fields import (  # noqa: F401
    FieldHeadNames,
    NeuRADField,
    NeuRADFieldConfig,
    NeuRADProposalField,
    NeuRADProposalFieldConfig,
)
from .losses import distortion_loss, lossfun_distortion, zipnerf_interlevel_loss  # noqa: F401
from .nff import LossSettings, NeuRadarHotPath, NeuRadarHotPathConfig, SamplingSettings, bench_loss, training_losses  # noqa: F401
from .ray_samplers import PDFSampler, PowerSampler, ProposalNetworkSampler  # noqa: F401
from .radars import generate_rays_from_fov  # noqa: F401
from .rays import Frustums, GaussiansStd, RayBundle, RaySamples  # noqa: F401
from .renderers import AccumulationRenderer, DepthRenderer, FeatureRenderer, render_depth_simple  # noqa: F401

__version__ = "0.1.0"
