"""The drop-in boundary end to end: the reference's OWN `NeuRadarModel` (unmodified, from oracle/_ref) is
constructed, its hot-path submodules are swapped by `neuradar_b200.plugin.convert_neuradar_model`, and the reference's
`get_nff_outputs` code then drives the B200 kernels with the reference's `RayBundle` - also under `torch.autocast`, as
`engine/trainer.py:564` runs it.  Compared with the untouched reference model on the CPU (its torch path)."""
import pytest
import torch

from oracle import ref_shim
from tests.parity_utils import FixedJitter, rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available(), reason="reference modules not materialised (oracle/build_ref.py)")]
DEV = "cuda"


def _reference_model(log2=12):
    ref_shim.install()
    import nerfstudio.model_components.losses as L

    class _NoVGG(torch.nn.Module):  # the real one downloads VGG19 weights; vgg loss is not on the hot path
        def forward(self, *a, **k):
            return torch.zeros(())

    L.VGGPerceptualLossPix2Pix = _NoVGG
    import nerfstudio.models.neuradar as M
    from nerfstudio.data.scene_box import SceneBox

    M.VGGPerceptualLossPix2Pix = _NoVGG
    cfg = M.NeuRadarModelConfig()
    cfg.implementation = "torch"
    cfg.field.grid.static.log2_hashmap_size = log2
    cfg.sampling.proposal_field_1.grid.static.log2_hashmap_size = log2
    cfg.sampling.proposal_field_2.grid.static.log2_hashmap_size = log2
    torch.manual_seed(3)
    sb = SceneBox(aabb=torch.tensor([[-100.0, -100.0, -100.0], [100.0, 100.0, 100.0]]))
    meta = {"duration": 20.0, "sensor_idx_to_name": {0: "cam", 1: "lidar", 2: "radar"}, "trajectories": []}
    model = M.NeuRadarModel(cfg, scene_box=sb, num_train_data=1, metadata=meta)
    with torch.no_grad():  # tables away from ~0 so that weights / features are not degenerate
        model.field.hashgrid.static_grid.hash_table.mul_(300.0)
        for p in model.proposal_fields:
            p.hashgrid.static_grid.hash_table.mul_(2000.0)
    return model


def _bundle(n, device):
    from nerfstudio.cameras.rays import RayBundle

    from neuradar_b200.synthetic import synthetic_rays

    r = synthetic_rays(n, seed=11)
    g = torch.Generator().manual_seed(5)
    md = {
        "is_lidar": r["is_lidar"], "is_radar": r["is_radar"],
        "directions_norm": torch.rand((n, 1), generator=g) * 60 + 1,
        "did_return": torch.rand((n, 1), generator=g) > 0.2,
        "sensor_idxs": (r["is_lidar"].long() + 2 * r["is_radar"].long()),
    }
    return RayBundle(origins=r["origins"].to(device), directions=r["directions"].to(device), pixel_area=r["pixel_area"].to(device),
                     nears=r["nears"].to(device), fars=r["fars"].to(device), times=r["times"].to(device),
                     metadata={k: v.to(device) for k, v in md.items()})


def _cpu_render_weights(self, outputs, ray_samples):
    """The dense nerfacc contract of `_render_weights` on CUDA (models/neuradar.py:1016): alpha * exclusive cumprod(1 - alpha).
    The model's own CPU branch returns a constant 0.5 (:1012-1014) and is no oracle."""
    from nerfstudio.field_components.field_heads import FieldHeadNames

    a = outputs[FieldHeadNames.ALPHA].squeeze(-1)
    T = torch.cumprod(torch.cat([torch.ones_like(a[..., :1]), 1 - a[..., :-1]], dim=-1), dim=-1)
    return a * T


def _loss(out):
    loss = out["features"].float().pow(2).mean() + 1e-3 * out["depth"].float().mean()
    for w in out["weights_list"][:-1]:
        loss = loss + w.float().pow(2).mean()
    for k in ("prop_weights_loss_0", "prop_weights_loss_1"):
        loss = loss + 1e-4 * out[k].float()
    return loss + 1e-4 * out["non_nearby_weights"].float().pow(2).sum()


@pytest.mark.parametrize("autocast", [False, True])
def test_reference_model_drives_the_b200_hot_path(autocast):
    import types

    from neuradar_b200 import plugin

    n = 768
    cpu_model = _reference_model()
    cpu_model.train()
    cpu_model._render_weights = types.MethodType(_cpu_render_weights, cpu_model)
    gpu_model = _reference_model()  # same seed, same parameters (a constructed NeuRadarModel cannot be deep-copied)
    gpu_model.load_state_dict(cpu_model.state_dict())
    gpu_model = gpu_model.to(DEV)  # its `_render_weights` calls nerfacc on CUDA: the compat module of this package
    plugin.convert_neuradar_model(gpu_model)
    gpu_model.train()
    s = cpu_model.config.sampling
    g = torch.Generator().manual_seed(9)
    jit = [torch.rand((n, s.num_proposal_samples[0] + 1), generator=g), torch.rand((n, 1), generator=g), torch.rand((n, 1), generator=g)]

    rb_cpu, rb_gpu = _bundle(n, "cpu"), _bundle(n, DEV)  # (the reference's own RayBundle type)
    with FixedJitter(jit):
        ref = cpu_model.get_nff_outputs(rb_cpu, calc_lidar_losses=True)
    _loss(ref).backward()

    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast), FixedJitter(jit):
        out = gpu_model.get_nff_outputs(rb_gpu, calc_lidar_losses=True)
        loss = _loss(out)
    loss.backward()

    tol = 1e-3
    for k in ("features", "depth", "accumulation", "prop_depth_0", "prop_depth_1", "prop_weights_loss_0", "prop_weights_loss_1",
              "non_nearby_weights"):
        assert out[k].shape == ref[k].shape, k
        assert rel_err(out[k].float(), ref[k]) <= tol, (k, rel_err(out[k].float(), ref[k]))
    assert torch.equal(out["non_nearby_lidar_ray_indices"].cpu(), ref["non_nearby_lidar_ray_indices"])
    for i in range(3):
        assert rel_err(out["weights_list"][i].float(), ref["weights_list"][i]) <= tol, i
    assert abs(float(loss) - float(_loss(ref))) <= tol * abs(float(_loss(ref)))
    # every parameter of the hot path that received a gradient in the reference receives the same one here
    ref_params = dict(cpu_model.named_parameters())
    checked = 0
    for name, p in gpu_model.named_parameters():
        rg = ref_params[name].grad
        if rg is None or float(rg.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        assert rel_err(p.grad, rg) <= tol, (name, rel_err(p.grad, rg))
        checked += 1
    assert checked >= 14  # main table, 10 MLP tensors, beta, proposal table + decoder, appearance embedding
