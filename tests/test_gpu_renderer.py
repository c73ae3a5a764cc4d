"""GPU: the ``Renderer`` module surface (/root/reference/meshreg/neurender/renderer.py:12-295) -- render / rgb /
silhouettes / depth modes, fill_back, lighting, anti-aliasing, per-call intrinsics -- against the oracle built from
the same restated helpers (oracle/nrfuncs.py + oracle/nmr.py), and the op-by-op get_opticalflow path with an
anti-aliased renderer (which cannot take the fused kernels)."""
import numpy as np
import pytest
import torch

import helpers
from helpers import onmr, onr
from handobjectconsist_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_faces(sc, S, fill_back=True, light=None, tex=None):
    faces = sc["faces"]
    if fill_back:
        faces, tex = onr.fill_back(faces, tex)
    if light is not None and tex is not None:
        tex = onr.lighting(onr.vertices_to_faces(sc["verts1"], faces), tex, *light)
    ndc = onr.projection(sc["verts1"], sc["K"], torch.eye(3)[None], torch.zeros(1, 1, 3), torch.zeros(1, 5), float(S))
    return onr.vertices_to_faces(ndc, faces).numpy(), (None if tex is None else tex.numpy())


@pytest.mark.parametrize("aa,no_light", [(False, True), (True, True), (False, False)])
def test_renderer_modes_match_oracle(aa, no_light):
    from handobjectconsist_b200.neurender.renderer import Renderer
    S, B = 48, 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=31, with_object=False)
    Fn = sc["faces"].shape[1]
    tex = torch.rand(B, Fn, 2, 2, 2, 3, generator=torch.Generator().manual_seed(0))
    r = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev), K=sc["K"].to(dev),
                 orig_size=S, anti_aliasing=aa, fill_back=True, near=0.1, no_light=no_light,
                 light_intensity_ambient=0.8, light_intensity_directional=0.4, light_direction=[0, 1, 0])
    v, f, t = sc["verts1"].to(dev), sc["faces"].to(dev), tex.to(dev)
    out = r(v, f, t)                                   # mode=None -> dict
    assert set(out) == {"rgb", "alpha", "depth", "face_inv_map", "face_index_map", "weight_map"}
    light = None if no_light else (0.8, 0.4, (1, 1, 1), (1, 1, 1), (0, 1, 0))
    Sr = 2 * S if aa else S
    of, ot = _oracle_faces(sc, S, True, light, tex)
    ora = onmr.rasterize_rgbad(of.astype(np.float32), ot.astype(np.float32), S, aa, 0.1, 100.0, 1e-3, (0, 0, 0))
    for k in ("rgb", "alpha", "depth"):
        assert np.abs(out[k].cpu().numpy() - ora[k]).max() <= 1e-4, k
    assert out["face_index_map"].shape == (B, Sr, Sr)
    mism = (out["face_index_map"].cpu().numpy() != ora["face_index_map"]).mean()
    assert mism == 0.0
    # the single-output modes return the same images
    assert torch.equal(r(v, f, t, mode="rgb"), out["rgb"])
    assert torch.equal(r(v, f, mode="silhouettes"), out["alpha"])
    assert torch.equal(r(v, f, mode="depth"), out["depth"])
    with pytest.raises(ValueError):
        r(v, f, t, mode="nope")


def test_renderer_look_at_camera_runs():
    from handobjectconsist_b200.neurender.renderer import Renderer
    dev = torch.device("cuda:0")
    hv, hf = synth.hand_template()
    v = torch.from_numpy(hv * 8.0)[None].to(dev)      # unit-ish object around the origin
    f = torch.from_numpy(hf)[None].to(dev)
    r = Renderer(image_size=32, camera_mode="look_at", anti_aliasing=False, viewing_angle=30)
    sil = r(v, f, mode="silhouettes")
    assert sil.shape == (1, 32, 32) and 0.0 < sil.mean().item() < 1.0


def test_opticalflow_op_by_op_path_with_lit_renderer():
    """A renderer with lighting cannot use the fused flow kernels (the light scales the displacement texture);
    get_opticalflow must fall back to the line-by-line mirror of the reference and stay differentiable."""
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.warping.opticalflow import get_opticalflow, _fused_path_ok
    S, B = 40, 2
    dev = torch.device("cuda:0")
    sc = synth.make_scene(B, S, S, seed=32)
    kw = dict(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
              K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1)
    lit = Renderer(no_light=False, light_intensity_ambient=1.0, light_intensity_directional=0.0, **kw)
    plain = Renderer(no_light=True, **kw)
    assert not _fused_path_ok(lit) and _fused_path_ok(plain)
    outs = []
    for r in (lit, plain):
        v1 = sc["verts1"].to(dev).requires_grad_(True)
        flows = get_opticalflow([v1, sc["verts2"].to(dev)], sc["faces"].to(dev), [sc["K"].to(dev)] * 2, r, (S, S),
                                ignore_face_idxs=sc["hand_ignore_faces"])
        (flows[0].sum() + flows[1].sum()).backward()
        assert torch.isfinite(v1.grad).all() and v1.grad.abs().max().item() > 0
        outs.append(flows)
    # ambient light of intensity 1 and no directional light leaves the textures unchanged: same flows
    for i in range(2):
        assert (outs[0][i] - outs[1][i]).abs().max().item() <= 1e-4
