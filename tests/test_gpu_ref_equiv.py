"""The GPU reference-equivalent baseline (baseline/ref_equiv: the oracle's per-item bodies with the reference's
one-thread-per-item launch structure) must compute what the CPU oracle computes -- otherwise the speed-up bench.py
quotes against it would compare different work."""
import numpy as np
import pytest
import torch

from handobjectconsist_b200 import synth

pytestmark = pytest.mark.gpu


def _scene(B, S, seed):
    sc = synth.make_scene(B, S, S, seed=seed)
    dev = torch.device("cuda:0")
    return sc, {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}


@pytest.mark.parametrize("detach_renders", [False, True])
def test_ref_equiv_matches_cpu_oracle(detach_renders):
    from baseline import ref_equiv
    from oracle import pipeline as opipe

    S, B = 48, 2
    sc, g = _scene(B, S, 3)
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, res = ref_equiv.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                       g["jitter_mask_ref"], g["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                       detach_renders=detach_renders, use_backward=True)
    loss.backward()
    c1 = sc["verts1"].clone().requires_grad_(True)
    loss_o, res_o = opipe.consist_step(c1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                       sc["jitter_mask_ref"], sc["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                       detach_renders=detach_renders, use_backward=True, grad_dtype=np.float32,
                                       warp_device=torch.device("cuda:0"))
    loss_o.backward()
    for i in range(2):
        assert (res["flows"][i].detach().cpu() - res_o["flows"][i].detach().cpu()).abs().max().item() <= 1e-5
    assert abs(loss.item() - loss_o.item()) <= 1e-5
    gn, go = v1.grad.cpu().numpy(), c1.grad.numpy()
    assert np.abs(gn - go).max() <= 1e-3 * max(np.abs(go).max(), 1e-30)  # float atomics: order of summation differs


def test_ref_equiv_matches_product():
    from baseline import ref_equiv
    from handobjectconsist_b200 import warpbranch
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion

    S, B = 64, 2
    sc, g = _scene(B, S, 5)
    dev = torch.device("cuda:0")
    renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                        no_light=True)
    v1 = g["verts1"].clone().requires_grad_(True)
    loss, _ = warpbranch.consist_step(v1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                      g["jitter_mask_ref"], g["jitter_mask"], renderer, PyramidCriterion("l1"), (S, S),
                                      sc["hand_ignore_faces"], detach_renders=False, use_backward=True)
    loss.backward()
    r1 = g["verts1"].clone().requires_grad_(True)
    loss_r, _ = ref_equiv.consist_step(r1, g["verts2"], g["faces"], g["K"], g["image_ref"], g["image"],
                                       g["jitter_mask_ref"], g["jitter_mask"], S, (S, S), sc["hand_ignore_faces"],
                                       detach_renders=False, use_backward=True)
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) <= 1e-4
    gn, go = v1.grad.cpu().numpy(), r1.grad.cpu().numpy()
    assert np.abs(gn - go).max() <= 1e-3 * max(np.abs(go).max(), 1e-30)
