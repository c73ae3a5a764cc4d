"""oracle/pipeline.py -- TEST INFRASTRUCTURE ONLY (checker / CPU baseline; never imported by the product).

CPU restatement of the whole consist path the way the reference composes it:

* ``RasterizeFunctionOracle`` ... torch autograd wrapper of the C restatement (oracle/nmr.py), same
  forward/backward plumbing as /root/reference/meshreg/neurender/rasterize.py:16-197
* ``rasterize_rgbad`` ............ rasterize.py:362-448 (permute, row flips, optional 2x SSAA) with
  ordinary differentiable torch ops
* ``render`` ..................... meshreg/neurender/renderer.py:237-295 (fill_back, projection,
  vertices_to_faces, detach_renders)
* ``get_opticalflow`` ............ meshreg/warping/opticalflow.py:51-156
* ``consist_step`` ............... meshreg/models/warpbranch.py:57-88 (flows -> pair_consist -> mean)

PARITY: the rasterizer arithmetic is UNPINNED (see nmr_oracle_impl.h); warp / loss are pinned to the
reference's own imgflowarp / lossutils through tests/golden (see oracle/warp.py).
"""
import numpy as np
import torch
from torch.autograd import Function

from . import nmr, nrfuncs, warp as owarp


class RasterizeFunctionOracle(Function):
    @staticmethod
    def forward(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb, return_alpha,
                return_depth, grad_dtype):
        fwd = nmr.rasterize_forward(faces.detach().cpu().numpy(),
                                    None if textures is None else textures.detach().cpu().numpy(), image_size, near,
                                    far, eps, background_color, return_rgb, return_alpha, return_depth)
        ctx.fwd = fwd
        ctx.grad_dtype = grad_dtype
        ctx.has_tex = textures is not None
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        empty = torch.tensor([])
        ctx.set_materialize_grads(False)
        out = (t(fwd["rgb_map"]) if return_rgb else empty, t(fwd["alpha_map"]) if return_alpha else empty,
               t(fwd["depth_map"]) if return_depth else empty, t(fwd["face_index_map"]), t(fwd["face_inv_map"]),
               t(fwd["weight_map"]))
        ctx.mark_non_differentiable(out[3])
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth, g_idx, g_inv, g_w):
        n = lambda g: None if g is None else g.detach().cpu().numpy()
        gf, gt = nmr.rasterize_backward(ctx.fwd, n(g_rgb), n(g_alpha), n(g_depth), dtype=ctx.grad_dtype)
        gf = torch.from_numpy(gf.astype(np.float32))
        gt = torch.from_numpy(gt.astype(np.float32)) if (gt is not None and ctx.has_tex) else None
        return (gf, gt) + (None,) * 9


def rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color, return_rgb=True,
                    return_alpha=True, return_depth=True, grad_dtype=np.float32, raster_fn=None):
    """``raster_fn``: a drop-in for ``RasterizeFunctionOracle.apply`` (same 11 arguments, same 6 outputs); the GPU
    reference-equivalent baseline (baseline/ref_equiv) plugs its one-thread-per-item CUDA drivers in here."""
    S = image_size * 2 if anti_aliasing else image_size
    raster_fn = RasterizeFunctionOracle.apply if raster_fn is None else raster_fn
    rgb, alpha, depth, idx, inv, w = raster_fn(faces, textures, S, near, far, eps, background_color, return_rgb,
                                               return_alpha, return_depth, grad_dtype)
    if return_rgb:
        rgb = rgb.permute(0, 3, 1, 2).flip(2)
    if return_alpha:
        alpha = alpha.flip(1)
    if return_depth:
        depth = depth.flip(1)
    if anti_aliasing:
        P = torch.nn.functional.avg_pool2d
        rgb = P(rgb, 2) if return_rgb else rgb
        alpha = P(alpha[:, None], 2)[:, 0] if return_alpha else alpha
        depth = P(depth[:, None], 2)[:, 0] if return_depth else depth
    return dict(rgb=rgb if return_rgb else None, alpha=alpha if return_alpha else None,
                depth=depth if return_depth else None, face_inv_map=inv, face_index_map=idx, weight_map=w)


def render(vertices, faces, textures, K, image_size, detach_renders=False, fill_back=True, anti_aliasing=False,
           near=0.1, far=100.0, eps=1e-3, background_color=(0, 0, 0), grad_dtype=np.float32, raster_fn=None):
    """Renderer.render with the WarpRegNet settings (R = I, t = 0, no distortion, no light)."""
    if fill_back:
        faces, textures = nrfuncs.fill_back(faces, textures)
        faces = faces.detach()
    R = torch.eye(3, dtype=vertices.dtype, device=vertices.device)[None]
    t = torch.zeros(1, 1, 3, dtype=vertices.dtype, device=vertices.device)
    dist = torch.zeros(1, 5, dtype=vertices.dtype, device=vertices.device)
    ndc = nrfuncs.projection(vertices, K, R, t, dist, float(image_size))
    f = nrfuncs.vertices_to_faces(ndc, faces)
    if detach_renders:
        f = f.detach()
    return rasterize_rgbad(f, textures, image_size, anti_aliasing, near, far, eps, background_color,
                           grad_dtype=grad_dtype, raster_fn=raster_fn)


def _ignore_mask(face_index_map, ignore_face_idxs):
    ign = face_index_map.new_tensor(list(ignore_face_idxs))
    m = (face_index_map.unsqueeze(-1) - ign).abs().min(-1)[0] != 0
    return m.flip(1).float().unsqueeze(1)


def get_opticalflow(verts_cam, faces, camintrs, image_size, orig_img_size=None, mask_occlusions=True,
                    detach_textures=False, detach_renders=True, ignore_face_idxs=None, grad_dtype=np.float32,
                    warp_device=None, raster_fn=None):
    loc1 = nrfuncs.batch_proj2d(verts_cam[0], camintrs[0])
    loc2 = nrfuncs.batch_proj2d(verts_cam[1], camintrs[1])
    d12 = loc2 - loc1
    tex = nrfuncs.batch_vertex_textures(faces, torch.cat([d12, torch.ones_like(d12[:, :, :1])], -1))
    if detach_textures:
        tex = tex.detach()
    out = render(verts_cam[0], faces, tex, camintrs[0], image_size, detach_renders, grad_dtype=grad_dtype,
                 raster_fn=raster_fn)
    mask1 = (out["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask1 = mask1 * _ignore_mask(out["face_index_map"], ignore_face_idxs)
    flow12 = out["rgb"] * mask1
    d21 = loc1 - loc2
    tex = nrfuncs.batch_vertex_textures(faces, torch.cat([d21, torch.ones_like(d21[:, :, :1])], -1))
    out = render(verts_cam[1], faces, tex, camintrs[1], image_size, detach_renders, grad_dtype=grad_dtype,
                 raster_fn=raster_fn)
    mask2 = (out["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask2 = mask2 * _ignore_mask(out["face_index_map"], ignore_face_idxs)
    flow21 = out["rgb"] * mask2
    if mask_occlusions:
        with torch.no_grad():
            mask2 = out["alpha"].unsqueeze(1)  # sic, opticalflow.py:139
            if warp_device is not None:  # the reference runs these ATen ops on CUDA
                o1, o2 = owarp.get_occlusion_mask(mask1.to(warp_device), mask2.to(warp_device),
                                                  flow12.to(warp_device), flow21.to(warp_device))
                o1, o2 = o1.to(mask1.device), o2.to(mask1.device)
            else:
                o1, o2 = owarp.get_occlusion_mask(mask1, mask2, flow12, flow21)
        mask1 = mask1 * o1.unsqueeze(1)
        mask2 = mask2 * o2.unsqueeze(1)
        flow12 = flow12 * mask1
        flow21 = flow21 * mask2
    flow12 = flow12.permute(0, 2, 3, 1)[:, :, :, :2]
    flow21 = flow21.permute(0, 2, 3, 1)[:, :, :, :2]
    if orig_img_size is not None:
        flow12 = flow12[:, : orig_img_size[1], : orig_img_size[0]]
        flow21 = flow21[:, : orig_img_size[1], : orig_img_size[0]]
    return [flow12, flow21]


def consist_step(verts1, verts2, faces, K, image_ref, image, jitter_mask_ref, jitter_mask, image_size, orig_img_size,
                 ignore_face_idxs=None, detach_renders=True, use_backward=True, grad_dtype=np.float32,
                 warp_device=None, raster_fn=None):
    """flows -> pair_consist -> mean over the batch (warpbranch.py:57-88 for one pair).

    ``warp_device``: run the warp / mask / loss ATen ops there (autograd crosses devices).  The
    reference's masks hold exact float tests (== 1, >= 0.99999) whose outcome depends on the rounding
    of ATen's CPU vs CUDA grid_sampler; the reference runs them on CUDA, so the GPU tests pass "cuda"."""
    flows = get_opticalflow([verts1, verts2], faces, [K, K], image_size, orig_img_size, True, False, detach_renders,
                            ignore_face_idxs, grad_dtype, warp_device, raster_fn)
    if warp_device is not None:
        mv = lambda t: t.to(warp_device)
        flows = [mv(f) for f in flows]
        image_ref, image, jitter_mask_ref, jitter_mask = mv(image_ref), mv(image), mv(jitter_mask_ref), mv(jitter_mask)
    loss, masks, warps, diffs = owarp.pair_consist(flows, image_ref, image, jitter_mask_ref, jitter_mask, use_backward)
    return loss.mean(), dict(flows=flows, loss=loss, masks=masks, warps=warps, diffs=diffs)
