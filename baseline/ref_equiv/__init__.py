"""baseline/ref_equiv -- BENCHMARK BASELINE ONLY: the reference's rasterizer launch structure
restated on the GPU, behind the reference's own wrapper plumbing, so that `bench.py` can quote this library against
"what a straightforward CUDA build of the reference's five kernels does on the same B200".

* ``RasterizeFunctionRefEquiv`` -- autograd Function with the buffer allocation / fill / clone / background /
  alpha plumbing of /root/reference/meshreg/neurender/rasterize.py:24-197 on CUDA tensors, calling the five
  one-thread-per-item kernels of ``ref_equiv.cu`` (= the CPU oracle's per-item bodies, compiled for the device).
* ``consist_step`` -- oracle.pipeline's composition (renderer.py:237-295, opticalflow.py:51-156,
  imgflowarp.py:58-115 with ATen ``grid_sample``, row flips, separate element-wise ops) on CUDA tensors.

It is a RESTATEMENT (the upstream wheel `neural-renderer-pytorch` is absent and not installable): labelled as such
wherever its numbers appear.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import torch
from torch.autograd import Function

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libref_equiv.so")
    srcs = [os.path.join(_HERE, "ref_equiv.cu"), os.path.join(_HERE, "..", "..", "oracle", "nmr_oracle_impl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libref_equiv.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libref_equiv.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(code, what):
    if code != 0:
        raise RuntimeError(f"{what}: CUDA error {code}")


class RasterizeFunctionRefEquiv(Function):
    """rasterize.py:16-197 on CUDA: same buffers, same fills, same five kernel calls, same saved tensors."""

    @staticmethod
    def forward(ctx, faces, textures, image_size, near, far, eps, background_color, return_rgb, return_alpha,
                return_depth, grad_dtype=None):
        L = lib()
        faces = faces.detach().contiguous().float()
        B, nf = faces.shape[:2]
        S = int(image_size)
        dev = faces.device
        c_i, c_f = ctypes.c_int, ctypes.c_float
        # rasterize.py:58-85
        face_index_map = torch.empty((B, S, S), dtype=torch.int32, device=dev).fill_(-1)
        weight_map = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev).fill_(0.0)
        depth_map = torch.empty((B, S, S), dtype=torch.float32, device=dev).fill_(far)
        rgb_map = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev).fill_(0) if return_rgb else torch.zeros(1, device=dev)
        alpha_map = torch.empty((B, S, S), dtype=torch.float32, device=dev).fill_(0) if return_alpha else torch.zeros(1, device=dev)
        face_inv_map = (torch.empty((B, S, S, 3, 3), dtype=torch.float32, device=dev).fill_(0) if return_depth
                        else torch.zeros(1, device=dev))
        faces_inv = torch.zeros_like(faces)
        _chk(L.refeq_forward_face_index_map(_p(faces), _p(face_index_map), _p(weight_map), _p(depth_map),
                                            _p(face_inv_map), _p(faces_inv), c_i(B), c_i(nf), c_i(S), c_f(near),
                                            c_f(far), c_i(int(return_depth)), _st()), "forward_face_index_map")
        sampling_index_map = sampling_weight_map = None
        ts = 0
        if return_rgb:
            textures = textures.detach().contiguous().float()
            ts = textures.shape[2]
            sampling_index_map = torch.empty((B, S, S, 8), dtype=torch.int32, device=dev).fill_(0)
            sampling_weight_map = torch.empty((B, S, S, 8), dtype=torch.float32, device=dev).fill_(0)
            _chk(L.refeq_forward_texture_sampling(_p(faces), _p(textures), _p(face_index_map), _p(weight_map),
                                                  _p(depth_map), _p(rgb_map), _p(sampling_index_map),
                                                  _p(sampling_weight_map), c_i(B), c_i(nf), c_i(S), c_i(ts), c_f(eps),
                                                  _st()), "forward_texture_sampling")
            # forward_background, rasterize.py:252-260
            bg = torch.as_tensor(background_color, dtype=torch.float32, device=dev)
            mask = (face_index_map >= 0).float()[:, :, :, None]
            bg = bg[None, None, None, :] if bg.ndimension() == 1 else bg[:, None, None, :]
            rgb_map = rgb_map * mask + (1 - mask) * bg
        if return_alpha:  # forward_alpha_map, rasterize.py:246-249
            alpha_map = (face_index_map >= 0).float()
        ctx.save_for_backward(faces, textures if return_rgb else None, face_index_map, weight_map, depth_map, rgb_map,
                              alpha_map, face_inv_map, sampling_index_map, sampling_weight_map)
        ctx.cfg = (B, nf, S, ts, float(eps), bool(return_rgb), bool(return_alpha), bool(return_depth))
        ctx.mark_non_differentiable(face_index_map)
        ctx.set_materialize_grads(False)
        # rasterize.py:118-125 clones its outputs
        return (rgb_map.clone(), alpha_map.clone(), depth_map.clone(), face_index_map, face_inv_map, weight_map)

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth, g_idx, g_inv, g_w):
        (faces, textures, face_index_map, weight_map, depth_map, rgb_map, alpha_map, face_inv_map, sampling_index_map,
         sampling_weight_map) = ctx.saved_tensors
        B, nf, S, ts, eps, rr, ra, rd = ctx.cfg
        L = lib()
        dev = faces.device
        c_i, c_f = ctypes.c_int, ctypes.c_float
        z = lambda shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        # rasterize.py:151-181: zero-filled gradients, materialised incoming gradients
        grad_faces = z(faces.shape)
        grad_textures = z(textures.shape) if rr else None
        g_rgb = (g_rgb.contiguous().float() if g_rgb is not None else z(rgb_map.shape)) if rr else z((1,))
        g_alpha = (g_alpha.contiguous().float() if g_alpha is not None else z(alpha_map.shape)) if ra else z((1,))
        g_depth = (g_depth.contiguous().float() if g_depth is not None else z(depth_map.shape)) if rd else z((1,))
        if rr or ra:
            _chk(L.refeq_backward_pixel_map(_p(faces), _p(face_index_map), _p(rgb_map), _p(alpha_map), _p(g_rgb),
                                            _p(g_alpha), _p(grad_faces), c_i(B), c_i(nf), c_i(S), c_f(eps), c_i(int(rr)),
                                            c_i(int(ra)), _st()), "backward_pixel_map")
        if rr:
            _chk(L.refeq_backward_textures(_p(face_index_map), _p(sampling_weight_map), _p(sampling_index_map),
                                           _p(g_rgb), _p(grad_textures), c_i(B), c_i(nf), c_i(S), c_i(ts), _st()),
                 "backward_textures")
        if rd:
            _chk(L.refeq_backward_depth_map(_p(faces), _p(depth_map), _p(face_index_map), _p(face_inv_map),
                                            _p(weight_map), _p(g_depth), _p(grad_faces), c_i(B), c_i(nf), c_i(S), _st()),
                 "backward_depth_map")
        return (grad_faces, grad_textures) + (None,) * 9


def consist_step(verts1, verts2, faces, K, image_ref, image, jitter_mask_ref, jitter_mask, image_size, orig_img_size,
                 ignore_face_idxs=None, detach_renders=True, use_backward=True):
    """The reference's composition of one frame pair (warpbranch.py:57-88) on CUDA tensors, op by op."""
    from oracle import pipeline as opipe  # the composition is shared with the CPU checker; only the kernels differ

    return opipe.consist_step(verts1, verts2, faces, K, image_ref, image, jitter_mask_ref, jitter_mask, image_size,
                              orig_img_size, ignore_face_idxs, detach_renders, use_backward,
                              raster_fn=RasterizeFunctionRefEquiv.apply)
