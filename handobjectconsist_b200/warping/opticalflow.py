"""Mesh-pair -> rendered optical flow -- same surface as /root/reference/meshreg/warping/opticalflow.py
(``get_opticalflows`` :10-48, ``get_opticalflow`` :51-156).

Per-vertex 2-D displacements between the two frames are painted on the mesh as a texture
([dx, dy, 1]) and rendered at each frame; the rendered field drives the image warp.  Reference
quirks are kept: ``mask_flow2`` is re-bound to the un-thresholded alpha inside the no-grad block
(opticalflow.py:139), ``face_index_map`` is not row-flipped by the rasterizer so the ignore mask is
flipped here (:114,132), and only front copies of the ignored faces (indices < F) are matched.
"""
from typing import List

import torch

from ..meshutils import batch_proj2d, batch_vertex_textures
from . import imgflowarp


def get_opticalflows(verts_cam: List[torch.Tensor], faces: torch.Tensor, camintrs: List[torch.Tensor], neurenderer,
                     orig_img_size=None, detach_textures: bool = False, detach_renders: bool = False,
                     ignore_face_idxs=None):
    """Flows between the first mesh and every other one (opticalflow.py:10-48)."""
    all_flows = []
    for vert_world, camintr in zip(verts_cam[1:], camintrs[1:]):
        flows = get_opticalflow([verts_cam[0], vert_world], faces, [camintrs[0], camintr], neurenderer,
                                orig_img_size=orig_img_size, detach_textures=detach_textures,
                                detach_renders=detach_renders, ignore_face_idxs=ignore_face_idxs)
        all_flows.append(flows)
    return all_flows


def _ignore_mask(face_index_map, ignore_face_idxs):
    """1 where the covering face is NOT in the ignore list, rows flipped to image order
    (opticalflow.py:110-115; ``isin`` replaces the [B,S,S,len(ignore)] abs-min temporary)."""
    ign = torch.as_tensor(list(ignore_face_idxs), dtype=face_index_map.dtype, device=face_index_map.device)
    keep = ~torch.isin(face_index_map, ign)
    return keep.flip(1).float().unsqueeze(1)


def get_opticalflow(verts_cam: List[torch.Tensor], faces: torch.Tensor, camintrs: List[torch.Tensor], neurenderer,
                    orig_img_size=None, mask_occlusions: bool = True, detach_textures: bool = False,
                    detach_renders: bool = True, ignore_face_idxs=None):
    """
    Rendered flow 1->2 at the pixels of mesh 1 and 2->1 at the pixels of mesh 2 (opticalflow.py:51-156).
    Returns [pred_flow12, pred_flow21], each [B,H,W,2] (cropped to ``orig_img_size`` = (W, H)).
    """
    locs2d_1 = batch_proj2d(verts_cam[0], camintrs[0])
    locs2d_2 = batch_proj2d(verts_cam[1], camintrs[1])
    displ_12 = locs2d_2 - locs2d_1
    sample_flows = torch.cat([displ_12, torch.ones_like(displ_12[:, :, :1])], -1)
    all_textures = batch_vertex_textures(faces, sample_flows)
    if detach_textures:
        all_textures = all_textures.detach()

    renderout = neurenderer(verts_cam[0], faces, all_textures, K=camintrs[0], detach_renders=detach_renders)
    mask_flow1 = (renderout["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask_flow1 = mask_flow1 * _ignore_mask(renderout["face_index_map"], ignore_face_idxs)
    pred_flow12 = renderout["rgb"] * mask_flow1

    displ_21 = locs2d_1 - locs2d_2
    sample_flows = torch.cat([displ_21, torch.ones_like(displ_21[:, :, :1])], -1)
    all_textures = batch_vertex_textures(faces, sample_flows)

    renderout = neurenderer(verts_cam[1], faces, all_textures, K=camintrs[1], detach_renders=detach_renders)
    mask_flow2 = (renderout["alpha"].unsqueeze(1) > 0.99999).float()
    if ignore_face_idxs is not None:
        mask_flow2 = mask_flow2 * _ignore_mask(renderout["face_index_map"], ignore_face_idxs)
    pred_flow21 = renderout["rgb"] * mask_flow2

    if mask_occlusions:
        with torch.no_grad():
            mask_flow2 = renderout["alpha"].unsqueeze(1)  # sic (opticalflow.py:139)
            occl_mask1, occl_mask2 = imgflowarp.get_occlusion_mask(mask_flow1, mask_flow2, pred_flow12, pred_flow21)
        mask_flow1 = mask_flow1 * occl_mask1.unsqueeze(1)
        mask_flow2 = mask_flow2 * occl_mask2.unsqueeze(1)
        pred_flow12 = pred_flow12 * mask_flow1
        pred_flow21 = pred_flow21 * mask_flow2
    pred_flow12 = pred_flow12.permute(0, 2, 3, 1)[:, :, :, :2]
    pred_flow21 = pred_flow21.permute(0, 2, 3, 1)[:, :, :, :2]
    if orig_img_size is not None:
        pred_flow12 = pred_flow12[:, : orig_img_size[1], : orig_img_size[0]]
        pred_flow21 = pred_flow21[:, : orig_img_size[1], : orig_img_size[0]]
    return [pred_flow12, pred_flow21]
