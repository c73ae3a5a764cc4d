"""Photometric-consistency branch -- same surface as /root/reference/meshreg/models/warpbranch.py:9-96.

``forward(samples, all_results, hand_face, renderer, image_size, criterion, ...)`` gathers the hand and
object vertices of every frame, swaps in ground truth for the reference frames (``gt_refs``), detaches
all but the first frame (``first_only``), renders the mesh-induced flows and returns the mean pair loss.
Samples may be keyed by the reference's own ``TransQueries`` / ``BaseQueries`` enums or by ours -- the
lookup goes by member NAME.
"""
import torch

from . import _config, consist
from .meshutils import batch_cat_meshes, cat_hand_object_pair
from .warping import imgflowarp, opticalflow


def _get(sample, name):
    for key, val in sample.items():
        if getattr(key, "name", key) == name:
            return val
    raise KeyError(name)


def _trans(sample, name):
    # TransQueries and BaseQueries share member names; the reference reads IMAGE / JITTERMASK / CAMINTR
    # from the transformed queries and faces / GT vertices from the base ones.
    for key, val in sample.items():
        if getattr(key, "name", key) == name and type(key).__name__ == "TransQueries":
            return val
    return _get(sample, name)


def _base(sample, name):
    for key, val in sample.items():
        if getattr(key, "name", key) == name and type(key).__name__ == "BaseQueries":
            return val
    return _get(sample, name)


def forward(samples, all_results, hand_face, renderer, image_size, criterion, gt_refs=True, first_only=True,
            hand_ignore_faces=None, use_backward=True, detach_renders=True, return_visuals=True, loss_only=False):
    """
    Args:
        use_backward: also compare the warp from the first to the other frame (warpbranch.py:21-25)
        detach_renders: extension -- the reference hard-codes True (warpbranch.py:66); False lets the
            NMR geometry gradient flow as well
        return_visuals: extension -- False skips the visualisation entries of ``pair_results`` (``warps``, ``diffs``
            and the masks' ``warp_mask`` are None): nothing the loss or its gradient depends on
        loss_only: extension (frame-pair path, with return_visuals=False) -- ``recons_flows`` and the masks'
            ``flow_mask`` are None as well: a step that only wants the loss and its gradient
    Returns (full_loss, pair_results) like the reference.
    """
    # Put inputs on GPU (stream-ordered copies when the host tensors are pinned)
    images = [_trans(sample, "IMAGE").cuda(non_blocking=True) for sample in samples]
    jitter_masks = [_trans(sample, "JITTERMASK").cuda(non_blocking=True) for sample in samples]
    camintrs = [_trans(sample, "CAMINTR").cuda(non_blocking=True) for sample in samples]

    obj_verts = [result["recov_objverts3d"] for result in all_results]
    obj_faces = [_base(sample, "OBJFACES").cuda(non_blocking=True).long() for sample in samples]
    hand_verts = [result["recov_handverts3d"] for result in all_results]
    if gt_refs:
        for sample_idx in range(1, len(samples)):
            obj_verts[sample_idx] = _base(samples[sample_idx], "OBJVERTS3D").cuda(non_blocking=True)
            hand_verts[sample_idx] = _base(samples[sample_idx], "HANDVERTS3D").cuda(non_blocking=True)
    verts_world = []
    last = len(samples) - 1
    if (len(samples) == 2 and _config.fused_pair and hand_face.dim() in (2, 3)
            and consist.pair_path_ok(renderer, criterion, images, jitter_masks, image_size)):
        # one frame pair, the renderer WarpRegNet builds, L1: the whole step behind one autograd node
        # (consist.py: 5 launches forward, 4 backward, both renders stacked along the batch)
        hand2, obj2 = (hand_verts[1].detach(), obj_verts[1].detach()) if first_only else (hand_verts[1], obj_verts[1])
        (warp_loss, warp_mean), flows, masks, warps, diffs = consist.pair_consist_step(
            hand_verts[0], obj_verts[0], hand2, obj2, hand_face.cuda(non_blocking=True), obj_faces[last], camintrs[0],
            camintrs[1], images[0], images[1], jitter_masks[0], jitter_masks[1], renderer, image_size,
            hand_ignore_faces=hand_ignore_faces, detach_renders=detach_renders, use_backward=use_backward,
            return_visuals=return_visuals, loss_only=loss_only)
        # the mean over the single pair's per-sample losses (warpbranch.py:88) comes out of the same launch
        pair_results = {"masks": [masks], "warps": [warps], "recons_flows": [flows], "diffs": [diffs],
                        "diff_losses": warp_loss.unsqueeze(0)}
        return warp_mean, pair_results
    if len(samples) == 2 and hand_face.dim() in (2, 3) and (hand_face.dim() == 2 or hand_face.shape[0] == 1):
        # one frame pair (the training setting): both concatenations and the face table in one launch
        verts_a, verts_b, all_faces = cat_hand_object_pair(hand_verts[0], obj_verts[0], hand_verts[1], obj_verts[1],
                                                           hand_face.cuda(non_blocking=True), obj_faces[last])
        verts_world = [verts_a, verts_b.detach() if first_only else verts_b]
    else:
        hand_faces_b = hand_face.repeat(obj_verts[0].shape[0], 1, 1).long()
        hand_faces = [hand_faces_b for _ in range(len(samples))]
        for seq_idx in range(len(samples)):
            if seq_idx == last:
                # the reference concatenates the faces of every frame but only the last frame's survive the loop
                # (warpbranch.py:50-52,57-60): build that one
                all_verts, all_faces, _ = batch_cat_meshes([hand_verts[seq_idx], obj_verts[seq_idx]],
                                                           [hand_faces[seq_idx], obj_faces[seq_idx]])
            else:
                all_verts = torch.cat([hand_verts[seq_idx], obj_verts[seq_idx]], 1)
            if first_only and seq_idx > 0:
                all_verts = all_verts.detach()
            verts_world.append(all_verts)

    recons_flows = opticalflow.get_opticalflows(verts_world, all_faces, camintrs, renderer, image_size,
                                                detach_textures=False, detach_renders=detach_renders,
                                                ignore_face_idxs=hand_ignore_faces)
    all_masks, all_warps, all_diffs, full_losses = [], [], [], []
    for recons_flow, image, jitter_mask in zip(recons_flows, images[1:], jitter_masks[1:]):
        warp_loss, masks, warps, diffs = imgflowarp.pair_consist(
            recons_flow, image_ref=images[0], image=image, jitter_mask_ref=jitter_masks[0], jitter_mask=jitter_mask,
            criterion=criterion, use_backward=use_backward)
        all_masks.append(masks)
        full_losses.append(warp_loss)
        all_warps.append(warps)
        all_diffs.append(diffs)
    # torch.stack of a single [B] tensor is that tensor with a leading axis: no copy kernel for one pair
    stack_losses = full_losses[0].unsqueeze(0) if len(full_losses) == 1 else torch.stack(full_losses)
    full_loss = stack_losses.mean()
    pair_results = {"masks": all_masks, "warps": all_warps, "recons_flows": recons_flows, "diffs": all_diffs,
                    "diff_losses": stack_losses}
    return full_loss, pair_results


def consist_step(verts1, verts2, faces, K, image_ref, image, jitter_mask_ref, jitter_mask, renderer, criterion,
                 orig_img_size, ignore_face_idxs=None, detach_renders=True, use_backward=True):
    """One frame pair given already-concatenated meshes: flows -> pair_consist -> batch mean, through the
    operator-by-operator surface (get_opticalflow, pair_consist).  (``forward`` takes the fused pair path of
    consist.py when it can; this entry keeps exercising the reference-shaped operators.)"""
    flows = opticalflow.get_opticalflow([verts1, verts2], faces, [K, K], renderer, orig_img_size,
                                        detach_textures=False, detach_renders=detach_renders,
                                        ignore_face_idxs=ignore_face_idxs)
    loss, masks, warps, diffs = imgflowarp.pair_consist(flows, image_ref, image, jitter_mask_ref, jitter_mask,
                                                        criterion, use_backward)
    return loss.mean(), dict(flows=flows, loss=loss, masks=masks, warps=warps, diffs=diffs)
