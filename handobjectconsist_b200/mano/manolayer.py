"""``ManoLayer`` -- same surface as ``manopth.manolayer.ManoLayer`` as the reference uses it
(/root/reference/meshreg/models/manobranch.py:70-85,139-145, meshreg/models/warpreg.py:54-60,
meshreg/models/manoutils.py:7-9): constructor keywords ``center_idx, flat_hand_mean, ncomps, side, mano_root,
use_pca, root_rot_mode, joint_rot_mode``; ``forward(th_pose_coeffs, th_betas, th_trans)`` returns
``(verts [B,778,3], joints [B,21,3])`` in millimetres; buffer ``th_faces``.

The skinning itself is one CUDA launch per direction (csrc/mano_lbs.cu through the C ABI).  The licensed
MANO_{RIGHT,LEFT}.pkl files cannot be shipped or downloaded here, so the parameter set is passed in as a dict
(``model=``, e.g. ``handobjectconsist_b200.synth.mano_model()`` or the arrays of a user's own MANO pickle:
``v_template, shapedirs, posedirs, J_regressor -> j_regressor, weights, hands_components, hands_mean, f``);
``mano_root`` is only used when ``model`` is not given and a pickle is actually there.
"""
import ctypes
import os
import pickle

import numpy as np
import torch
from torch.autograd import Function

from .. import _lib

TIP_IDS = {"right": (745, 317, 444, 556, 673), "left": (745, 317, 445, 556, 673)}


def _load_pickle(mano_root, side):
    path = os.path.join(mano_root, f"MANO_{side.upper()}.pkl")
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: the MANO model is licence-gated (https://mano.is.tue.mpg.de). Pass the arrays as "
            "`model=` (a dict) or use handobjectconsist_b200.synth.mano_model() for a synthetic stand-in.")
    with open(path, "rb") as f:
        raw = pickle.load(f, encoding="latin1")
    g = lambda k: np.asarray(getattr(raw[k], "r", raw[k]), dtype=np.float64)
    jreg = raw["J_regressor"]
    jreg = np.asarray(jreg.todense() if hasattr(jreg, "todense") else jreg, dtype=np.float64)
    return dict(v_template=g("v_template"), shapedirs=g("shapedirs"), posedirs=g("posedirs"), j_regressor=jreg,
                weights=g("weights"), hands_components=g("hands_components"), hands_mean=g("hands_mean"),
                faces=np.asarray(raw["f"], dtype=np.int64))


class _ManoFunction(Function):
    @staticmethod
    def forward(ctx, pose, betas, trans, layer):
        _lib.require_cuda(pose, betas, trans, what="ManoLayer")
        L = _lib.lib()
        p = pose.detach().contiguous().float()
        b = None if betas is None else betas.detach().contiguous().float()
        t = None if trans is None else trans.detach().contiguous().float()
        B = p.shape[0]
        dev = p.device
        st = layer._struct(dev)
        with torch.cuda.device(dev):
            verts = torch.empty((B, layer.num_verts, 3), dtype=torch.float32, device=dev)
            joints = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
            _lib.check(L.hoc_mano_forward(ctypes.byref(st), _lib.ptr(p), _lib.ptr(b), _lib.ptr(t), B, _lib.ptr(verts),
                                          _lib.ptr(joints), _lib.stream_ptr()), "hoc_mano_forward")
        ctx.layer = layer
        ctx.save_for_backward(p, b, t)
        ctx.set_materialize_grads(False)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        p, b, t = ctx.saved_tensors
        if g_verts is None and g_joints is None:
            return None, None, None, None
        L = _lib.lib()
        layer = ctx.layer
        dev = p.device
        B = p.shape[0]
        st = layer._struct(dev)
        c = lambda g: None if g is None else g.contiguous().float()
        g_verts, g_joints = c(g_verts), c(g_joints)
        with torch.cuda.device(dev):
            gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
            gb = torch.empty_like(b) if (b is not None and ctx.needs_input_grad[1]) else None
            gt = torch.empty_like(t) if (t is not None and ctx.needs_input_grad[2]) else None
            ws_bytes = L.hoc_mano_backward_workspace_bytes(B)
            ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
            _lib.check(L.hoc_mano_backward(ctypes.byref(st), _lib.ptr(p), _lib.ptr(b), _lib.ptr(t), _lib.ptr(g_verts),
                                           _lib.ptr(g_joints), B, _lib.ptr(gp), _lib.ptr(gb), _lib.ptr(gt),
                                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "hoc_mano_backward")
        return gp, gb, gt, None


class ManoLayer(torch.nn.Module):
    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side="right", mano_root="mano/models",
                 use_pca=True, root_rot_mode="axisang", joint_rot_mode="axisang", robust_rot=False, model=None):
        super().__init__()
        if root_rot_mode != "axisang" or joint_rot_mode != "axisang":
            raise NotImplementedError("only axis-angle rotations (the modes the reference uses) are implemented")
        if side not in TIP_IDS:
            raise ValueError("side must be 'right' or 'left'")
        self.center_idx, self.side, self.use_pca = center_idx, side, use_pca
        self.flat_hand_mean, self.robust_rot = flat_hand_mean, robust_rot
        self.rot = 3
        self.ncomps = ncomps if use_pca else 45
        if model is None:
            model = _load_pickle(mano_root, side)
        t = lambda a: torch.as_tensor(np.asarray(a.cpu() if torch.is_tensor(a) else a), dtype=torch.float32).contiguous()
        self.register_buffer("th_v_template", t(model["v_template"]))
        self.register_buffer("th_shapedirs", t(model["shapedirs"]))
        self.register_buffer("th_posedirs", t(model["posedirs"]))
        self.register_buffer("th_J_regressor", t(model["j_regressor"]))
        self.register_buffer("th_weights", t(model["weights"]))
        # constants derived once per model for the kernels (include/hoc_b200.h, hoc_mano_model): transposed pose blend
        # shapes, and the joint regression folded through the template / the shape blend shapes
        self.register_buffer("th_posedirs_t", self.th_posedirs.reshape(-1, 135).t().contiguous())
        self.register_buffer("th_J_template", (self.th_J_regressor @ self.th_v_template).contiguous())
        self.register_buffer("th_J_shapedirs",
                             torch.einsum("jv,vck->jck", self.th_J_regressor, self.th_shapedirs).contiguous())
        comps = t(model["hands_components"])
        self.register_buffer("th_comps", comps)
        self.register_buffer("th_selected_comps", comps[: self.ncomps].contiguous())
        mean = torch.zeros(45) if flat_hand_mean else t(model["hands_mean"])
        self.register_buffer("th_hands_mean", mean.reshape(45).contiguous())
        faces = model["faces"]
        self.register_buffer("th_faces", torch.as_tensor(np.asarray(faces.cpu() if torch.is_tensor(faces) else faces),
                                                         dtype=torch.int64))
        self.num_verts = self.th_v_template.shape[0]
        self.tip_ids = tuple(model.get("tip_ids", TIP_IDS[side]))
        if self.th_posedirs.shape != (self.num_verts, 3, 135) or self.th_weights.shape != (self.num_verts, 16):
            raise ValueError("model tensors do not have MANO's shapes")

    def _struct(self, device):
        if self.th_v_template.device != device:
            raise RuntimeError(f"ManoLayer buffers live on {self.th_v_template.device}, input on {device}: call .cuda()")
        st = _lib.ManoModelStruct()
        st.v_template = self.th_v_template.data_ptr()
        st.shapedirs = self.th_shapedirs.data_ptr()
        st.posedirs = self.th_posedirs.data_ptr()
        st.j_regressor = self.th_J_regressor.data_ptr()
        st.posedirs_t = self.th_posedirs_t.data_ptr()
        st.j_template = self.th_J_template.data_ptr()
        st.j_shapedirs = self.th_J_shapedirs.data_ptr()
        st.weights = self.th_weights.data_ptr()
        st.hands_components = self.th_selected_comps.data_ptr()
        st.hands_mean = self.th_hands_mean.data_ptr()
        st.num_verts = self.num_verts
        st.ncomps = self.ncomps
        st.use_pca = int(self.use_pca)
        st.center_idx = -1 if self.center_idx is None else int(self.center_idx)
        for k in range(5):
            st.tip_ids[k] = int(self.tip_ids[k])
        return st

    def forward(self, th_pose_coeffs, th_betas=None, th_trans=None, root_palm=False, share_betas=False):
        """(verts, joints) in millimetres.  ``th_betas`` / ``th_trans`` with a single element (the reference passes
        ``torch.Tensor([0])``, manobranch.py:128-129) mean "zero shape" / "centre on center_idx"."""
        if bool(root_palm) or bool(share_betas):
            raise NotImplementedError("root_palm / share_betas are not used by the reference and not implemented")
        if not th_pose_coeffs.is_cuda:
            raise TypeError("ManoLayer supports only cuda Tensors (this package has no CPU path)")
        expect = self.rot + self.ncomps
        if th_pose_coeffs.dim() != 2 or th_pose_coeffs.shape[1] != expect:
            raise ValueError(f"pose must be [B, {expect}], got {tuple(th_pose_coeffs.shape)}")
        betas = None if (th_betas is None or th_betas.numel() == 1) else th_betas.to(th_pose_coeffs.device)
        trans = None
        if th_trans is not None and th_trans.numel() > 1:
            th_trans = th_trans.to(th_pose_coeffs.device)
            if bool(torch.norm(th_trans) != 0):  # manopth's own test (a device sync, only when a translation is given)
                trans = th_trans
        return _ManoFunction.apply(th_pose_coeffs, betas, trans, self)
