"""Closed-hand topology -- same surface as /root/reference/meshreg/models/manoutils.py:6-33.

MANO's hand mesh (778 vertices, 1538 triangles) is open at the wrist; the photometric-consistency branch renders a
CLOSED mesh (14 extra triangles fan over the wrist ring) and masks the flow that lands on those 14 triangles, because
the wrist cap is not a surface of the real hand (warpreg.py:61-63, opticalflow.py:110-116).  The 14 triangles are
indices into MANO's fixed vertex numbering: a property of the MANO topology that a drop-in must reproduce, like the
fingertip vertex ids of ``mano.manolayer.TIP_IDS``.
"""
import torch

MANO_FACE_NB = 1538

# the wrist ring of the right-hand MANO template, triangulated as a fan (manoutils.py:10-27)
_WRIST_RING_FAN = (
    (92, 38, 122), (234, 92, 122), (239, 234, 122), (279, 239, 122), (215, 279, 122), (215, 122, 118),
    (215, 118, 117), (215, 117, 119), (215, 119, 120), (215, 120, 108), (215, 108, 79), (215, 79, 78),
    (215, 78, 121), (214, 215, 121),
)


def wrist_closing_faces():
    """[14, 3] int64: the triangles that close the wrist."""
    return torch.tensor(_WRIST_RING_FAN, dtype=torch.int64)


def get_closed_faces(mano_faces=None, mano_root="assets/mano"):
    """``(closed_faces [1552,3] int64, hand_ignore_faces)`` like the reference.

    ``mano_faces``: MANO's own face table ([1538,3], e.g. ``ManoLayer.th_faces``).  The reference reads it from the
    licence-gated MANO_RIGHT.pkl under ``mano_root``; so does this function when ``mano_faces`` is not given (and
    raises FileNotFoundError with the download hint when the file is not there).
    ``hand_ignore_faces`` are the positions of the added triangles in the closed table -- valid because they are
    appended at the end (manoutils.py:29-31): 1538 .. 1551 for the real MANO table.
    """
    if mano_faces is None:
        from .mano.manolayer import ManoLayer

        mano_faces = ManoLayer(joint_rot_mode="axisang", use_pca=False, mano_root=mano_root, center_idx=None,
                               flat_hand_mean=True).th_faces
    mano_faces = torch.as_tensor(mano_faces).long().cpu()
    if mano_faces.dim() != 2 or mano_faces.shape[1] != 3:
        raise ValueError(f"mano_faces must be [F,3], got {tuple(mano_faces.shape)}")
    closing = wrist_closing_faces()
    closed_faces = torch.cat([mano_faces, closing])
    first = mano_faces.shape[0]
    hand_ignore_faces = list(range(first, first + closing.shape[0]))
    return closed_faces, hand_ignore_faces
