"""Synthetic inputs of the shapes the hot path sees: a MANO-sized closed hand
mesh (778 vertices / 1552 faces, the last 14 play the wrist-closing faces of
/root/reference/meshreg/models/manoutils.py:10-33), a 1502-vertex / 3000-face object, FPHAB-like
pinhole intrinsics, frame pairs that differ by a small rigid motion, random images and jitter masks.
Datasets and the licensed MANO model are not available offline, so bench.py and the tests use these.
Everything is generated with numpy from an explicit seed and is deterministic.
"""
import numpy as np
import torch

HAND_VERTS, HAND_FACES = 778, 1552
OBJ_VERTS, OBJ_FACES = 1502, 3000
HAND_IGNORE_FACES = list(range(1538, 1552))


def icosphere(subdivisions):
    """Unit icosphere: (verts [V,3] float64, faces [F,3] int64), outward (counter-clockwise) winding."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.asarray(v, dtype=np.float64) / np.linalg.norm(v) for v in verts]
    for _ in range(subdivisions):
        cache = {}
        new_faces = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.stack(verts), np.asarray(faces, dtype=np.int64)


def split_edges(verts, faces, count, rng):
    """``count`` edge splits of a closed manifold mesh: each adds 1 vertex and 2 faces (Euler-safe)."""
    verts = [v for v in verts]
    faces = [tuple(f) for f in faces]
    for _ in range(count):
        fi = int(rng.integers(len(faces)))
        a, b, c = faces[fi]
        # neighbour across edge (a, b) holds it reversed: (b, a, d) up to rotation
        fj, d = -1, -1
        for j, g in enumerate(faces):
            if j == fi:
                continue
            for r in range(3):
                if g[r] == b and g[(r + 1) % 3] == a:
                    fj, d = j, g[(r + 2) % 3]
                    break
            if fj >= 0:
                break
        assert fj >= 0, "mesh is not closed"
        m = len(verts)
        verts.append((verts[a] + verts[b]) / 2.0)
        faces[fi] = (a, m, c)
        faces[fj] = (b, m, d)
        faces.append((m, b, c))
        faces.append((m, a, d))
    return np.stack(verts), np.asarray(faces, dtype=np.int64)


_MESH_CACHE = {}


def hand_template(seed=0):
    """(verts [778,3] metres, faces [1552,3]) -- elongated closed blob of MANO's size."""
    key = ("hand", seed)
    if key not in _MESH_CACHE:
        rng = np.random.default_rng(seed)
        v, f = icosphere(3)  # 642 / 1280
        v, f = split_edges(v, f, HAND_VERTS - v.shape[0], rng)
        assert v.shape[0] == HAND_VERTS and f.shape[0] == HAND_FACES
        v = v / np.linalg.norm(v, axis=1, keepdims=True)
        v = v * np.array([0.045, 0.09, 0.015])  # ~ 9 x 18 x 3 cm
        _MESH_CACHE[key] = (v.astype(np.float32), f)
    return _MESH_CACHE[key]


def object_template(seed=1):
    """(verts [1502,3] metres, faces [3000,3]) -- star-convex blob, radius 4-8 cm."""
    key = ("obj", seed)
    if key not in _MESH_CACHE:
        rng = np.random.default_rng(seed)
        v, f = icosphere(3)
        v, f = split_edges(v, f, OBJ_VERTS - v.shape[0], rng)
        assert v.shape[0] == OBJ_VERTS and f.shape[0] == OBJ_FACES
        v = v / np.linalg.norm(v, axis=1, keepdims=True)
        lobes = sum(rng.normal() * np.sin((k + 1) * np.arctan2(v[:, 1], v[:, 0]) + rng.uniform(0, 6.28)) *
                    np.cos((k + 1) * np.arccos(np.clip(v[:, 2], -1, 1))) for k in range(3))
        radius = 0.06 + 0.012 * lobes / max(np.abs(lobes).max(), 1e-6)
        _MESH_CACHE[key] = ((v * radius[:, None]).astype(np.float32), f)
    return _MESH_CACHE[key]


def _rodrigues(rvec):
    theta = np.linalg.norm(rvec)
    if theta < 1e-12:
        return np.eye(3)
    k = rvec / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * K @ K


def camera_intrinsics(batch, width, height, device="cpu"):
    """FPHAB-quarter intrinsics (fhbhands.py:83,115,129) rescaled to ``width`` x ``height``."""
    fx = 1395.749 / 4.0 * width / 480.0
    K = np.array([[fx, 0, width / 2.0 - 0.5], [0, fx, height / 2.0 - 0.5], [0, 0, 1]], dtype=np.float32)
    return torch.from_numpy(np.tile(K[None], (batch, 1, 1))).to(device)


def make_scene(batch, width=256, height=256, seed=0, device="cpu", with_object=True, motion=True):
    """One synthetic frame pair per sample.

    Returns a dict: verts1 / verts2 [B,V,3] camera-space metres (V = 2280 with the object), faces
    [B,F,3] int64 (hand first, F = 4552), K [B,3,3], image_ref / image [B,3,H,W] in (-0.5, 0.5),
    jitter_mask_ref / jitter_mask [B,3,H,W] (ones with a random zero border band),
    hand_ignore_faces (list).
    """
    rng = np.random.default_rng(seed)
    hv, hf = hand_template()
    ov, of = object_template()
    verts1, verts2 = [], []
    for _ in range(batch):
        parts1, parts2 = [], []
        centre = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.3, 0.55)])
        for name, tmpl in (("hand", hv), ("obj", ov)):
            if name == "obj" and not with_object:
                continue
            R = _rodrigues(rng.normal(size=3) * 1.5)
            off = centre + (rng.uniform(-0.05, 0.05, size=3) if name == "obj" else 0.0)
            p1 = tmpl @ R.T + off
            if motion:
                dR = _rodrigues(rng.normal(size=3) * 0.05)
                p2 = (p1 - off) @ dR.T + off + rng.normal(size=3) * 0.005
            else:
                p2 = p1.copy()
            parts1.append(p1)
            parts2.append(p2)
        verts1.append(np.concatenate(parts1))
        verts2.append(np.concatenate(parts2))
    faces = np.concatenate([hf, of + hv.shape[0]]) if with_object else hf
    img = rng.uniform(-0.5, 0.5, size=(2, batch, 3, height, width)).astype(np.float32)
    jit = np.ones((2, batch, 3, height, width), dtype=np.float32)
    for i in range(2):
        for b in range(batch):
            band = rng.integers(0, max(int(0.1 * min(height, width)), 1), size=4)
            jit[i, b, :, : band[0], :] = 0
            jit[i, b, :, height - band[1]:, :] = 0
            jit[i, b, :, :, : band[2]] = 0
            jit[i, b, :, :, width - band[3]:] = 0
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
    return dict(
        verts1=t(np.stack(verts1)), verts2=t(np.stack(verts2)),
        faces=t(np.tile(faces[None], (batch, 1, 1)), torch.int64),
        K=camera_intrinsics(batch, width, height, device),
        image_ref=t(img[0]), image=t(img[1]), jitter_mask_ref=t(jit[0]), jitter_mask=t(jit[1]),
        hand_ignore_faces=HAND_IGNORE_FACES,
    )


def mano_model(seed=0, device="cpu"):
    """Synthetic parameter set with MANO's shapes (the licensed MANO_RIGHT.pkl is not available): the hand
    template, 16 joints on five three-link chains, softmax skinning weights, small shape / pose blend shapes,
    an orthonormal PCA basis and a non-zero mean pose.  Returns a dict of float32 tensors + ``tip_ids`` /
    ``faces`` (the 1538 non-closing faces)."""
    rng = np.random.default_rng(seed + 100)
    v, f = hand_template()
    v = v.astype(np.float64)
    # joints: root at the wrist end, five chains fanning out along +y
    joints = [np.array([0.0, -0.08, 0.0])]
    for finger in range(5):
        x = (finger - 2) * 0.018
        for link in range(3):
            joints.append(np.array([x * (1 + 0.2 * link), -0.02 + 0.035 * link, 0.0]))
    J = np.stack(joints)  # [16,3]
    d2 = ((v[:, None, :] - J[None]) ** 2).sum(-1)
    w = np.exp(-d2 / (0.02 ** 2))
    w = w / w.sum(1, keepdims=True)
    # J_regressor: rows that reproduce J from the template as well as a sparse, positive regressor can
    jr = np.exp(-d2.T / (0.012 ** 2))
    jr = jr / jr.sum(1, keepdims=True)
    q, _ = np.linalg.qr(rng.normal(size=(45, 45)))
    model = dict(
        v_template=v, shapedirs=rng.normal(size=(778, 3, 10)) * 2e-3, posedirs=rng.normal(size=(778, 3, 135)) * 2e-4,
        j_regressor=jr, weights=w, hands_components=q, hands_mean=rng.normal(size=45) * 0.15)
    out = {k: torch.from_numpy(np.ascontiguousarray(a)).float().to(device) for k, a in model.items()}
    out["tip_ids"] = (745, 317, 444, 556, 673)
    out["faces"] = torch.from_numpy(f[:1538].copy()).to(device)
    return out
