"""CPU oracle for the SuperPoint -> SuperGlue matching path of palsol/SimpleSfm.

TEST INFRASTRUCTURE ONLY.  This module is a functional CPU restatement of the
reference algorithm (the reference is eager PyTorch; its arithmetic lives in the
un-vendored third-party dependency `torch`, requirement unpinned in
reference `requirements.txt:22`; run here with torch 2.11.0+cu128 on CPU).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker / CPU baseline.
The product (`simplesfm_b200`) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the UNMODIFIED reference modules run in
the build container: `tests/golden/make_golden.py` imports `/root/reference`,
runs it on seeded synthetic weights/images and stores per-stage tensors in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here
against those files.

Every function cites the reference file:line it restates (paths relative to the
reference root).  State dicts use the reference's key layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# SuperPoint (simple_sfm/models/superpoint.py)
# --------------------------------------------------------------------------------------

_SP_ENCODER = (("conv1a", "conv1b", True), ("conv2a", "conv2b", True),
               ("conv3a", "conv3b", True), ("conv4a", "conv4b", False))


def sp_encoder(w: Dict[str, Tensor], image: Tensor) -> Tensor:
    """Shared VGG encoder: (1,1,H,W) -> (1,128,H/8,W/8).  superpoint.py:112-122."""
    x = image
    for a, b, pool in _SP_ENCODER:
        x = F.relu(F.conv2d(x, w[a + ".weight"], w[a + ".bias"], padding=1))
        x = F.relu(F.conv2d(x, w[b + ".weight"], w[b + ".bias"], padding=1))
        if pool:
            x = F.max_pool2d(x, 2, 2)
    return x


def sp_detector_logits(w: Dict[str, Tensor], feat: Tensor) -> Tensor:
    """convPa+ReLU+convPb: (1,128,h,w) -> (1,65,h,w).  superpoint.py:125-126."""
    c = F.relu(F.conv2d(feat, w["convPa.weight"], w["convPa.bias"], padding=1))
    return F.conv2d(c, w["convPb.weight"], w["convPb.bias"])


def sp_score_map(logits: Tensor) -> Tensor:
    """softmax over 65 channels, drop the dustbin, 8x8 pixel shuffle -> (1,8h,8w).
    superpoint.py:127-130."""
    p = torch.softmax(logits, 1)[:, :64]
    b, _, h, w = p.shape
    p = p.permute(0, 2, 3, 1).reshape(b, h, w, 8, 8)
    return p.permute(0, 1, 3, 2, 4).reshape(b, h * 8, w * 8)


def simple_nms(scores: Tensor, radius: int) -> Tensor:
    """Iterated (1 + 2 rounds) max-pool NMS with float equality.  superpoint.py:9-24."""
    assert radius >= 0
    k = 2 * radius + 1

    def wmax(t):
        return F.max_pool2d(t, kernel_size=k, stride=1, padding=radius)

    zero = torch.zeros_like(scores)
    keep = scores == wmax(scores)
    for _ in range(2):
        near = wmax(keep.float()) > 0
        rest = torch.where(near, zero, scores)
        keep = keep | ((rest == wmax(rest)) & ~near)
    return torch.where(keep, scores, zero)


def sp_select_keypoints(nms_scores: Tensor, threshold: float, border: int,
                        max_keypoints: int, mask: Optional[Tensor] = None
                        ) -> Tuple[Tensor, Tensor]:
    """mask -> threshold -> nonzero (row-major) -> border removal -> top-k ->
    (x, y) float keypoints sorted by score descending.
    superpoint.py:133-154,165-171.  `nms_scores` is (H, W); mask is bool (H, W).

    Tie order: torch.topk / argsort leave equal-score order unspecified
    (superpoint.py:38,165); this oracle fixes it to (score desc, row-major pixel
    index asc), which is one valid output of the reference.
    """
    s = nms_scores.clone()
    H, W = s.shape
    if mask is not None:
        s[~mask] = 0                                              # :133-134
    rc = torch.nonzero(s > threshold)                             # :137-139, row-major
    val = s[rc[:, 0], rc[:, 1]]                                   # :140
    ok = ((rc[:, 0] >= border) & (rc[:, 0] < H - border) &        # :27-32
          (rc[:, 1] >= border) & (rc[:, 1] < W - border))
    rc, val = rc[ok], val[ok]
    order = torch.sort(val, descending=True, stable=True).indices  # :165 (+ tie rule)
    if max_keypoints >= 0 and max_keypoints < len(rc):            # :35-39,148-151
        order = order[:max_keypoints]
    rc, val = rc[order], val[order]
    return torch.flip(rc, [1]).float(), val                       # :154


def sp_descriptor_map(w: Dict[str, Tensor], feat: Tensor) -> Tensor:
    """convDa+ReLU+convDb+L2 normalise over channels.  superpoint.py:157-159."""
    c = F.relu(F.conv2d(feat, w["convDa.weight"], w["convDa.bias"], padding=1))
    d = F.conv2d(c, w["convDb.weight"], w["convDb.bias"])
    return F.normalize(d, p=2, dim=1)


def default_align_corners() -> bool:
    """The reference keys align_corners on the 3rd character of torch.__version__
    (superpoint.py:49): '2.11.0' -> '1' -> not > 2 -> argument omitted -> False."""
    return int(torch.__version__[2]) > 2


def sample_descriptors(kpts_xy: Tensor, dmap: Tensor, s: int = 8,
                       align_corners: Optional[bool] = None) -> Tensor:
    """Bilinear sampling of the (1,256,h,w) descriptor map at (K,2) keypoints, then
    L2 normalise -> (256,K).  superpoint.py:42-54.  Written out tap by tap (not via
    grid_sample) so it is an independent restatement; zero padding outside."""
    if align_corners is None:
        align_corners = default_align_corners()
    _, c, h, w = dmap.shape
    kp = kpts_xy.to(torch.float32)
    g = kp - s / 2 + 0.5
    g = g / torch.tensor([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5], dtype=torch.float32)
    g = g * 2 - 1
    if align_corners:
        px = (g[:, 0] + 1) / 2 * (w - 1)
        py = (g[:, 1] + 1) / 2 * (h - 1)
    else:
        px = ((g[:, 0] + 1) * w - 1) / 2
        py = ((g[:, 1] + 1) * h - 1) / 2
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    fx, fy = px - x0, py - y0
    x0, y0 = x0.long(), y0.long()
    out = torch.zeros(c, kp.shape[0], dtype=torch.float32)
    m = dmap[0]
    for dy, dx, wt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                       (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xi, yi = x0 + dx, y0 + dy
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        v = m[:, yi.clamp(0, h - 1), xi.clamp(0, w - 1)]
        out += v * (wt * ok.float())[None]
    return F.normalize(out, p=2, dim=0)


def superpoint_forward(w: Dict[str, Tensor], image: Tensor, mask: Optional[Tensor] = None,
                       nms_radius: int = 4, keypoint_threshold: float = 0.005,
                       max_keypoints: int = -1, remove_borders: int = 4,
                       align_corners: Optional[bool] = None) -> Dict[str, Tensor]:
    """SuperPoint.forward for one image (1,1,H,W); mask (1,1,H,W) bool.
    superpoint.py:109-171."""
    feat = sp_encoder(w, image)
    smap = simple_nms(sp_score_map(sp_detector_logits(w, feat)), nms_radius)[0]
    m = None if mask is None else mask[0, 0]
    kpts, scores = sp_select_keypoints(smap, keypoint_threshold, remove_borders,
                                       max_keypoints, m)
    dmap = sp_descriptor_map(w, feat)
    desc = sample_descriptors(kpts, dmap, 8, align_corners)
    return {"keypoints": kpts, "scores": scores, "descriptors": desc}


# --------------------------------------------------------------------------------------
# SuperGlue (simple_sfm/models/superglue.py)
# --------------------------------------------------------------------------------------

BN_EPS = 1e-5  # torch BatchNorm1d default, superglue.py:57


def _bn(x: Tensor, w: Dict[str, Tensor], prefix: str, bn_mode: str) -> Tensor:
    """BatchNorm1d on (B,C,N).  The reference never calls .eval(), so
    as run it normalises with batch statistics over (B,N) (biased variance);
    bn_mode='eval' uses the stored running statistics.  superglue.py:56-57."""
    g, b = w[prefix + ".weight"], w[prefix + ".bias"]
    if bn_mode == "train":
        mean = x.mean(dim=(0, 2), keepdim=True)
        var = x.var(dim=(0, 2), unbiased=False, keepdim=True)
    elif bn_mode == "eval":
        mean = w[prefix + ".running_mean"][None, :, None]
        var = w[prefix + ".running_var"][None, :, None]
    else:
        raise ValueError(bn_mode)
    return (x - mean) / torch.sqrt(var + BN_EPS) * g[None, :, None] + b[None, :, None]


def _mlp(x: Tensor, w: Dict[str, Tensor], prefix: str, n_layers: int, bn_mode: str) -> Tensor:
    """Conv1d(k=1) [+BN+ReLU] stack; sequential indices 0,3,6.. are convs, 1,4,7.. BN.
    superglue.py:48-59."""
    for i in range(n_layers):
        x = F.conv1d(x, w[f"{prefix}.{3 * i}.weight"], w[f"{prefix}.{3 * i}.bias"])
        if i < n_layers - 1:
            x = F.relu(_bn(x, w, f"{prefix}.{3 * i + 1}", bn_mode))
    return x


def normalize_keypoints(kpts: Tensor, image_shape: Sequence[int]) -> Tensor:
    """(kpts - [W/2,H/2]) / (0.7*max(W,H)).  superglue.py:62-69."""
    h, w = image_shape[-2], image_shape[-1]
    size = torch.tensor([[float(w), float(h)]], dtype=kpts.dtype)
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center[:, None, :]) / scaling[:, None, :]


def keypoint_encoder(w: Dict[str, Tensor], kpts_n: Tensor, scores: Tensor, bn_mode: str) -> Tensor:
    """MLP([3,32,64,128,256,256]) on cat(x,y,score): (B,K,2),(B,K) -> (B,256,K).
    superglue.py:72-82."""
    x = torch.cat([kpts_n.transpose(1, 2), scores.unsqueeze(1)], dim=1)
    return _mlp(x, w, "kenc.encoder", 5, bn_mode)


def attention(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q^T k / sqrt(d)) v per head on (B,d,h,N) tensors.  superglue.py:85-89."""
    d = q.shape[1]
    s = torch.einsum("bdhn,bdhm->bhnm", q, k) / d ** 0.5
    return torch.einsum("bhnm,bdhm->bdhn", torch.softmax(s, dim=-1), v)


def gnn_layer_delta(w: Dict[str, Tensor], layer: int, x: Tensor, src: Tensor, bn_mode: str) -> Tensor:
    """AttentionalPropagation.forward: mlp(cat[x, merge(attn(q(x),k(src),v(src)))]).
    Heads are interleaved: channel c = d*4 + h.  superglue.py:92-120."""
    p = f"gnn.layers.{layer}"
    b = x.shape[0]
    q = F.conv1d(x, w[p + ".attn.proj.0.weight"], w[p + ".attn.proj.0.bias"]).view(b, 64, 4, -1)
    k = F.conv1d(src, w[p + ".attn.proj.1.weight"], w[p + ".attn.proj.1.bias"]).view(b, 64, 4, -1)
    v = F.conv1d(src, w[p + ".attn.proj.2.weight"], w[p + ".attn.proj.2.bias"]).view(b, 64, 4, -1)
    msg = attention(q, k, v).contiguous().view(b, 256, -1)
    msg = F.conv1d(msg, w[p + ".attn.merge.weight"], w[p + ".attn.merge.bias"])
    return _mlp(torch.cat([x, msg], dim=1), w, p + ".mlp", 2, bn_mode)


def gnn(w: Dict[str, Tensor], d0: Tensor, d1: Tensor, bn_mode: str,
        n_layers: int = 18, trace: Optional[List] = None) -> Tuple[Tensor, Tensor]:
    """18 alternating self/cross layers; both sides read the pre-update other side.
    superglue.py:123-139."""
    for i in range(n_layers):
        cross = (i % 2 == 1)
        s0, s1 = (d1, d0) if cross else (d0, d1)
        e0 = gnn_layer_delta(w, i, d0, s0, bn_mode)
        e1 = gnn_layer_delta(w, i, d1, s1, bn_mode)
        d0, d1 = d0 + e0, d1 + e1
        if trace is not None:
            trace.append((d0.clone(), d1.clone()))
    return d0, d1


def log_sinkhorn(Z: Tensor, log_mu: Tensor, log_nu: Tensor, iters: int) -> Tuple[Tensor, Tensor]:
    """u, v after `iters` log-space Sinkhorn sweeps from u=v=0.  superglue.py:142-148
    (returns the potentials; the caller adds them to Z)."""
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return u, v


def log_optimal_transport(scores: Tensor, alpha: Tensor, iters: int) -> Tensor:
    """(B,m,n) scores -> (B,m+1,n+1) log assignment with dustbins.  superglue.py:151-171."""
    b, m, n = scores.shape
    a = alpha.to(scores).reshape(1, 1, 1)
    Z = torch.cat([torch.cat([scores, a.expand(b, m, 1)], -1),
                   torch.cat([a.expand(b, 1, n), a.expand(b, 1, 1)], -1)], 1)
    # the reference computes these constants in float32 tensors (superglue.py:165-167)
    one = scores.new_tensor(1)
    nrm = -((m * one) + (n * one)).log()
    log_mu = torch.cat([nrm.expand(m), (n * one).log()[None] + nrm])[None].expand(b, -1)
    log_nu = torch.cat([nrm.expand(n), (m * one).log()[None] + nrm])[None].expand(b, -1)
    u, v = log_sinkhorn(Z, log_mu, log_nu, iters)
    return Z + u.unsqueeze(2) + v.unsqueeze(1) - nrm


def mutual_matches(Zfull: Tensor, match_threshold: float) -> Dict[str, Tensor]:
    """Mutual nearest neighbour + threshold on Z[:, :-1, :-1].  superglue.py:273-283.
    max() ties resolve to the first index (CPU behaviour)."""
    S = Zfull[:, :-1, :-1]
    max0, max1 = S.max(2), S.max(1)
    i0, i1 = max0.indices, max1.indices
    ar0 = torch.arange(i0.shape[1])[None]
    ar1 = torch.arange(i1.shape[1])[None]
    mutual0 = ar0 == i1.gather(1, i0)
    mutual1 = ar1 == i0.gather(1, i1)
    zero = S.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values.exp(), zero)
    ms1 = torch.where(mutual1, ms0.gather(1, i1), zero)
    valid0 = mutual0 & (ms0 > match_threshold)
    valid1 = mutual1 & valid0.gather(1, i1)
    return {"matches0": torch.where(valid0, i0, i0.new_tensor(-1)),
            "matches1": torch.where(valid1, i1, i1.new_tensor(-1)),
            "matching_scores0": ms0, "matching_scores1": ms1}


def superglue_forward(w: Dict[str, Tensor], data: Dict[str, Tensor], sinkhorn_iterations: int = 100,
                      match_threshold: float = 0.2, bn_mode: str = "train",
                      trace: Optional[Dict] = None) -> Dict[str, Tensor]:
    """SuperGlue.forward on a batch of pairs.  superglue.py:235-290.
    data: descriptors0/1 (B,256,K), keypoints0/1 (B,K,2), scores0/1 (B,K),
    image0/1 (only .shape[-2:] is read, superglue.py:250-251)."""
    k0, k1 = data["keypoints0"], data["keypoints1"]
    if k0.shape[1] == 0 or k1.shape[1] == 0:                       # :240-247
        s0, s1 = k0.shape[:-1], k1.shape[:-1]
        return {"matches0": torch.full(s0, -1, dtype=torch.int32),
                "matches1": torch.full(s1, -1, dtype=torch.int32),
                "matching_scores0": torch.zeros(s0), "matching_scores1": torch.zeros(s1)}
    n0 = normalize_keypoints(k0, data["image0"].shape)
    n1 = normalize_keypoints(k1, data["image1"].shape)
    d0 = data["descriptors0"] + keypoint_encoder(w, n0, data["scores0"], bn_mode)
    d1 = data["descriptors1"] + keypoint_encoder(w, n1, data["scores1"], bn_mode)
    if trace is not None:
        trace["kenc"] = (d0.clone(), d1.clone())
        trace["layers"] = []
    d0, d1 = gnn(w, d0, d1, bn_mode, trace=None if trace is None else trace["layers"])
    m0 = F.conv1d(d0, w["final_proj.weight"], w["final_proj.bias"])
    m1 = F.conv1d(d1, w["final_proj.weight"], w["final_proj.bias"])
    scores = torch.einsum("bdn,bdm->bnm", m0, m1) / 256 ** 0.5
    if trace is not None:
        trace["scores"] = scores.clone()
    Z = log_optimal_transport(scores, w["bin_score"], sinkhorn_iterations)
    if trace is not None:
        trace["Z"] = Z.clone()
    return mutual_matches(Z, match_threshold)


# --------------------------------------------------------------------------------------
# Matcher / SimpleMatcher (simple_sfm/matchers/*.py)
# --------------------------------------------------------------------------------------

def pair_list(image_ids: Sequence[int], pair_filter=None) -> np.ndarray:
    """All id pairs in lexicographic order, filtered.  matcher.py:143-144."""
    out = [(a, b) for i, a in enumerate(image_ids) for b in list(image_ids)[i + 1:]]
    arr = np.array(out).reshape(-1, 2)
    if pair_filter is not None:
        arr = np.array([p for p in arr if pair_filter(p)]).reshape(-1, 2)
    return arr


def superglue_match_table(w: Dict[str, Tensor], frames: List[Tuple[Tensor, Tensor, Tensor, Tensor]],
                          pairs: np.ndarray, batch: int, sinkhorn_iterations: int,
                          match_threshold: float, bn_mode: str = "train"
                          ) -> Dict[Tuple[int, int], np.ndarray]:
    """Chunk the pair list, truncate every side to the chunk-minimum keypoint count,
    run SuperGlue, emit (L,2) uint32 [index0, index1] tables keyed by 1-based ids.
    matcher.py:60-113.  frames[i] = (descriptors (256,K), keypoints (K,2), scores (K,), image)."""
    table = {}
    for c in range(0, len(pairs), batch):
        chunk = pairs[c:c + batch] - 1
        f0 = [frames[i] for i in chunk[:, 0]]
        f1 = [frames[i] for i in chunk[:, 1]]
        k0 = min(len(f[1]) for f in f0)
        k1 = min(len(f[1]) for f in f1)
        data = {"descriptors0": torch.stack([f[0][:, :k0] for f in f0]),
                "descriptors1": torch.stack([f[0][:, :k1] for f in f1]),
                "keypoints0": torch.stack([f[1][:k0] for f in f0]),
                "keypoints1": torch.stack([f[1][:k1] for f in f1]),
                "scores0": torch.stack([f[2][:k0] for f in f0]),
                "scores1": torch.stack([f[2][:k1] for f in f1]),
                "image0": torch.stack([f[3] for f in f0]),
                "image1": torch.stack([f[3] for f in f1])}
        res = superglue_forward(w, data, sinkhorn_iterations, match_threshold, bn_mode)
        for i, (a, b) in enumerate(chunk):
            m = res["matches0"][i]
            idx = torch.nonzero(m > -1)[:, 0]
            table[(int(a) + 1, int(b) + 1)] = torch.stack([idx, m[idx].long()], 1).numpy().astype(np.uint32)
    return table


def simple_match_two_way(desc1: Tensor, desc2: Tensor, nn_thresh: float) -> np.ndarray:
    """Mutual nearest neighbour on L2 distance of unit descriptors -> (3,L) [i, j, dist].
    simple_matcher.py:11-64 / 66-114."""
    if desc1.shape[1] == 0 or desc2.shape[1] == 0:
        return np.zeros((3, 0))
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    d = torch.sqrt(2 - 2 * torch.clamp(desc1.t() @ desc2, -1, 1))
    j = torch.argmin(d, dim=1)
    i_back = torch.argmin(d, dim=0)
    best = d[torch.arange(d.shape[0]), j]
    keep = (best < nn_thresh) & (torch.arange(len(j)) == i_back[j])
    out = np.zeros((3, int(keep.sum())))
    out[0] = torch.nonzero(keep)[:, 0].numpy()
    out[1] = j[keep].numpy()
    out[2] = best[keep].numpy()
    return out


# --------------------------------------------------------------------------------------
# COLMAP database payloads (simple_sfm/colmap_utils/colmap_bd_utils.py)
# --------------------------------------------------------------------------------------

def colmap_pair_id(id1: int, id2: int) -> int:
    """colmap_bd_utils.py:86-91."""
    a, b = (id2, id1) if id1 > id2 else (id1, id2)
    return a * 2147483647 + b


def colmap_keypoint_blob(kpts_xy: np.ndarray) -> Tuple[int, int, bytes]:
    """float32 [x, y, 1, 0] rows.  colmap_bd_utils.py:324-331 (tostring -> tobytes)."""
    k = np.concatenate([kpts_xy, np.ones((len(kpts_xy), 1)), np.zeros((len(kpts_xy), 1))], 1).astype(np.float32)
    return k.shape[0], k.shape[1], k.tobytes()


def colmap_match_blob(id1: int, id2: int, matches: np.ndarray) -> Tuple[int, int, int, bytes]:
    """uint32 (L,2), columns swapped when id1 > id2.  colmap_bd_utils.py:354-367."""
    m = matches[:, [1, 0]] if id1 > id2 else matches
    m = np.ascontiguousarray(m)
    return colmap_pair_id(id1, id2), m.shape[0], m.shape[1], m.tobytes()


# --------------------------------------------------------------------------------------
# Post-match consumer: sparse depth from a COLMAP sparse model
# (simple_sfm/dataset/simple_sfm_dataset_utils.py:147-261)
# --------------------------------------------------------------------------------------

def read_points3d_bin(path: str) -> List[Tuple[np.ndarray, float, np.ndarray]]:
    """points3D.bin -> [(xyz, error, image_ids)] in file order.  colmap_utils/read_write_colmap_data.py:336-363."""
    import struct
    out = []
    with open(path, "rb") as fid:
        (n,) = struct.unpack("<Q", fid.read(8))
        for _ in range(n):
            rec = struct.unpack("<QdddBBBd", fid.read(43))
            (tl,) = struct.unpack("<Q", fid.read(8))
            tr = struct.unpack("<" + "ii" * tl, fid.read(8 * tl))
            out.append((np.array(rec[1:4]), float(rec[7]), np.array(tr[0::2], dtype=np.int64)))
    return out


def simple_sfm_json_cameras(views: dict):
    """cameras/camera_multiple.py:212-276: per view sorted by id -> (w2c 4x4, K 3x3, [W, H], id, file_path).
    K[1][2] is c_x as in the reference (:252)."""
    cams = []
    for view in sorted(views["frames"], key=lambda item: item["id"]):
        k = view["intrinsic"]
        c2w = np.array(view["transform_matrix"])
        c2w[2, :] *= 1
        c2w = c2w[[1, 0, 2, 3], :]
        c2w[0:3, 1] *= -1
        c2w[0:3, 2] *= -1
        K = np.array([[k["f_x"], 0, k["c_x"]], [0, k["f_y"], k["c_x"]], [0, 0, 1]], dtype=np.float64)
        cams.append((np.linalg.inv(c2w), K, [k["original_resolution_x"], k["original_resolution_y"]],
                     int(view["id"]), view["file_path"]))
    return cams


def sparse_depth_for_camera(points, w2c: np.ndarray, K: np.ndarray, size, camera_id: int, scale: float = 1,
                            error_threshold: float = 2.0):
    """One iteration of the camera loop (simple_sfm_dataset_utils.py:195-233) -> (sparse_depth (1,W,H) f64,
    mask (1,W,H) bool, points_bounds).  Projection: utils/coord_conversion.py:61-97, 138-159."""
    W, H = int(size[0] // scale), int(size[1] // scale)
    sel = [xyz for xyz, err, image_ids in points if error_threshold > err and camera_id + 1 in image_ids]   # :204-209
    depth = np.zeros((W, H, 1), np.float64) - 1
    bounds = [-1, 1]
    if sel:
        p = np.array(sel, np.float64)
        cam = p @ w2c[:3, :3].T + w2c[:3, 3]                                  # world_to_cam
        z = cam[:, 2:]
        den = np.where(z >= 0, np.maximum(z, 1e-6), np.minimum(z, -1e-6))      # cam_to_film
        film = cam[:, :2] / den
        pix = film * np.diagonal(K[:2, :2]) + K[:2, 2]                         # film_to_pixel
        px = np.concatenate([pix, z], 1) * np.array([[W, H, 1]], np.float64)   # :211
        m = (px[:, 0] >= 0) & (px[:, 0] <= W) & (px[:, 1] >= 0) & (px[:, 1] <= H) & (px[:, 2] >= 0)
        px = px[m]
        xy = px[:, :2].astype(np.int64)                                        # .long() truncation
        if len(px):
            if (xy[:, 0] >= W).any() or (xy[:, 1] >= H).any():
                raise IndexError("projected point on the x == W / y == H edge")
            for (x, y), d in zip(xy, px[:, 2]):                                # indexed write: the last point wins
                depth[x, y, 0] = d
            bounds = [px[:, 2].min(), px[:, 2].max()]
    return depth.transpose(2, 0, 1), (depth != -1).transpose(2, 0, 1), bounds
