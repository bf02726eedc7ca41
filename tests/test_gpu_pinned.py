"""The BENCHMARKED configuration pinned directly against the CPU oracle ("next round" item 1).

`bench.py` runs SuperGlue on 5-pair Matcher chunks at 2048 keypoints, train-mode BatchNorm, 20 Sinkhorn
iterations, match threshold 0.4 (reference defaults: matcher.py:33-41, superglue.py:56-57 run in train mode
because the reference never calls .eval()).  These tests run exactly that chunk shape through the C ABI and
compare it with `oracle/sfm_oracle.py` (itself pinned on the reference goldens in tests/test_oracle_golden.py):

* fp32 path: `matches0/1` bit-exact, or every differing row is a numerical near-tie of the oracle's own top-2
  assignment scores / sits on the match threshold (the reference's float32 arithmetic decides those by rounding);
* bf16 path (stated precision): match-set agreement with the ORACLE >= 99.5 % (BASELINE.json north_star);
* the module-level functions the reference exports (`log_sinkhorn_iterations`, `remove_borders`,
  `top_k_keypoints`, `arange_like`) against the goldens; `Matcher.match_datasets` with a real image-folder dataset.
"""
import os

import numpy as np
import pytest
import torch

import sfm_oracle as O
from conftest import load_golden, t

pytestmark = pytest.mark.gpu

BENCH = dict(B=5, K=2048, iters=20, thr=0.4, bn="train")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _chunk(B, K, seed):
    """One Matcher chunk of the bench scene family: pairs (1,2) .. (1,B+1) of a feature scene."""
    from simplesfm_b200 import synthetic
    fr = synthetic.feature_scene(B + 1, K, seed=seed)
    idx0, idx1 = [0] * B, list(range(1, B + 1))
    return {"descriptors0": torch.stack([fr[i][0] for i in idx0]), "descriptors1": torch.stack([fr[i][0] for i in idx1]),
            "keypoints0": torch.stack([fr[i][1] for i in idx0]), "keypoints1": torch.stack([fr[i][1] for i in idx1]),
            "scores0": torch.stack([fr[i][2] for i in idx0]), "scores1": torch.stack([fr[i][2] for i in idx1]),
            "image0": torch.zeros(B, 1, 480, 640), "image1": torch.zeros(B, 1, 480, 640)}


def _set_agreement(a, b):
    """|A & B| / |A | B| over the (pair, i, j) match sets of two matches0 arrays."""
    a, b = np.asarray(a), np.asarray(b)
    sa = {(p, i, int(j)) for p, i in zip(*np.nonzero(a > -1)) for j in [a[p, i]]}
    sb = {(p, i, int(j)) for p, i in zip(*np.nonzero(b > -1)) for j in [b[p, i]]}
    return len(sa & sb) / max(len(sa | sb), 1), len(sb)


@pytest.fixture(scope="module")
def bench_chunk_oracle(sg_weights):
    """The oracle on one bench-shaped chunk (about 4 s on the GPU box's host cores), shared by both precisions."""
    torch.set_num_threads(os.cpu_count() or 1)
    data = _chunk(BENCH["B"], BENCH["K"], seed=0)
    tr = {}
    with torch.no_grad():
        ref = O.superglue_forward(sg_weights, data, BENCH["iters"], BENCH["thr"], BENCH["bn"], trace=tr)
    return data, ref, tr


def test_fp32_bench_chunk_bit_exact_vs_oracle(sg_weights, dev, bench_chunk_oracle):
    """(a) fp32 path, 5 pairs x 2048 keypoints, train BN, 20 iterations, thr 0.4 against the oracle."""
    from simplesfm_b200.models.superglue import SuperGlue
    data, ref, tr = bench_chunk_oracle
    m = SuperGlue(sg_weights, sinkhorn_iterations=BENCH["iters"], match_threshold=BENCH["thr"], bn_mode=BENCH["bn"],
                  precision="fp32", device=dev)
    out = m({k: v.to(dev) for k, v in data.items()})
    a0, b0 = out["matches0"].cpu().numpy(), ref["matches0"].numpy()
    a1, b1 = out["matches1"].cpu().numpy(), ref["matches1"].numpy()
    n_ref = int((b0 > -1).sum())
    assert n_ref > 2000, n_ref
    bad0 = np.argwhere(a0 != b0)
    Z = tr["Z"][:, :-1, :-1]
    for bi, i in bad0:
        # a differing row must be a near-tie of the oracle's top-2 assignment scores, or its score sits on the
        # match threshold, or the column it points at is itself such a near-tie (mutual check)
        top2 = torch.topk(Z[bi, i], 2).values
        j = int(torch.argmax(Z[bi, i]))
        ctop2 = torch.topk(Z[bi, :, j], 2).values
        near = float(top2[0] - top2[1]) < 1e-4 or float(ctop2[0] - ctop2[1]) < 1e-4 \
            or abs(float(top2[0].exp()) - BENCH["thr"]) < 1e-4
        assert near, (int(bi), int(i), top2.tolist())
    assert len(bad0) <= 2, f"{len(bad0)} of {a0.size} rows differ"
    if len(bad0) == 0:
        assert np.array_equal(a1, b1)
    s_gpu, s_ref = out["matching_scores0"].cpu(), ref["matching_scores0"]
    torch.testing.assert_close(s_gpu, s_ref, rtol=1e-3, atol=1e-5)
    print(f"fp32 bench chunk vs oracle: {n_ref} matches, {len(bad0)} differing rows")


def test_bf16_bench_chunk_agreement_vs_oracle(sg_weights, dev, bench_chunk_oracle):
    """(b) bf16 (stated precision) path on the same chunk: >= 99.5 % match-set agreement with the ORACLE."""
    from simplesfm_b200.models.superglue import SuperGlue
    data, ref, _ = bench_chunk_oracle
    m = SuperGlue(sg_weights, sinkhorn_iterations=BENCH["iters"], match_threshold=BENCH["thr"], bn_mode=BENCH["bn"],
                  precision="bf16", device=dev)
    out = m({k: v.to(dev) for k, v in data.items()})
    agree, n_ref = _set_agreement(out["matches0"].cpu().numpy(), ref["matches0"].numpy())
    print(f"bf16 bench chunk vs oracle: agreement {agree:.4f} over {n_ref} oracle matches")
    assert n_ref > 2000
    assert agree >= 0.995


@pytest.mark.parametrize("mode,iters,thr,bar", [("train", 20, 0.4, 0.995), ("train", 20, 0.2, 0.985), ("eval", 100, 0.2, 0.985)])
def test_bf16_1024kp_agreement_vs_oracle(sg_weights, dev, mode, iters, thr, bar):
    """bf16 path against the oracle at 1024 keypoints, 4 pairs: both BatchNorm modes, both iteration counts.
    At the Matcher's threshold (0.4) the stated 99.5 % bar holds.  At SuperGlue's own default 0.2 this synthetic scene
    crowds the threshold (a fifth of its matches score within 0.05 of it) and the measured agreement is 99.0 - 99.7 %
    depending on the rounding pattern of a kernel version; what is asserted there is that EVERY disagreement is a
    borderline decision (score next to the threshold, or a near-tie of the oracle's top-2 assignment scores)."""
    from simplesfm_b200.models.superglue import SuperGlue
    torch.set_num_threads(os.cpu_count() or 1)
    data = _chunk(4, 1024, seed=3)
    tr = {}
    with torch.no_grad():
        ref = O.superglue_forward(sg_weights, data, iters, thr, mode, trace=tr)
    m = SuperGlue(sg_weights, sinkhorn_iterations=iters, match_threshold=thr, bn_mode=mode, precision="bf16", device=dev)
    out = m({k: v.to(dev) for k, v in data.items()})
    a, b = out["matches0"].cpu().numpy(), ref["matches0"].numpy()
    agree, n_ref = _set_agreement(a, b)
    print(f"bf16 vs oracle [{mode}, {iters} it, thr {thr}]: agreement {agree:.4f} over {n_ref} oracle matches")
    assert n_ref > 500
    assert agree >= bar
    s_gpu, s_ref = out["matching_scores0"].cpu().numpy(), ref["matching_scores0"].numpy()
    Z = tr["Z"][:, :-1, :-1].numpy()
    for bi, i in np.argwhere(a != b):
        top2 = np.sort(Z[bi, i])[-2:]
        ctop2 = np.sort(Z[bi, :, int(np.argmax(Z[bi, i]))])[-2:]
        borderline = abs(s_ref[bi, i] - thr) < 0.05 or abs(s_gpu[bi, i] - thr) < 0.05 or \
            (top2[1] - top2[0]) < 0.1 or (ctop2[1] - ctop2[0]) < 0.1
        assert borderline, (int(bi), int(i), float(s_ref[bi, i]), float(s_gpu[bi, i]), top2.tolist())


# ---------------------------------------------------------------------------------------------- module functions
def test_log_sinkhorn_iterations_vs_golden(dev):
    """Module-level `log_sinkhorn_iterations` (superglue.py:142-148) on ARBITRARY couplings and marginals
    against the reference's output, and on the dustbin couplings `log_optimal_transport` builds."""
    from simplesfm_b200.models import superglue as SG
    g = load_golden("module_functions")
    Z, mu, nu = t(g["lsi_Z"]).to(dev), t(g["lsi_log_mu"]).to(dev), t(g["lsi_log_nu"]).to(dev)
    torch.testing.assert_close(SG.log_sinkhorn_iterations(Z, mu, nu, 25).cpu(), t(g["lsi_out_25"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(SG.log_sinkhorn_iterations(Z, mu, nu, 0).cpu(), t(g["lsi_out_0"]), rtol=0, atol=0)
    # oracle restatement on the same inputs
    u, v = O.log_sinkhorn(t(g["lsi_Z"]), t(g["lsi_log_mu"]), t(g["lsi_log_nu"]), 25)
    torch.testing.assert_close(t(g["lsi_Z"]) + u.unsqueeze(2) + v.unsqueeze(1), t(g["lsi_out_25"]), rtol=1e-5, atol=1e-5)
    # the couplings of log_optimal_transport (superglue.py:151-171): Z_ot = log_sinkhorn(couplings) - norm
    f = load_golden("sg_functions")
    scores = t(f["ot_scores"]).to(dev)
    b, m, n = scores.shape
    norm = -float(np.log(m + n))
    for alpha, iters, key in ((1.0, 20, "ot_it20"), (1.0, 100, "ot_it100"), (2.5, 5, "ot_alpha2_it5")):
        couplings = torch.full((b, m + 1, n + 1), alpha, device=dev)
        couplings[:, :m, :n] = scores
        log_mu = torch.cat([torch.full((m,), norm), torch.tensor([np.log(n) + norm], dtype=torch.float32)]).to(dev)
        log_nu = torch.cat([torch.full((n,), norm), torch.tensor([np.log(m) + norm], dtype=torch.float32)]).to(dev)
        Zo = SG.log_sinkhorn_iterations(couplings, log_mu[None].expand(b, -1), log_nu[None].expand(b, -1), iters)
        torch.testing.assert_close((Zo - norm).cpu(), t(f[key]), rtol=1e-4, atol=1e-4)


def test_superpoint_module_functions_vs_golden(dev):
    """`remove_borders`, `top_k_keypoints` (superpoint.py:27-39) and `arange_like` (superglue.py:174-175)
    against the reference's outputs."""
    from simplesfm_b200.models import superglue as SG
    from simplesfm_b200.models import superpoint as SP
    g = load_golden("module_functions")
    kp, sc = t(g["rb_in_kpts"]).to(dev), t(g["rb_in_scores"]).to(dev)
    k2, s2 = SP.remove_borders(kp, sc, int(g["rb_border"]), int(g["rb_h"]), int(g["rb_w"]))
    assert torch.equal(k2.cpu(), t(g["rb_out_kpts"])) and torch.equal(s2.cpu(), t(g["rb_out_scores"]))
    k3, s3 = SP.top_k_keypoints(kp, sc, int(g["tk_k"]))
    # torch.topk leaves the order of equal scores open: compare as (score, keypoint) sets
    got = sorted(zip(s3.cpu().tolist(), map(tuple, k3.cpu().tolist())))
    want = sorted(zip(t(g["tk_out_scores"]).tolist(), map(tuple, t(g["tk_out_kpts"]).tolist())))
    assert got == want
    k4, s4 = SP.top_k_keypoints(kp, sc, len(kp) + 5)        # k >= len: passthrough
    assert torch.equal(k4, kp) and torch.equal(s4, sc)
    x = torch.zeros(3, 7, 5, device=dev)
    assert torch.equal(SG.arange_like(x, 0).cpu(), t(g["arange_d0"]))
    assert torch.equal(SG.arange_like(x, 2).cpu(), t(g["arange_d2"]))
    assert torch.equal(SG.arange_like(torch.zeros(2, 7, device=dev), 1).cpu(), t(load_golden("sg_functions")["arange"]))


# ---------------------------------------------------------------------------------------------- match_datasets
class ImageFolders(torch.utils.data.Dataset):
    """Stand-in with the interface of the reference's `ImageFoldersDataset` (utils/datasets.py:16-87), which is
    not on the GPU box: `folders_dict` {'rgb': dir[, 'mask': dir]} of *.jpg files, `transforms_dict` per type;
    items are dicts {type: transformed image, 'path_<type>': path, 'image_name': stem}.  `order` fixes the item
    order (the reference iterates in `glob` order, which the golden recorded)."""

    def __init__(self, folders_dict, transforms_dict, order):
        self.transforms = transforms_dict
        self.meta = {}
        for name in order:
            for key, path in folders_dict.items():
                p = os.path.join(path, name)
                if os.path.exists(p):
                    self.meta.setdefault(name.split(".")[0], {})[key] = p

    def __len__(self):
        return len(self.meta)

    def __getitem__(self, idx):
        from PIL import Image
        name = list(self.meta.keys())[idx]
        out = {"image_name": name}
        for kind, path in self.meta[name].items():
            with Image.open(path) as img:
                out[kind] = self.transforms[kind](img)
            out["path_" + kind] = path
        return out


def _to_grey_tensor(img):
    return torch.from_numpy(np.asarray(img.convert("L"), dtype=np.float32) / 255.0)[None]


@pytest.mark.parametrize("masked", [False, True])
def test_match_datasets_vs_reference_golden(tmp_path, sp_weights, sg_weights, dev, masked):
    """`Matcher.match_datasets` (matcher.py:167-247) over two image-folder datasets (3 + 2 JPEG files, optional
    masks) against what the REFERENCE Matcher returned for the same files through its own `ImageFoldersDataset`
    (tests/golden/make_golden.py::golden_module_functions): names, split index, image shapes, keypoint sets
    and match tables."""
    from simplesfm_b200.matchers import Matcher
    g = load_golden("module_functions")
    tag = "md_masked" if masked else "md"
    names_ref = [str(x) for x in g[f"{tag}_names"]]
    sets = []
    for d in ("a", "b"):
        folders, tf = {"rgb": str(tmp_path / d / "rgb")}, {"rgb": _to_grey_tensor}
        if masked:
            folders["mask"], tf["mask"] = str(tmp_path / d / "mask"), _to_grey_tensor
        for kind in ("rgb", "mask"):
            (tmp_path / d / kind).mkdir(parents=True)
        for key in g:
            if key.startswith("jpg_") and key.split("_")[2].startswith(d):
                _, kind, stem = key.split("_")
                (tmp_path / d / kind / (stem + ".jpg")).write_bytes(g[key].tobytes())
        sets.append(ImageFolders(folders, tf, [n for n in names_ref if n.startswith(d)]))
    m = Matcher(sp_weights, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg_weights, match_threshold=0.2,
                sinkhorn_iterations=20, super_glue_batch=2, device=dev)
    frames, table, names, split = m.match_datasets(sets, None, image_name_path_shift=1)
    assert split == g[f"{tag}_split"].tolist()
    assert [names[i + 1] for i in range(len(names))] == names_ref
    assert len(frames) == 5 and len(table) == 10
    for i, f in enumerate(frames):
        assert list(f.image.shape) == g[f"{tag}_image_shape_{i}"].tolist()
        got = {tuple(x) for x in f.keypoints_poses.cpu().numpy().tolist()}
        want = {tuple(x) for x in g[f"{tag}_kpts_{i}"].tolist()}
        assert got == want, (i, len(got ^ want))
    # tables through keypoint coordinates (ranks of float32-equal scores are left open by torch.argsort; a rank
    # swap can also move the chunk-minimum truncation by one keypoint, hence a set-agreement bar)
    inter = union = 0
    for (a, b), tab in table.items():
        assert tab.dtype == np.uint32 and isinstance(a, np.int64)
        ka, kb = frames[a - 1].keypoints_poses.cpu().numpy(), frames[b - 1].keypoints_poses.cpu().numpy()
        ra, rb = g[f"{tag}_kpts_{a - 1}"], g[f"{tag}_kpts_{b - 1}"]
        got = {(tuple(ka[i]), tuple(kb[j])) for i, j in tab.tolist()}
        want = {(tuple(ra[i]), tuple(rb[j])) for i, j in g[f"{tag}_table_{int(a)}_{int(b)}"].tolist()}
        inter += len(got & want)
        union += len(got | want)
    print(f"match_datasets[{tag}] vs reference: {inter}/{union} matches agree")
    assert union > 500 and inter / union >= 0.99, (inter, union)
