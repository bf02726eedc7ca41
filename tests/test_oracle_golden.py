"""CPU: the oracle restatement must reproduce the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  This is the pin that makes the oracle trustworthy."""
import numpy as np
import pytest
import torch

import sfm_oracle as O
from conftest import load_golden, t, canon_keypoints


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def test_superpoint_stages(sp_weights):
    g = load_golden("sp_small_stages")
    img = t(g["image"])
    feat = O.sp_encoder(sp_weights, img)
    close(feat, g["feat"], 1e-5, 1e-5)
    logits = O.sp_detector_logits(sp_weights, t(g["feat"]))
    close(logits, g["logits"], 1e-5, 1e-5)
    smap = O.sp_score_map(t(g["logits"]))
    close(smap, g["score_map"], 1e-6, 1e-8)
    assert torch.equal(O.simple_nms(t(g["score_map"]), 4), t(g["nms"]))
    close(O.sp_descriptor_map(sp_weights, t(g["feat"])), g["dmap"], 1e-5, 1e-6)


@pytest.mark.parametrize("tag,mk,masked", [("all", -1, False), ("top200", 200, False), ("masked", -1, True)])
def test_superpoint_forward(sp_weights, tag, mk, masked):
    g = load_golden(f"sp_small_{tag}")
    mask = t(g["mask"]) if masked else None
    out = O.superpoint_forward(sp_weights, t(g["image"]), mask, 4, 0.005, mk, 4)
    rk, rs, ro = canon_keypoints(g["keypoints"], g["scores"])
    ok, os_, oo = canon_keypoints(out["keypoints"].numpy(), out["scores"].numpy())
    assert rk.shape == ok.shape
    assert np.array_equal(rk, ok)                     # keypoint indices bit-exact
    close(os_, rs, 1e-6, 0)
    close(out["descriptors"].numpy()[:, oo], g["descriptors"][:, ro], 1e-4, 1e-6)
    assert np.all(np.diff(out["scores"].numpy()) <= 0)


def test_superpoint_480x640(sp_weights):
    from simplesfm_b200 import synthetic
    g = load_golden("sp_480x640_top1024")
    out = O.superpoint_forward(sp_weights, synthetic.textured_image(1), None, 4, 0.005, 1024, 4)
    assert np.array_equal(out["keypoints"].numpy(), g["keypoints"])   # scores are distinct here
    close(out["scores"], g["scores"], 1e-6, 0)
    close(out["descriptors"][:, :128], g["descriptors_first128"], 1e-4, 1e-6)


def test_superpoint_functions():
    g = load_golden("sp_functions")
    for r in (4, 2, 0):
        assert torch.equal(O.simple_nms(t(g["nms_in"]), r), t(g[f"nms_r{r}"]))
    # plateau: all interior plateau pixels survive
    assert (O.simple_nms(t(g["nms_in"]), 4)[0, 14:26, 14:26] == 0.7).all()
    out = O.sample_descriptors(t(g["samp_kpts"]), t(g["samp_dmap"]), 8)
    close(out, g["samp_out"], 1e-5, 1e-6)
    assert O.default_align_corners() is False         # torch 2.11 -> '1' -> False (hazard H2)


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("iters", [20, 100])
def test_superglue_forward(sg_weights, mode, iters):
    g = load_golden(f"sg_b2_{mode}_it{iters}")
    data = {k: t(g[k]) for k in ("descriptors0", "descriptors1", "keypoints0", "keypoints1", "scores0", "scores1")}
    data["image0"] = data["image1"] = torch.zeros(2, 1, int(g["image_hw"][0]), int(g["image_hw"][1]))
    tr = {}
    out = O.superglue_forward(sg_weights, data, iters, 0.2, mode, trace=tr)
    close(tr["scores"], g["scores"], 2e-4, 2e-4)
    close(tr["Z"], g["Z"], 2e-4, 2e-4)
    assert np.array_equal(out["matches0"].numpy(), g["matches0"])     # mutual matches bit-exact
    assert np.array_equal(out["matches1"].numpy(), g["matches1"])
    assert out["matches0"].dtype == torch.int64
    close(out["matching_scores0"], g["matching_scores0"], 1e-4, 1e-6)
    close(out["matching_scores1"], g["matching_scores1"], 1e-4, 1e-6)
    if iters == 20:
        close(tr["kenc"][0] - data["descriptors0"], g["kenc_0"], 1e-4, 1e-5)
        close(tr["kenc"][1] - data["descriptors1"], g["kenc_1"], 1e-4, 1e-5)
        # layer deltas: delta_l = desc_after_l - desc_before_l
        prev0 = tr["kenc"][0]
        close(tr["layers"][0][0] - prev0, g["delta_l0_0"], 1e-3, 1e-4)
        close(tr["layers"][17][1] - tr["layers"][16][1], g["delta_l17_1"], 1e-3, 1e-4)


def test_superglue_empty_side(sg_weights):
    g = load_golden("sg_empty_side")
    data = {"keypoints0": torch.zeros(1, 8, 2), "keypoints1": torch.zeros(1, 0, 2)}
    out = O.superglue_forward(sg_weights, data)
    assert str(out["matches0"].dtype) == str(g["dtypes"][0]) == "torch.int32"
    for k in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
        assert np.array_equal(out[k].numpy(), g[k])


def test_superglue_functions():
    g = load_golden("sg_functions")
    sc = t(g["ot_scores"])
    close(O.log_optimal_transport(sc, torch.tensor(1.0), 20), g["ot_it20"], 1e-5, 1e-5)
    close(O.log_optimal_transport(sc, torch.tensor(1.0), 100), g["ot_it100"], 1e-5, 1e-5)
    close(O.log_optimal_transport(sc, torch.tensor(2.5), 5), g["ot_alpha2_it5"], 1e-5, 1e-5)
    close(O.attention(t(g["att_q"]), t(g["att_k"]), t(g["att_v"])), g["att_out"], 1e-5, 1e-6)
    close(O.normalize_keypoints(t(g["nk_in"]), (2, 1, 480, 640)), g["nk_out"], 1e-6, 1e-7)
    close(O.normalize_keypoints(t(g["nk_in"]), (2, 1, 640, 480)), g["nk_out_tall"], 1e-6, 1e-7)


def _frames(g, n):
    return [(t(g[f"desc_{i}"]), t(g[f"kpts_{i}"]), t(g[f"scores_{i}"]), torch.zeros(1, 480, 640)) for i in range(n)]


@pytest.mark.parametrize("batch", [4, 1])
def test_matcher_tables(sg_weights_outdoor, batch):
    g = load_golden("matcher_4frames")
    frames = _frames(g, 4)
    pairs = O.pair_list([1, 2, 3, 4])
    table = O.superglue_match_table(sg_weights_outdoor, frames, pairs, batch, 20, 0.2)
    assert len(table) == 6
    for (a, b), tab in table.items():
        ref = g[f"sg_b{batch}_{a}_{b}"]
        assert tab.dtype == np.uint32 and ref.dtype == np.uint32
        assert np.array_equal(tab, ref), (a, b)
    filt = O.pair_list([1, 2, 3, 4], lambda p: abs(p[0] - p[1]) <= 2)
    assert np.array_equal(filt, g["sg_filtered_keys"])


def test_simple_matcher():
    g = load_golden("matcher_4frames")
    frames = _frames(g, 4)
    out = O.simple_match_two_way(frames[0][0], frames[1][0], 0.7)
    close(out, g["simple_two_way_np"], 1e-5, 1e-6)
    close(out, g["simple_two_way_cuda"], 1e-5, 1e-6)
    for a in range(1, 5):
        for b in range(a + 1, 5):
            o = O.simple_match_two_way(frames[a - 1][0], frames[b - 1][0], 0.7)
            assert np.array_equal(o.T[:, :2].astype(np.uint32), g[f"simple_{a}_{b}"])
    assert O.simple_match_two_way(frames[0][0][:, :0], frames[1][0], 0.7).shape == (3, 0)
    with pytest.raises(ValueError):
        O.simple_match_two_way(frames[0][0], frames[1][0], -1.0)


def test_matcher_images_e2e(sp_weights, sg_weights_outdoor):
    g = load_golden("matcher_images_e2e")
    frames = []
    for i in range(3):
        im = t(g[f"image_{i}"])
        r = O.superpoint_forward(sp_weights, im, None, 4, 0.005, -1, 4)
        assert np.array_equal(canon_keypoints(r["keypoints"].numpy(), r["scores"].numpy())[0],
                              canon_keypoints(g[f"kpts_{i}"], g[f"scores_{i}"])[0])
        frames.append((r["descriptors"], r["keypoints"], r["scores"], im[0]))
    table = O.superglue_match_table(sg_weights_outdoor, frames, O.pair_list([1, 2, 3]), 2, 20, 0.2)
    for (a, b), tab in table.items():
        assert np.array_equal(tab, g[f"table_{a}_{b}"])


def test_colmap_payloads():
    assert O.colmap_pair_id(3, 7) == 3 * 2147483647 + 7 == O.colmap_pair_id(7, 3)
    r, c, blob = O.colmap_keypoint_blob(np.array([[1.0, 2.0], [3.0, 4.0]], np.float32))
    assert (r, c) == (2, 4)
    assert np.array_equal(np.frombuffer(blob, np.float32).reshape(2, 4), [[1, 2, 1, 0], [3, 4, 1, 0]])
    pid, r, c, blob = O.colmap_match_blob(5, 2, np.array([[1, 9], [4, 8]], np.uint32))
    assert np.array_equal(np.frombuffer(blob, np.uint32).reshape(2, 2), [[9, 1], [8, 4]])


def test_module_functions_golden():
    """Oracle pieces behind the module-level functions the reference exports (superglue.py:142-148 with
    arbitrary couplings; superpoint.py:27-39) against tests/golden/module_functions.npz."""
    g = load_golden("module_functions")
    Z, mu, nu = t(g["lsi_Z"]), t(g["lsi_log_mu"]), t(g["lsi_log_nu"])
    u, v = O.log_sinkhorn(Z, mu, nu, 25)
    torch.testing.assert_close(Z + u.unsqueeze(2) + v.unsqueeze(1), t(g["lsi_out_25"]), rtol=1e-5, atol=1e-5)
    u0, v0 = O.log_sinkhorn(Z, mu, nu, 0)
    assert torch.equal(Z + u0.unsqueeze(2) + v0.unsqueeze(1), t(g["lsi_out_0"]))
    kp, sc = t(g["rb_in_kpts"]), t(g["rb_in_scores"])
    b, h, w = int(g["rb_border"]), int(g["rb_h"]), int(g["rb_w"])
    keep = (kp[:, 0] >= b) & (kp[:, 0] < h - b) & (kp[:, 1] >= b) & (kp[:, 1] < w - b)
    assert torch.equal(kp[keep], t(g["rb_out_kpts"])) and torch.equal(sc[keep], t(g["rb_out_scores"]))
    top = torch.topk(sc, int(g["tk_k"])).values
    assert torch.equal(torch.sort(top).values, torch.sort(t(g["tk_out_scores"])).values)
