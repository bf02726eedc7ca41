"""Layer-0 hoisting: with eval-mode BatchNorm the keypoint encoder and GNN layer 0 (a
self layer, superglue.py:131-139) depend on one image alone, so the bf16 path computes them once per image
(`sfm_superglue_hoist_layer0`) and matches pairs from the stored residual streams (`sfm_superglue_forward_hoisted`).
The results must be BIT-IDENTICAL to the plain forward pass."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_hoisted_forward_equals_plain_forward(sg_weights, dev):
    from simplesfm_b200 import synthetic
    from simplesfm_b200.models.superglue import SuperGlue
    n, K, B = 6, 384, 5
    fr = synthetic.feature_scene(n, K, seed=12)
    desc = torch.stack([f[0] for f in fr]).to(dev)
    kpts = torch.stack([f[1] for f in fr]).to(dev)
    scores = torch.stack([f[2] for f in fr]).to(dev)
    i0 = torch.tensor([0, 0, 1, 2, 4], dtype=torch.int32, device=dev)
    i1 = torch.tensor([1, 5, 3, 3, 5], dtype=torch.int32, device=dev)
    m = SuperGlue(sg_weights, sinkhorn_iterations=20, match_threshold=0.2, bn_mode="eval", precision="bf16", device=dev)
    assert m.can_hoist()
    plain = [x.clone() for x in m.match_records(desc, kpts, scores, i0, desc, kpts, scores, i1, B, K, K, (480, 640),
                                                (480, 640))[:4]]
    store = m.hoist_layer0(desc, kpts, scores, n, K, (480, 640))
    assert store.shape == (n, K, 256) and torch.isfinite(store).all()
    hoisted = m.match_hoisted(store, i0, store, i1, B, K, K)[:4]
    assert int((plain[0] > -1).sum()) > 100
    for a, b in zip(plain, hoisted):
        assert torch.equal(a, b)
    # train-mode BatchNorm (the reference default) has batch statistics over the pair batch: no hoisting there
    assert not SuperGlue(sg_weights, bn_mode="train", precision="bf16", device=dev).can_hoist()
    assert not SuperGlue(sg_weights, bn_mode="eval", precision="fp32", device=dev).can_hoist()


def test_matcher_tables_identical_with_and_without_hoisting(sp_weights, sg_weights, dev):
    from simplesfm_b200 import synthetic
    from simplesfm_b200.matchers import Matcher
    from simplesfm_b200.matchers.matcher import Frame
    fr = synthetic.feature_scene(7, 512, seed=4)
    img = torch.zeros(1, 480, 640, device=dev)
    frames = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in fr]
    names = {i + 1: str(i) for i in range(7)}
    tabs = {}
    for hoist in (False, True):
        m = Matcher(sp_weights, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg_weights,
                    match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=3, bn_mode="eval", precision="bf16",
                    device=dev, hoist_layer0=hoist)
        tabs[hoist] = m.match_frames(frames, names, lambda p: abs(p[0] - p[1]) != 3)[1]
    assert len(tabs[True]) == len(tabs[False]) == 17
    assert sum(len(v) for v in tabs[False].values()) > 500
    for k in tabs[False]:
        assert np.array_equal(tabs[False][k], tabs[True][k])
    # ragged keypoint counts (chunk-minimum truncation) silently take the plain path
    frames[2] = Frame(fr[2][0][:, :400].to(dev), fr[2][1][:400].to(dev), fr[2][2][:400].to(dev), img)
    m = Matcher(sp_weights, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg_weights,
                match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=3, bn_mode="eval", precision="bf16",
                device=dev)
    t = m.match_frames(frames, names, None)[1]
    assert len(t) == 21 and max(int(v[:, 0].max()) for (a, b), v in t.items() if a == 3 and len(v)) < 400
