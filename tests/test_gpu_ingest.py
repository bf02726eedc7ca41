"""Image ingestion: the three image sources of `Matcher` (matcher.py:167-400) through the
batched pipeline (threaded decode, 8-bit upload, device-side /255, batched SuperPoint) must give exactly what
the reference's one-image-at-a-time loop gives."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import t  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _grey_u8(seed, H=240, W=320, shift=(0, 0)):
    from simplesfm_b200 import synthetic
    im = synthetic.textured_image(seed, H, W, shift=shift)[0, 0].numpy()
    return np.clip(np.rint(im * 255), 0, 255).astype(np.uint8)


def _same_frames(fa, fb):
    assert len(fa) == len(fb)
    for a, b in zip(fa, fb):
        assert torch.equal(a.keypoints_poses, b.keypoints_poses)
        assert torch.equal(a.scores, b.scores)
        assert torch.equal(a.descriptors, b.descriptors)
        assert a.image.shape == b.image.shape and torch.equal(a.image, b.image)


def _same_tables(ta, tb):
    assert set(ta) == set(tb)
    for k in ta:
        assert np.array_equal(ta[k], tb[k]), k


def test_u8_to_unit_f32_matches_numpy(dev):
    from simplesfm_b200._lib import check, lib, ptr, stream_ptr
    for n in (256 * 7 + 5, 16, 3):
        x = torch.arange(n, dtype=torch.int64).remainder(256).to(torch.uint8)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        xd = x.to(dev)
        check(lib().sfm_u8_to_unit_f32(ptr(xd), ptr(out), n, stream_ptr(dev)))
        assert np.array_equal(out.cpu().numpy(), x.numpy().astype("float32") / 255)


@pytest.mark.parametrize("masked", [False, True])
def test_simple_sfm_dataset_batched_equals_per_image(tmp_path, sp_weights, sg_weights, dev, masked):
    from PIL import Image
    from simplesfm_b200.matchers import Matcher
    frames = []
    for i in range(5):
        Image.fromarray(_grey_u8(20 + i, shift=(i, -i))).save(tmp_path / f"im_{i}.png")
        item = {"file_path": f"./im_{i}.png"}
        if masked:
            m = np.zeros((240, 320), np.uint8)
            m[20:200, 30 + 10 * i:300] = 255
            Image.fromarray(m).save(tmp_path / f"mask_{i}.png")
            item["mask_path"] = f"./mask_{i}.png"
        frames.append(item)
    (tmp_path / "views_data.json").write_text(json.dumps({"frames": frames}))
    kw = dict(matcher_type="super_glue", super_glue_weights_path=sg_weights, match_threshold=0.2,
              sinkhorn_iterations=10, super_glue_batch=2, max_keypoints=256, device=dev)
    ref = Matcher(sp_weights, 4, 0.005, extract_batch=1, **kw)
    new = Matcher(sp_weights, 4, 0.005, extract_batch=4, decode_threads=3, **kw)
    mask_name = "mask_path" if masked else None
    fa, ta, na = ref.match_simple_sfm_dataset(str(tmp_path), dataset_mask_name=mask_name)
    fb, tb, nb = new.match_simple_sfm_dataset(str(tmp_path), dataset_mask_name=mask_name)
    assert na == nb == {i + 1: f"im_{i}.png" for i in range(5)}
    assert all(len(f.keypoints_poses) > 20 for f in fa)
    _same_frames(fa, fb)
    _same_tables(ta, tb)
    if masked:   # every keypoint lies inside its mask rectangle
        for i, f in enumerate(fb):
            kp = f.keypoints_poses.cpu().numpy()
            assert (kp[:, 0] >= 30 + 10 * i).all() and (kp[:, 0] < 300).all()
            assert (kp[:, 1] >= 20).all() and (kp[:, 1] < 200).all()


class _FakeStream:
    """Minimal stand-in for utils/video_streamer.py VideoStreamer: next_frame() -> (f32 image, status, path)."""

    def __init__(self, images):
        self.images, self.i = images, 0

    def next_frame(self):
        if self.i >= len(self.images):
            return None, False, ""
        im = self.images[self.i]
        self.i += 1
        return im, True, f"/video/frame_{self.i:05d}.png"


def test_video_stream_batched_equals_per_image(sp_weights, dev):
    from simplesfm_b200.matchers import Matcher
    imgs = [_grey_u8(40 + i).astype("float32") / 255 for i in range(5)]
    masks = [(np.arange(320)[None, :] > 40 * i).astype("float32") * np.ones((240, 1), "float32") for i in range(5)]
    out = []
    for eb in (1, 3):   # 5 frames in batches of 3: one full, one ragged
        m = Matcher(sp_weights, 4, 0.005, matcher_type="simple", match_threshold=0.7, device=dev, extract_batch=eb)
        out.append(m.match_video_stream(_FakeStream(imgs), vs_mask=_FakeStream(masks)))
    (fa, ta, na), (fb, tb, nb) = out
    assert na == nb and na[1] == "frame_00001.png" and len(na) == 5
    _same_frames(fa, fb)
    _same_tables(ta, tb)


def test_unbounded_keypoints_and_mixed_shapes(sp_weights, dev):
    """max_keypoints = -1 (one count read-back per batch) and a shape change in the middle of the stream."""
    from simplesfm_b200.ingest import FrameExtractor
    from simplesfm_b200.models.superpoint import SuperPoint
    sp = SuperPoint(sp_weights, max_keypoints=-1, device=dev)
    items = [(_grey_u8(60 + i), None) for i in range(3)] + [(_grey_u8(70 + i, 160, 240), None) for i in range(2)]
    got = FrameExtractor(sp, batch_size=4).extract(items)
    assert len(got) == 5
    for (im, _), e in zip(items, got):
        x = torch.from_numpy(im.astype("float32") / 255)[None, None].to(dev)
        r = sp({"image": x})
        assert torch.equal(r["keypoints"], e.keypoints) and torch.equal(r["scores"], e.scores)
        assert torch.equal(r["descriptors"], e.descriptors)
        assert torch.equal(e.image, x[0])


def test_config3_miniature_images_to_colmap_db(tmp_path, sp_weights, sg_weights_outdoor, dev):
    """BASELINE config 3 in miniature (generate_sparse_colmap defaults: outdoor weights, masks, batch 4, 20 Sinkhorn
    iterations; `simple_sfm_dataset_utils.py:82-139`): image + mask files -> Matcher.match_simple_sfm_dataset ->
    ColmapDatabaseWriter, checked against the CPU oracle run on the same decoded pixels and the oracle's blob
    restatement of `colmap_bd_utils.py:291-372`."""
    import sqlite3
    import sfm_oracle as O
    from PIL import Image
    from simplesfm_b200.colmap_db import ColmapDatabaseWriter
    from simplesfm_b200.matchers import Matcher
    n, H, W = 4, 160, 240
    frames_meta, imgs, masks = [], [], []
    for i in range(n):
        im = _grey_u8(90, H, W, shift=(2 * i, -3 * i))
        m = np.zeros((H, W), np.uint8)
        m[8:H - 8, 8 + 4 * i:W - 8] = 255
        Image.fromarray(im).save(tmp_path / f"view_{i}.png")
        Image.fromarray(m).save(tmp_path / f"view_{i}_mask.png")
        frames_meta.append({"file_path": f"./view_{i}.png", "mask_path": f"./view_{i}_mask.png"})
        imgs.append(im)
        masks.append(m)
    (tmp_path / "views_data.json").write_text(json.dumps({"frames": frames_meta}))
    thr, mthr = 0.005, 0.2    # synthetic weights: the defaults 1e-4 / 0.9 would keep everything / nothing
    m = Matcher(sp_weights, 4, thr, matcher_type="super_glue", super_glue_weights_path=sg_weights_outdoor,
                match_threshold=mthr, sinkhorn_iterations=20, super_glue_batch=4, max_keypoints=200, device=dev)
    frames, table, names = m.match_simple_sfm_dataset(str(tmp_path), dataset_mask_name="mask_path")
    db = ColmapDatabaseWriter(str(tmp_path / "colmap"), images_folder_path=str(tmp_path), camera_size=(W, H))
    db.replace_images_data(names)
    db.replace_keypoints(names, frames)
    db.replace_and_verificate_matches(table, names, run_verification=False)

    # oracle on the same pixels
    ofr = []
    for im, mk in zip(imgs, masks):
        x = torch.from_numpy(im.astype("float32") / 255)[None, None]
        o = O.superpoint_forward(sp_weights, x, torch.from_numpy(mk > 0)[None, None], 4, thr, 200, 4)
        ofr.append((o["descriptors"], o["keypoints"], o["scores"], x[0]))
    otab = O.superglue_match_table(sg_weights_outdoor, ofr, O.pair_list(list(range(1, n + 1))), 4, 20, mthr, "train")

    con = sqlite3.connect(str(tmp_path / "colmap" / "database.db"))
    assert [r[0] for r in con.execute("SELECT name FROM images ORDER BY image_id")] == [f"view_{i}.png" for i in range(n)]
    total = 0
    for i in range(n):
        kp = db.read_keypoints(i + 1)
        okp = ofr[i][1].numpy()
        assert kp.shape == (len(okp), 4) and (kp[:, 2] == 1).all() and (kp[:, 3] == 0).all()
        assert {tuple(r) for r in kp[:, :2].tolist()} == {tuple(r) for r in okp.tolist()}
    assert set(table) == set(otab)
    for (a, b), ref in otab.items():
        got = db.read_matches(int(a), int(b))
        ka, kb_ = db.read_keypoints(int(a))[:, :2], db.read_keypoints(int(b))[:, :2]
        oa, ob = ofr[a - 1][1].numpy(), ofr[b - 1][1].numpy()
        assert {(tuple(ka[i]), tuple(kb_[j])) for i, j in got.tolist()} == \
               {(tuple(oa[i]), tuple(ob[j])) for i, j in ref.tolist()}, (a, b)
        total += len(ref)
    assert total > 20
    assert open(tmp_path / "colmap" / "match.txt").read().splitlines()[0] == "view_0.png view_1.png"


def test_device_pair_list_equals_itertools(dev):
    """`sfm_pair_list`: any range of the lexicographic (i < j) schedule, generated on the device,
    equals `itertools.combinations` (matcher.py:143-144); 2 000 images = 1 999 000 pairs exercise the closed form."""
    import itertools
    from simplesfm_b200._lib import check, lib, ptr, stream_ptr
    L = lib()
    for n in (2, 3, 7, 50, 333, 2000):
        total = n * (n - 1) // 2
        a, b = np.triu_indices(n, k=1)
        for first, count in ((0, total), (total // 3, total - total // 3), (total - 1, 1), (5 % total, 0)):
            i0 = torch.full((max(count, 1),), -1, dtype=torch.int32, device=dev)
            i1 = torch.full((max(count, 1),), -1, dtype=torch.int32, device=dev)
            check(L.sfm_pair_list(n, first, count, ptr(i0), ptr(i1), stream_ptr(dev)))
            assert np.array_equal(i0.cpu().numpy()[:count], a[first:first + count])
            assert np.array_equal(i1.cpu().numpy()[:count], b[first:first + count])
    assert list(zip(*np.triu_indices(7, k=1))) == list(itertools.combinations(range(7), 2))
    with pytest.raises(Exception):
        check(L.sfm_pair_list(5, 8, 5, ptr(i0), ptr(i1), stream_ptr(dev)))


def test_matcher_on_image_sizes_not_multiple_of_8(tmp_path, sp_weights, sg_weights, dev):
    """The whole path (PNG files -> batched ingestion -> SuperPoint fp32 -> SuperGlue fp32 -> tables) on 150 x 203
    images: batched extraction equals the per-image loop, the frames equal the CPU oracle's SuperPoint on the same
    pixels (keypoint sets bit-exact), and the match table equals the oracle's for the extracted frames."""
    import sfm_oracle as O
    from PIL import Image
    from simplesfm_b200.matchers import Matcher
    H, W = 150, 203
    items = []
    for i in range(3):
        Image.fromarray(_grey_u8(40 + i, H + 2, W + 5, shift=(2 * i, i))[:H, :W]).save(tmp_path / f"odd_{i}.png")
        items.append({"file_path": f"./odd_{i}.png"})
    (tmp_path / "views_data.json").write_text(json.dumps({"frames": items}))
    kw = dict(matcher_type="super_glue", super_glue_weights_path=sg_weights, match_threshold=0.2,
              sinkhorn_iterations=10, super_glue_batch=2, max_keypoints=300, device=dev)
    fa, ta, _ = Matcher(sp_weights, 4, 0.005, extract_batch=1, **kw).match_simple_sfm_dataset(str(tmp_path))
    fb, tb, _ = Matcher(sp_weights, 4, 0.005, extract_batch=3, **kw).match_simple_sfm_dataset(str(tmp_path))
    _same_frames(fa, fb)
    _same_tables(ta, tb)
    oframes = []
    for i, f in enumerate(fb):
        assert tuple(f.image.shape) == (1, H, W)
        px = np.asarray(Image.open(tmp_path / f"odd_{i}.png")).astype(np.float32) / 255
        ref = O.superpoint_forward(sp_weights, torch.from_numpy(px)[None, None], None, 4, 0.005, 300, 4)
        got = {(int(x), int(y)) for x, y in f.keypoints_poses.cpu().numpy()}
        want = {(int(x), int(y)) for x, y in ref["keypoints"].numpy()}
        assert len(want) > 50 and got == want
        assert max(x for x, _ in got) < W // 8 * 8 and max(y for _, y in got) < H // 8 * 8
        oframes.append((f.descriptors.cpu(), f.keypoints_poses.cpu(), f.scores.cpu(), torch.zeros(1, H, W)))
    otab = O.superglue_match_table(sg_weights, oframes, O.pair_list([1, 2, 3]), 2, 10, 0.2, "train")
    assert set(otab) == set(tb)
    n = 0
    for k in otab:
        assert np.array_equal(np.asarray(otab[k]), tb[k]), k
        n += len(tb[k])
    assert n > 0
