"""`generate_sparse_depth_from_colmap` (reference simple_sfm/dataset/simple_sfm_dataset_utils.py:147-261).

Golden: tests/golden/sparse_depth.npz = the input files (views_data.json + a COLMAP sparse model written by the
reference's own writer) and what the UNMODIFIED reference function produced for them at scale 1 / threshold 2 and
scale 2 / threshold 3 (tests/golden/make_golden.py::golden_sparse_depth).
CPU tests: the oracle restatement and the host-side track flattening against the golden.
GPU tests: the CUDA path (through the C ABI) against the golden, and against the oracle on a larger random model
with many pixel collisions."""
import json
import os

import numpy as np
import pytest
import torch

import sfm_oracle as O
from conftest import load_golden

CASES = (("s1", 1, 2.0), ("s2", 2, 3.0))


def _write_dataset(tmp_path, g):
    root = tmp_path / "ds"
    os.makedirs(root / "colmap" / "sparse")
    (root / "views_data.json").write_bytes(g["views_json"].tobytes())
    for name in ("cameras.bin", "images.bin", "points3D.bin"):
        (root / "colmap" / "sparse" / name).write_bytes(g["model_" + name.replace(".", "_")].tobytes())
    return root


def test_oracle_matches_reference_golden(tmp_path):
    g = load_golden("sparse_depth")
    root = _write_dataset(tmp_path, g)
    points = O.read_points3d_bin(str(root / "colmap" / "sparse" / "points3D.bin"))
    cams = O.simple_sfm_json_cameras(json.load(open(root / "views_data.json")))
    assert len(points) == 400 and len(cams) == 6
    for tag, scale, thr in CASES:
        hits = 0
        for w2c, K, size, cid, name in cams:
            depth, mask, _ = O.sparse_depth_for_camera(points, w2c, K, size, cid, scale, thr)
            assert np.array_equal(mask, g[f"{tag}_mask_{cid}"])
            np.testing.assert_allclose(depth, g[f"{tag}_depth_{cid}"], rtol=1e-12, atol=1e-12)
            hits += int(mask.sum())
        assert hits > 100


def test_host_entries_match_oracle_selection(tmp_path):
    """The flattened (camera, point) entries equal the reference's nested selection loop (:204-209), including
    repeated image ids in a track and the strict error comparison."""
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    g = load_golden("sparse_depth")
    root = _write_dataset(tmp_path, g)
    xyz, err, tp, ti = U.read_points3d_tracks(root / "colmap" / "sparse")
    points = O.read_points3d_bin(str(root / "colmap" / "sparse" / "points3D.bin"))
    assert np.array_equal(xyz, np.array([p[0] for p in points]))
    assert np.array_equal(err, np.array([p[1] for p in points]))
    views = U.read_simple_sfm_views(root / "views_data.json")
    cams = O.simple_sfm_json_cameras(json.load(open(root / "views_data.json")))
    assert views["ids"] == [c[3] for c in cams]
    for k, c in enumerate(cams):
        assert np.array_equal(views["extrinsics"][k].reshape(3, 4), c[0][:3])
        assert views["intrinsics"][k].tolist() == [c[1][0, 0], c[1][1, 1], c[1][0, 2], c[1][1, 2]]
    for thr in (2.0, 3.0, float(err[5])):                        # the last: `thr > error` must exclude equality
        ecam, ept = U.camera_point_entries(err, tp, ti, views["ids"], thr)
        for k, cid in enumerate(views["ids"]):
            want = [j for j, (_, e, ids) in enumerate(points) if thr > e and cid + 1 in ids]
            assert ept[ecam == k].tolist() == want


def test_points3d_text_reader(tmp_path):
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    d = tmp_path / "m"
    os.makedirs(d)
    (d / "cameras.txt").write_text("# none\n")
    (d / "images.txt").write_text("# none\n")
    (d / "points3D.txt").write_text("# header\n7 0.5 -1.25 3 10 20 30 0.75 2 11 3 4\n9 1 2 3 0 0 0 2.5 1 0\n")
    xyz, err, tp, ti = U.read_points3d_tracks(d)
    assert xyz.tolist() == [[0.5, -1.25, 3.0], [1.0, 2.0, 3.0]] and err.tolist() == [0.75, 2.5]
    assert tp.tolist() == [0, 0, 1] and ti.tolist() == [2, 3, 1]
    with pytest.raises(FileNotFoundError):
        U.read_points3d_tracks(tmp_path)


@pytest.mark.gpu
def test_cuda_path_matches_reference_golden(tmp_path):
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    g = load_golden("sparse_depth")
    for tag, scale, thr in CASES:
        root = _write_dataset(tmp_path / tag, g)
        U.generate_sparse_depth_from_colmap(str(root), scale=scale, error_threshold=thr)
        out = json.load(open(root / "views_data.json"))
        assert [f["sparse_depth_path"] for f in out["frames"]] == g[f"{tag}_paths"].tolist()
        for f in out["frames"]:
            d = np.load(root / f["sparse_depth_path"], allow_pickle=True).item()
            assert d["sparse_depth"].dtype == np.float64 and d["sparse_depth_mask"].dtype == np.bool_
            assert np.array_equal(d["sparse_depth_mask"], g[f"{tag}_mask_{f['id']}"])
            np.testing.assert_allclose(d["sparse_depth"], g[f"{tag}_depth_{f['id']}"], rtol=1e-12, atol=1e-12)
            name = f["file_path"].split("/")[-1] + "_" + str(f["id"])
            assert (root / "sparse_depth" / "sparse_depth_vis" / f"sparse_depth_vis_{name}.jpg").exists()


@pytest.mark.gpu
def test_cuda_path_matches_oracle_with_collisions():
    """40 000 points into 12 cameras of 48 x 36 pixels: most pixels are hit several times, so the
    last-point-wins rule of the reference's indexed write decides nearly every value."""
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    rng = np.random.RandomState(3)
    n_cam, n_pts, W, H = 12, 40000, 48, 36
    xyz = rng.uniform(-1, 1, (n_pts, 3))
    err = rng.uniform(0, 4, n_pts)
    tracks = [rng.randint(1, n_cam + 3, size=rng.randint(1, 6)) for _ in range(n_pts)]
    points = [(xyz[j], float(err[j]), tracks[j]) for j in range(n_pts)]
    ids = list(range(n_cam))
    extr, K = [], []
    for i in range(n_cam):
        a = 0.5 * i
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        w2c = np.eye(4)
        w2c[:3, :3], w2c[:3, 3] = R, [0.05 * i, -0.02 * i, 2.2]
        extr.append(w2c)
        K.append(np.array([[0.8 + 0.01 * i, 0, 0.5], [0, 1.1, 0.48], [0, 0, 1.0]]))
    tp = np.repeat(np.arange(n_pts), [len(t_) for t_ in tracks])
    ti = np.concatenate(tracks)
    ecam, ept = U.camera_point_entries(err, tp, ti, ids, 2.0)
    intr = np.array([[k[0, 0], k[1, 1], k[0, 2], k[1, 2]] for k in K])
    depth, bounds = U.sparse_depth_maps(xyz, ecam, ept, np.array([e[:3].reshape(12) for e in extr]), intr, W, H)
    depth, bounds = depth.cpu().numpy(), bounds.cpu().numpy()
    filled = 0
    for i in range(n_cam):
        want, mask, wb = O.sparse_depth_for_camera(points, extr[i], K[i], (W, H), i, 1, 2.0)
        assert np.array_equal(depth[i] != -1, mask[0])
        np.testing.assert_allclose(depth[i], want[0], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(bounds[i], wb, rtol=1e-12)
        filled += int(mask.sum())
    assert filled > 0.6 * n_cam * W * H and len(ecam) > 3 * filled      # most pixels are contested


@pytest.mark.gpu
def test_cuda_path_edge_cases():
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    eye = np.array([[1.0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]])
    intr = np.array([[1.0, 1.0, 0.5, 0.5]])
    none = np.zeros(0, np.int32)
    # no entries at all: every pixel is -1, bounds say "no point"
    depth, bounds = U.sparse_depth_maps(np.zeros((0, 3)), none, none, eye, intr, 8, 6)
    assert (depth == -1).all() and torch.isinf(bounds[0, 0])
    # behind the camera / outside the frustum / exactly on the lower edges (kept, pixel 0)
    xyz = np.array([[0.0, 0.0, -1.0], [5.0, 0.0, 1.0], [-0.5, -0.5, 1.0], [0.0, 0.0, 2.0]])
    cam = np.zeros(4, np.int32)
    depth, bounds = U.sparse_depth_maps(xyz, cam, np.arange(4, dtype=np.int32), eye, intr, 8, 6)
    d = depth.cpu().numpy()[0]
    assert d[0, 0] == 1.0 and d[4, 3] == 2.0 and (d != -1).sum() == 2
    assert bounds.cpu().numpy()[0].tolist() == [1.0, 2.0]
    # a point exactly on x == W: the reference's indexed write raises IndexError
    with pytest.raises(IndexError):
        U.sparse_depth_maps(np.array([[0.5, 0.0, 1.0]]), np.zeros(1, np.int32), np.zeros(1, np.int32), eye, intr, 8, 6)


@pytest.mark.gpu
def test_mixed_image_sizes_and_text_model_vs_oracle(tmp_path):
    """Views of two different resolutions (one launch per image size) and a text-format sparse model, end to end
    through `generate_sparse_depth_from_colmap`, against the oracle's per-camera loop."""
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    rng = np.random.RandomState(11)
    root = tmp_path / "ds"
    os.makedirs(root / "sparse_txt")
    n_cam, n_pts = 5, 3000
    frames = []
    for i in range(n_cam):
        a = 0.3 * (i - 2)
        c2w = np.eye(4)
        c2w[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        c2w[:3, 3] = [-2.4 * np.sin(a), 0.05 * i, -2.4 * np.cos(a)]
        m = c2w.copy()
        m[0:3, 1] *= -1
        m[0:3, 2] *= -1
        m = m[[1, 0, 2, 3], :]
        W, H = ((40, 30), (56, 44))[i % 2]
        frames.append({"id": i, "file_path": f"images/v{i}.png", "transform_matrix": m.tolist(),
                       "intrinsic": {"f_x": 1.0, "f_y": 1.3, "c_x": 0.5, "c_y": 0.5,
                                     "original_resolution_x": W, "original_resolution_y": H}})
    with open(root / "views_data.json", "w") as fh:
        json.dump({"frames": frames[::-1]}, fh)
    lines = ["# 3D point list"]
    for k in range(n_pts):
        x, y, z = (float(v) for v in rng.uniform(-0.9, 0.9, 3))
        ids = rng.randint(1, n_cam + 1, size=rng.randint(1, 4))
        track = " ".join(f"{int(i)} {int(rng.randint(0, 9))}" for i in ids)
        lines.append(f"{k + 1} {x!r} {y!r} {z!r} 9 9 9 {float(rng.uniform(0.1, 3.0))!r} {track}")
    (root / "sparse_txt" / "points3D.txt").write_text("\n".join(lines) + "\n")
    (root / "sparse_txt" / "cameras.txt").write_text("# none\n")
    (root / "sparse_txt" / "images.txt").write_text("# none\n")
    U.generate_sparse_depth_from_colmap(str(root), sparse_colmap_output_path=str(root / "sparse_txt"), scale=1,
                                        error_threshold=1.5)
    xyz, err, tp, ti = U.read_points3d_tracks(root / "sparse_txt")
    points = [(xyz[j], float(err[j]), ti[tp == j]) for j in range(n_pts)]
    out = json.load(open(root / "views_data.json"))
    cams = {c[3]: c for c in O.simple_sfm_json_cameras({"frames": frames})}
    total = 0
    for f in out["frames"]:
        w2c, K, size, cid, name = cams[f["id"]]
        want, mask, _ = O.sparse_depth_for_camera(points, w2c, K, size, cid, 1, 1.5)
        d = np.load(root / f["sparse_depth_path"], allow_pickle=True).item()
        assert d["sparse_depth"].shape == (1, size[0], size[1])
        assert np.array_equal(d["sparse_depth_mask"], mask)
        np.testing.assert_allclose(d["sparse_depth"], want, rtol=1e-12, atol=1e-12)
        total += int(mask.sum())
    assert total > 1000
