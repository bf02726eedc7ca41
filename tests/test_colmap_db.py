"""CPU: the binary-free COLMAP database writer produces the reference's payloads
(simple_sfm/colmap_utils/colmap_bd_utils.py:291-372), checked against the oracle's blob restatement."""
import os
import sqlite3

import numpy as np
import torch

import sfm_oracle as O


def test_colmap_writer_roundtrip(tmp_path):
    from simplesfm_b200.colmap_db import ColmapDatabaseWriter
    from simplesfm_b200.matchers.matcher import Frame
    g = torch.Generator().manual_seed(0)
    frames = [Frame(torch.zeros(256, k), torch.randint(4, 600, (k, 2), generator=g).float(), torch.rand(k),
                    torch.zeros(1, 480, 640)) for k in (7, 0, 12)]
    names = {1: "a.png", 2: "b.png", 3: "c.png"}
    table = {(np.int64(1), np.int64(3)): np.array([[0, 5], [3, 11]], np.uint32),
             (np.int64(3), np.int64(2)): np.zeros((0, 2), np.uint32),
             (np.int64(3), np.int64(1)): np.array([[2, 6]], np.uint32)}
    db = ColmapDatabaseWriter(str(tmp_path), images_folder_path=str(tmp_path), camera_size=(640, 480))
    db.replace_images_data(names)
    db.replace_keypoints(names, frames)
    db.replace_and_verificate_matches({k: v for k, v in list(table.items())[:2]}, names, run_verification=False)
    con = sqlite3.connect(os.path.join(str(tmp_path), "database.db"))
    assert con.execute("SELECT image_id, name, camera_id, prior_qw FROM images ORDER BY image_id").fetchall() == \
        [(1, "a.png", 1, None), (2, "b.png", 1, None), (3, "c.png", 1, None)]
    for i, f in enumerate(frames):
        r, c, blob = O.colmap_keypoint_blob(f.keypoints_poses.numpy())
        row = con.execute("SELECT rows, cols, data FROM keypoints WHERE image_id = ?", (i + 1,)).fetchone()
        assert row == (r, c, blob)
    assert con.execute("SELECT COUNT(*) FROM descriptors").fetchone()[0] == 0
    for (a, b), m in list(table.items())[:2]:
        pid, r, c, blob = O.colmap_match_blob(int(a), int(b), m)
        assert con.execute("SELECT rows, cols, data FROM matches WHERE pair_id = ?", (pid,)).fetchone() == (r, c, blob)
    assert np.array_equal(db.read_matches(1, 3), table[(1, 3)])
    # swapped ids: columns swapped, same pair id (colmap_bd_utils.py:359-360)
    db.replace_and_verificate_matches({(np.int64(3), np.int64(1)): table[(3, 1)]}, names, run_verification=False)
    assert np.array_equal(db.read_matches(1, 3), np.array([[6, 2]], np.uint32))
    assert open(os.path.join(str(tmp_path), "match.txt")).read() == "c.png a.png\n"
    assert db.image_ids_to_pair_id(5, 2) == 2 * 2147483647 + 5 == O.colmap_pair_id(5, 2)
    assert con.execute("SELECT model, width, height FROM cameras").fetchall() == [(2, 640, 480)]


def test_camera_from_first_image_and_external_steps(tmp_path):
    """Without an explicit size the shared camera comes from the first image file (what `colmap feature_extractor`
    records, colmap_bd_utils.py:256-273): width, height, f = 0.85 * max(w, h); without any size the writer raises.
    The COLMAP subprocess steps are pass-throughs that fail loudly when the binary is missing."""
    import shutil
    import sqlite3
    import pytest
    from PIL import Image
    from simplesfm_b200.colmap_db import ColmapDatabaseWriter
    imgs = tmp_path / "images"
    imgs.mkdir()
    Image.fromarray(np.zeros((300, 500), np.uint8)).save(imgs / "b.png")
    Image.fromarray(np.zeros((360, 720), np.uint8)).save(imgs / "a.png")       # first in name order
    db = ColmapDatabaseWriter(str(tmp_path / "colmap"), images_folder_path=str(imgs))
    con = sqlite3.connect(str(tmp_path / "colmap" / "database.db"))
    model, w, h, params, prior = con.execute("SELECT model, width, height, params, prior_focal_length FROM cameras").fetchone()
    assert (model, w, h, prior) == (2, 720, 360, 0)
    assert np.frombuffer(params, np.float64).tolist() == [0.85 * 720, 360.0, 180.0, 0.0]
    with pytest.raises(ValueError):
        ColmapDatabaseWriter(str(tmp_path / "colmap2"))
    with pytest.raises(ValueError):
        ColmapDatabaseWriter(str(tmp_path / "colmap3"), images_folder_path=str(tmp_path / "missing"))
    if shutil.which("colmap") is None:
        with pytest.raises(FileNotFoundError):
            db.triangulation(db.db_path, str(imgs), db.sparse_path, db.sparse_path)
        with pytest.raises(FileNotFoundError):
            db.run_mapper()
