"""Post-match consumer of the matching path: sparse depth maps from a COLMAP sparse model.

Mirror of `generate_sparse_depth_from_colmap` (reference simple_sfm/dataset/simple_sfm_dataset_utils.py:147-261):
same arguments, same files written (`sparse_depth/sparse_depth_<name>_<id>.npy`, the visualisation JPEGs, the
`sparse_depth_path` entries of `views_data.json`).  The reference walks every 3-D point for every camera in Python
(`O(cameras x points)` dict look-ups) and projects camera by camera; here the host flattens the tracks of the model
once into (camera, point) entries (numpy, no per-point Python) and ONE launch of `sfm_sparse_depth` projects and
scatters all cameras of an image size in float64 (csrc/sparse_depth.cu).  No CPU fallback: the projection needs the
CUDA library.
"""
from __future__ import annotations

import json
import logging
import os
import struct
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np
import torch

from .._lib import check, lib, ptr, stream_ptr

logger = logging.getLogger(__name__)


# ---- COLMAP points3D readers (reference colmap_utils/read_write_colmap_data.py:330-407; standard COLMAP layout) -----
def read_points3d_tracks(model_path) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """points3D.{bin,txt} of a sparse model -> (xyz (P,3) f64, error (P) f64, track_point (T) i64, track_image (T) i64)
    in FILE order (the order of the reference's `points3D` dict).  The format is detected like `read_model`
    (:410-427): all three `.bin` files, else all three `.txt` files."""
    model_path = str(model_path)

    def has(ext):
        return all(os.path.isfile(os.path.join(model_path, f"{n}.{ext}")) for n in ("cameras", "images", "points3D"))

    if has("bin"):
        raw = open(os.path.join(model_path, "points3D.bin"), "rb").read()
        (n_pts,) = struct.unpack_from("<Q", raw, 0)
        xyz = np.empty((n_pts, 3), np.float64)
        err = np.empty(n_pts, np.float64)
        tr_len = np.empty(n_pts, np.int64)
        tracks: List[np.ndarray] = []
        off = 8
        for k in range(n_pts):                                       # one fixed 43-byte record + a variable track
            xyz[k] = struct.unpack_from("<3d", raw, off + 8)
            err[k] = struct.unpack_from("<d", raw, off + 35)[0]
            (tl,) = struct.unpack_from("<Q", raw, off + 43)
            tracks.append(np.frombuffer(raw, "<i4", 2 * tl, off + 51)[0::2])
            tr_len[k] = tl
            off += 51 + 8 * tl
    elif has("txt"):
        xs, es, tracks, tl_ = [], [], [], []
        with open(os.path.join(model_path, "points3D.txt")) as fh:
            for line in fh:
                line = line.strip()
                if not line or line[0] == "#":
                    continue
                el = line.split()
                xs.append([float(v) for v in el[1:4]])
                es.append(float(el[7]))
                tracks.append(np.array(el[8::2], np.int64))
                tl_.append(len(tracks[-1]))
        xyz = np.array(xs, np.float64).reshape(-1, 3)
        err = np.array(es, np.float64)
        tr_len = np.array(tl_, np.int64)
    else:
        raise FileNotFoundError(f"no COLMAP sparse model (cameras/images/points3D .bin or .txt) in {model_path}")
    track_image = np.concatenate(tracks).astype(np.int64) if tracks else np.zeros(0, np.int64)
    track_point = np.repeat(np.arange(len(tr_len), dtype=np.int64), tr_len)
    return xyz, err, track_point, track_image


# ---- views_data.json -> camera arrays (reference cameras/camera_multiple.py:212-276) -------------------------------
def read_simple_sfm_views(views_data_json_path) -> Dict[str, object]:
    """Per view, sorted by id: world-to-camera [R|t] rows (12 f64), (f_x, f_y, c_x, c_x') as the reference builds its
    intrinsic matrix -- NOTE its second principal-point entry is `c_x` again (camera_multiple.py:252), reproduced
    here --, image size, id and file path."""
    with open(views_data_json_path) as fh:
        views = json.load(fh)
    views = sorted(views["frames"], key=lambda item: item["id"])
    extr, intr, sizes, ids, names = [], [], [], [], []
    for view in views:
        k = view["intrinsic"]
        c2w = np.array(view["transform_matrix"], dtype=np.float64)
        c2w = c2w[[1, 0, 2, 3], :]
        c2w[0:3, 1] *= -1
        c2w[0:3, 2] *= -1
        w2c = np.linalg.inv(c2w)
        extr.append(w2c[:3, :].reshape(12))
        intr.append([k["f_x"], k["f_y"], k["c_x"], k["c_x"]])
        sizes.append([k["original_resolution_x"], k["original_resolution_y"]])
        ids.append(int(view["id"]))
        names.append(view["file_path"])
    return {"extrinsics": np.array(extr, np.float64).reshape(-1, 12), "intrinsics": np.array(intr, np.float64).reshape(-1, 4),
            "sizes": sizes, "ids": ids, "names": names}


def camera_point_entries(err: np.ndarray, track_point: np.ndarray, track_image: np.ndarray, camera_ids,
                         error_threshold: float) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's selection (:204-209) as arrays: point p belongs to camera c when `error_threshold > error[p]`
    and `camera_id + 1` occurs in its track (a membership test: repeated observations count once).  Returns
    (entry_cam index into camera_ids, entry_point) with the points of a camera in model order."""
    keep = error_threshold > err[track_point]
    tp, ti = track_point[keep], track_image[keep]
    cam_of_image = {int(cid) + 1: k for k, cid in enumerate(camera_ids)}
    lut_size = int(max(ti.max(initial=0), max(cam_of_image, default=0))) + 2
    lut = np.full(lut_size, -1, np.int64)
    for image_id, k in cam_of_image.items():
        if 0 <= image_id < lut_size:
            lut[image_id] = k
    ok = ti >= 0
    cam = np.where(ok, lut[np.where(ok, ti, 0)], -1)
    sel = cam >= 0
    pairs = np.unique(np.stack([cam[sel], tp[sel]], 1), axis=0)       # sorted by (camera, point), duplicates removed
    return pairs[:, 0].astype(np.int32), pairs[:, 1].astype(np.int32)


def sparse_depth_maps(xyz, entry_cam, entry_point, extrinsics, intrinsics, W: int, H: int, device="cuda"):
    """All cameras of one image size: (depth (n_cams, W, H) f64 on the device, -1 where no point landed; bounds
    (n_cams, 2) f64 = min / max depth over every point inside the frustum, +inf / 0 for a camera without one)
    (reference :195-231).  A point that projects exactly onto x == W or y == H raises IndexError like the
    reference's indexed write."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("sparse_depth_maps needs a CUDA device (libsfm_b200 has no CPU fallback)")
    n_cams = int(len(extrinsics))
    f64 = dict(dtype=torch.float64, device=device)
    d_xyz = torch.as_tensor(np.ascontiguousarray(xyz, np.float64)).to(device)
    d_cam = torch.as_tensor(np.ascontiguousarray(entry_cam, np.int32)).to(device)
    d_pt = torch.as_tensor(np.ascontiguousarray(entry_point, np.int32)).to(device)
    d_ex = torch.as_tensor(np.ascontiguousarray(extrinsics, np.float64)).to(device)
    d_in = torch.as_tensor(np.ascontiguousarray(intrinsics, np.float64)).to(device)
    depth = torch.empty(n_cams, W, H, **f64)
    bounds = torch.empty(max(n_cams, 1), 2, **f64)
    winner = torch.empty(max(n_cams * W * H, 1), dtype=torch.int64, device=device)
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().sfm_sparse_depth(ptr(d_xyz), int(len(xyz)), ptr(d_cam), ptr(d_pt), int(len(entry_cam)), ptr(d_ex),
                                     ptr(d_in), n_cams, int(W), int(H), ptr(depth), ptr(bounds), ptr(winner),
                                     ptr(flag),
                                     stream_ptr(device)))
    if int(flag.item()):
        raise IndexError(f"a point projects onto the x == {W} / y == {H} edge: index out of bounds "
                         f"(the reference's indexed write fails the same way)")
    return depth, bounds[:n_cams]


def generate_sparse_depth_from_colmap(
        dataset_path: str,
        sparse_colmap_output_path: str = None,
        scale: float = 1,
        error_threshold: float = 2.0,
        device="cuda",
):
    """
    Generate binary files with sparse depth based on the point cloud,
    which is restored during COLMAP SFM (sparse reconstruction).

    Output binary files are dicts with two items:
        'sparse_depth' - sparse depth with float values
        'sparse_depth_mask' - binary mask of pixels where sparse depth has values.

    :param dataset_path: path to dataset (folder with 'views_data.json' file)
    :param sparse_colmap_output_path: path to sparse COLMAP reconstruction
    :param scale: scale factor
    :param error_threshold: points with a larger error will be filtered.
    """
    from PIL import Image

    colmap_path = Path(dataset_path, "colmap/sparse") if sparse_colmap_output_path is None else sparse_colmap_output_path
    views_data_json_path = Path(dataset_path, "views_data.json")
    xyz, err, track_point, track_image = read_points3d_tracks(colmap_path)
    views = read_simple_sfm_views(views_data_json_path)

    save_sparse_depth_path = Path(dataset_path, "sparse_depth")
    os.makedirs(save_sparse_depth_path, exist_ok=True)
    save_sparse_depth_vis_path = Path(save_sparse_depth_path, "sparse_depth_vis")
    os.makedirs(save_sparse_depth_vis_path, exist_ok=True)
    logger.info(f"Num founded 3d points {len(xyz)}.")

    ids = views["ids"]
    entry_cam, entry_point = camera_point_entries(err, track_point, track_image, ids, error_threshold)
    sizes = [(int(w // scale), int(h // scale)) for w, h in views["sizes"]]
    depth_of: Dict[int, np.ndarray] = {}
    bounds_of: Dict[int, np.ndarray] = {}
    for size in sorted(set(sizes)):                                   # one launch per image size
        cams = np.array([k for k, s in enumerate(sizes) if s == size], np.int64)
        local = np.full(len(ids), -1, np.int64)
        local[cams] = np.arange(len(cams))
        sel = local[entry_cam] >= 0
        maps, bounds = sparse_depth_maps(xyz, local[entry_cam[sel]], entry_point[sel], views["extrinsics"][cams],
                                         views["intrinsics"][cams], size[0], size[1], device=device)
        maps, bounds = maps.cpu().numpy(), bounds.cpu().numpy()
        for j, k in enumerate(cams):
            depth_of[int(k)], bounds_of[int(k)] = maps[j], bounds[j]

    sparse_depth_path_dict = {}
    for k, camera_id in enumerate(ids):
        sparse_depth = depth_of[k][:, :, None]                          # (W, H, 1) like the reference tensor
        depth_mask = sparse_depth != -1
        if not depth_mask.any():
            logger.info(f"There is no confident 3d points for camera {camera_id}.")
        # bounds over every point that passed the frustum test, overwritten ones included (:228)
        points_bounds = list(bounds_of[k]) if np.isfinite(bounds_of[k][0]) else [-1, 1]
        camera_name_index = views["names"][k].split("/")[-1] + "_" + str(camera_id)
        img_depth = (np.repeat(sparse_depth.transpose(1, 0, 2), 3, axis=2) - points_bounds[0]) / (
                points_bounds[1] - points_bounds[0] + 10e-6)
        Image.fromarray((img_depth * 255).astype(np.uint8)).save(
            Path(save_sparse_depth_vis_path, f"sparse_depth_vis_{camera_name_index}.jpg"))
        np.save(Path(save_sparse_depth_path, f"sparse_depth_{camera_name_index}"),
                {"sparse_depth": np.ascontiguousarray(sparse_depth.transpose(2, 0, 1)),
                 "sparse_depth_mask": np.ascontiguousarray(depth_mask.transpose(2, 0, 1))})
        sparse_depth_path_dict[camera_id] = Path("sparse_depth", f"sparse_depth_{camera_name_index}.npy")

    with open(views_data_json_path, encoding="UTF-8") as file:
        views_data_json = json.load(file)
    for i in range(len(views_data_json["frames"])):
        views_data_json["frames"][i]["sparse_depth_path"] = str(
            sparse_depth_path_dict[views_data_json["frames"][i]["id"]])
    with open(views_data_json_path, "w") as outfile:
        json.dump(views_data_json, outfile, indent=2)
