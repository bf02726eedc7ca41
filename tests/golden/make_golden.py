"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference has no tests or fixtures of its own, so parity is pinned on
outputs of its own modules executed here on CPU with seeded synthetic weights in the real
key layout.  The `.npz` files written next to this script are committed; nothing on the GPU
box reads /root/reference.  The only shim applied is `.cuda()` -> no-op (the reference calls
bare `.cuda()` at matcher.py:48,58,210,...), nothing else is patched.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from simplesfm_b200 import synthetic as S  # noqa: E402

torch.nn.Module.cuda = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.empty_cache = lambda: None

from simple_sfm.models import superpoint as RSP  # noqa: E402
from simple_sfm.models import superglue as RSG  # noqa: E402
from simple_sfm.matchers.matcher import Matcher, Frame  # noqa: E402
from simple_sfm.matchers.simple_matcher import SimpleMatcher  # noqa: E402

torch.set_num_threads(8)
TMP = tempfile.mkdtemp()


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KB")


def golden_superpoint():
    wpath = S.save_state_dict(S.superpoint_state_dict(0), os.path.join(TMP, "sp.pth"))
    # --- small image, every stage -------------------------------------------------------
    img = S.textured_image(3, 120, 160)
    mask = torch.ones(1, 1, 120, 160, dtype=torch.bool)
    mask[:, :, 30:70, 50:120] = False
    for tag, mk, use_mask in (("all", -1, False), ("top200", 200, False), ("masked", -1, True)):
        m = RSP.SuperPoint(wpath, nms_radius=4, keypoint_threshold=0.005, max_keypoints=mk)
        stages = {}
        with torch.no_grad():
            x = img
            for a, b, pool in (("conv1a", "conv1b", True), ("conv2a", "conv2b", True),
                               ("conv3a", "conv3b", True), ("conv4a", "conv4b", False)):
                x = m.relu(getattr(m, b)(m.relu(getattr(m, a)(x))))
                if pool:
                    x = m.pool(x)
            stages["feat"] = x
            logits = m.convPb(m.relu(m.convPa(x)))
            stages["logits"] = logits
            sc = torch.nn.functional.softmax(logits, 1)[:, :-1]
            b_, _, h, w = sc.shape
            sc = sc.permute(0, 2, 3, 1).reshape(b_, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(b_, h * 8, w * 8)
            stages["score_map"] = sc
            stages["nms"] = RSP.simple_nms(sc, 4)
            stages["dmap"] = torch.nn.functional.normalize(m.convDb(m.relu(m.convDa(x))), p=2, dim=1)
            data = {"image": img}
            if use_mask:
                data["mask"] = mask
            out = m(data)
        if tag == "all":
            npz("sp_small_stages", image=img, **stages)
        npz(f"sp_small_{tag}", image=img, mask=mask if use_mask else np.zeros(0),
            keypoints=out["keypoints"], scores=out["scores"], descriptors=out["descriptors"])
    # --- full 480x640, final outputs (descriptors of the 128 best only) ------------------
    img = S.textured_image(1, 480, 640)
    m = RSP.SuperPoint(wpath, nms_radius=4, keypoint_threshold=0.005, max_keypoints=1024)
    with torch.no_grad():
        out = m({"image": img})
    npz("sp_480x640_top1024", seed=1, keypoints=out["keypoints"], scores=out["scores"],
        descriptors_first128=out["descriptors"][:, :128])
    # --- free functions ------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    smap = torch.rand(2, 64, 96, generator=g)
    smap[0, 10:30, 10:30] = 0.7                       # plateau: every interior pixel survives
    smap[1, 5, 5] = 2.0
    npz("sp_functions",
        nms_in=smap, nms_r4=RSP.simple_nms(smap, 4), nms_r2=RSP.simple_nms(smap, 2), nms_r0=RSP.simple_nms(smap, 0),
        samp_kpts=(kp := torch.stack([torch.randint(0, 160, (50,), generator=g),
                                     torch.randint(0, 120, (50,), generator=g)], 1).float()),
        samp_dmap=(dm := torch.nn.functional.normalize(torch.randn(1, 256, 15, 20, generator=g), dim=1)),
        samp_out=RSP.sample_descriptors(kp.clone()[None], dm, 8)[0],
        torch_version=np.array(torch.__version__))


def golden_superpoint_odd():
    """Image sides that are not multiples of 8 (the reference accepts them when no mask is given: the convolutions
    run at the full size, the pools floor, keypoints come from the (h*8, w*8) score map; superpoint.py:112-171)."""
    wpath = S.save_state_dict(S.superpoint_state_dict(0), os.path.join(TMP, "sp.pth"))
    for tag, (H, W), mk in (("77x133", (77, 133), -1), ("101x90_top150", (101, 90), 150)):
        img = S.textured_image(7, H + 3, W + 3)[:, :, :H, :W].contiguous()
        m = RSP.SuperPoint(wpath, nms_radius=4, keypoint_threshold=0.005, max_keypoints=mk)
        with torch.no_grad():
            x = img
            for a, b, pool in (("conv1a", "conv1b", True), ("conv2a", "conv2b", True),
                               ("conv3a", "conv3b", True), ("conv4a", "conv4b", False)):
                x = m.relu(getattr(m, b)(m.relu(getattr(m, a)(x))))
                if pool:
                    x = m.pool(x)
            sc = torch.nn.functional.softmax(m.convPb(m.relu(m.convPa(x))), 1)[:, :-1]
            b_, _, h, w = sc.shape
            sc = sc.permute(0, 2, 3, 1).reshape(b_, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(b_, h * 8, w * 8)
            out = m({"image": img})
        npz(f"sp_odd_{tag}", image=img, feat=x, score_map=sc, keypoints=out["keypoints"], scores=out["scores"],
            descriptors=out["descriptors"])


def sg_inputs(B, K0, K1, seed):
    fr = S.feature_scene(2 * B, max(K0, K1), seed=seed)
    return {"descriptors0": torch.stack([fr[i][0][:, :K0] for i in range(B)]),
            "descriptors1": torch.stack([fr[B + i][0][:, :K1] for i in range(B)]),
            "keypoints0": torch.stack([fr[i][1][:K0] for i in range(B)]),
            "keypoints1": torch.stack([fr[B + i][1][:K1] for i in range(B)]),
            "scores0": torch.stack([fr[i][2][:K0] for i in range(B)]),
            "scores1": torch.stack([fr[B + i][2][:K1] for i in range(B)]),
            "image0": torch.zeros(B, 1, 480, 640), "image1": torch.zeros(B, 1, 480, 640)}


def golden_superglue():
    wpath = S.save_state_dict(S.superglue_state_dict(0, "indoor"), os.path.join(TMP, "sg.pth"))
    data = sg_inputs(2, 96, 80, seed=11)
    for mode in ("train", "eval"):
        for iters in (20, 100):
            m = RSG.SuperGlue(wpath, sinkhorn_iterations=iters, match_threshold=0.2)
            if mode == "eval":
                m.eval()
            keep = {}
            hooks = []
            for li in (0, 1, 17):
                hooks.append(m.gnn.layers[li].register_forward_hook(
                    lambda mod, inp, out, li=li: keep.setdefault(f"delta_l{li}", []).append(out.clone())))
            hooks.append(m.kenc.register_forward_hook(
                lambda mod, inp, out: keep.setdefault("kenc", []).append(out.clone())))
            hooks.append(m.final_proj.register_forward_hook(
                lambda mod, inp, out: keep.setdefault("mdesc", []).append(out.clone())))
            with torch.no_grad():
                out = m(data)
                md0, md1 = keep["mdesc"]
                scores = torch.einsum("bdn,bdm->bnm", md0, md1) / 16.0
                Z = RSG.log_optimal_transport(scores, m.bin_score, iters)
            extra = {}
            for k, v in keep.items():
                extra[k + "_0"], extra[k + "_1"] = v[0], v[1]
            if iters == 100:
                extra = {}   # per-layer tensors are identical; keep the file small
            npz(f"sg_b2_{mode}_it{iters}", **{k: v for k, v in data.items() if "image" not in k},
                image_hw=np.array([480, 640]), scores=scores, Z=Z, **extra, **out)
    # empty side
    m = RSG.SuperGlue(wpath)
    d = sg_inputs(1, 8, 8, seed=3)
    d["keypoints1"] = d["keypoints1"][:, :0]
    d["descriptors1"] = d["descriptors1"][:, :, :0]
    d["scores1"] = d["scores1"][:, :0]
    with torch.no_grad():
        out = m(d)
    npz("sg_empty_side", **{k: v for k, v in out.items()},
        dtypes=np.array([str(out["matches0"].dtype), str(out["matching_scores0"].dtype)]))
    # free functions
    g = torch.Generator().manual_seed(9)
    sc = torch.randn(3, 37, 29, generator=g) * 3
    q = torch.randn(2, 64, 4, 33, generator=g)
    k = torch.randn(2, 64, 4, 21, generator=g)
    v = torch.randn(2, 64, 4, 21, generator=g)
    kp = torch.rand(2, 17, 2, generator=g) * 600
    npz("sg_functions",
        ot_scores=sc, ot_alpha=np.float32(1.0),
        ot_it20=RSG.log_optimal_transport(sc, torch.tensor(1.0), 20),
        ot_it100=RSG.log_optimal_transport(sc, torch.tensor(1.0), 100),
        ot_alpha2_it5=RSG.log_optimal_transport(sc, torch.tensor(2.5), 5),
        att_q=q, att_k=k, att_v=v, att_out=RSG.attention(q, k, v)[0],
        nk_in=kp, nk_out=RSG.normalize_keypoints(kp, (2, 1, 480, 640)),
        nk_out_tall=RSG.normalize_keypoints(kp, (2, 1, 640, 480)),
        arange=RSG.arange_like(torch.zeros(2, 7), 1))


def golden_matcher():
    sp = S.save_state_dict(S.superpoint_state_dict(0), os.path.join(TMP, "sp.pth"))
    sg = S.save_state_dict(S.superglue_state_dict(1, "outdoor"), os.path.join(TMP, "sgo.pth"))
    fr = S.feature_scene(4, 160, seed=21)
    counts = [160, 131, 160, 97]                     # ragged: exercises chunk-min truncation
    frames = [Frame(fr[i][0][:, :c], fr[i][1][:c], fr[i][2][:c], torch.zeros(1, 480, 640))
              for i, c in enumerate(counts)]
    names = {i + 1: f"img_{i:03d}.png" for i in range(4)}
    save = {}
    for i, f in enumerate(frames):
        save[f"desc_{i}"], save[f"kpts_{i}"], save[f"scores_{i}"] = f.descriptors, f.keypoints_poses, f.scores
    for batch in (4, 1):
        m = Matcher(sp, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg,
                    match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=batch)
        _, table, _ = m.match_frames(frames, names, lambda p: True)
        for (a, b), t in table.items():
            save[f"sg_b{batch}_{int(a)}_{int(b)}"] = t
        save[f"sg_b{batch}_keytype"] = np.array(str(type(next(iter(table.keys()))[0])))
    m = Matcher(sp, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg,
                match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=4)
    _, table, _ = m.match_frames(frames, names, lambda p: abs(p[0] - p[1]) <= 2)
    save["sg_filtered_keys"] = np.array(sorted((int(a), int(b)) for a, b in table.keys()))
    m = Matcher(sp, 4, 0.005, matcher_type="simple", match_threshold=0.7)
    _, table, _ = m.match_frames(frames, names, lambda p: True)
    for (a, b), t in table.items():
        save[f"simple_{int(a)}_{int(b)}"] = t
    sm = SimpleMatcher(0.7)
    save["simple_two_way_np"] = sm.match_two_way(frames[0].descriptors, frames[1].descriptors)
    save["simple_two_way_cuda"] = sm.match_two_way_cuda(frames[0].descriptors, frames[1].descriptors)
    npz("matcher_4frames", counts=np.array(counts), **save)
    # full image -> frames -> tables through the reference Matcher's own extraction code path
    imgs = [S.textured_image(40, 120, 160, shift=(0, 0)), S.textured_image(40, 120, 160, shift=(3, 5)),
            S.textured_image(40, 120, 160, shift=(-4, 2))]
    m = Matcher(sp, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg,
                match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=2)
    frs = []
    with torch.no_grad():
        for im in imgs:
            r = m.super_point_extractor({"image": im})
            frs.append(Frame(r["descriptors"], r["keypoints"], r["scores"], im[0]))
    _, table, _ = m.match_frames(frs, {1: "a", 2: "b", 3: "c"}, lambda p: True)
    save = {f"image_{i}": im for i, im in enumerate(imgs)}
    for i, f in enumerate(frs):
        save[f"kpts_{i}"], save[f"scores_{i}"] = f.keypoints_poses, f.scores
    for (a, b), t in table.items():
        save[f"table_{int(a)}_{int(b)}"] = t
    npz("matcher_images_e2e", **save)


def golden_module_functions():
    """Module-level functions the reference exports besides the classes (superpoint.py:27-39,
    superglue.py:142-148,174-175) and `Matcher.match_datasets` (matcher.py:167-247) over the reference's own
    `ImageFoldersDataset` (utils/datasets.py:16).  The JPEG files are stored byte for byte so that the GPU test
    decodes exactly what the reference decoded."""
    import io
    from PIL import Image
    from simple_sfm.utils.datasets import ImageFoldersDataset
    g = torch.Generator().manual_seed(4321)
    kp = torch.stack([torch.randint(0, 120, (400,), generator=g), torch.randint(0, 160, (400,), generator=g)], 1)
    sc = torch.rand(400, generator=g)
    rb_k, rb_s = RSP.remove_borders(kp, sc, 6, 120, 160)
    tk_k, tk_s = RSP.top_k_keypoints(kp, sc, 123)
    Z = torch.randn(2, 37, 53, generator=g) * 2
    log_mu = torch.log_softmax(torch.randn(2, 37, generator=g), -1)
    log_nu = torch.log_softmax(torch.randn(2, 53, generator=g), -1)
    save = dict(rb_in_kpts=kp, rb_in_scores=sc, rb_border=np.int64(6), rb_h=np.int64(120), rb_w=np.int64(160),
                rb_out_kpts=rb_k, rb_out_scores=rb_s, tk_k=np.int64(123), tk_out_kpts=tk_k, tk_out_scores=tk_s,
                lsi_Z=Z, lsi_log_mu=log_mu, lsi_log_nu=log_nu,
                lsi_out_25=RSG.log_sinkhorn_iterations(Z, log_mu, log_nu, 25),
                lsi_out_0=RSG.log_sinkhorn_iterations(Z, log_mu, log_nu, 0),
                arange_d0=RSG.arange_like(torch.zeros(3, 7, 5), 0), arange_d2=RSG.arange_like(torch.zeros(3, 7, 5), 2))
    # ---- match_datasets: two folders (3 + 2 images) with masks, reference dataset class + reference Matcher
    sp = S.save_state_dict(S.superpoint_state_dict(0), os.path.join(TMP, "sp.pth"))
    sg = S.save_state_dict(S.superglue_state_dict(0, "indoor"), os.path.join(TMP, "sgi.pth"))

    def grey(img):
        return torch.from_numpy(np.asarray(img.convert("L"), dtype=np.float32) / 255.0)[None]

    for masked in (False, True):
        sets = []
        tag = "md_masked" if masked else "md"
        for d, seeds in (("a", (0, 1, 2)), ("b", (3, 4))):
            rgb, msk = os.path.join(TMP, tag, d, "rgb"), os.path.join(TMP, tag, d, "mask")
            os.makedirs(rgb)
            os.makedirs(msk)
            for i, sd in enumerate(seeds):
                im = (S.textured_image(40 + (sd % 2), 120, 160, shift=(sd, 2 * sd))[0, 0] * 255).round().byte().numpy()
                mk = np.zeros((120, 160), np.uint8)
                mk[10:110, 20:150] = 255
                for arr, folder, q in ((im, rgb, 95), (mk, msk, 100)):
                    buf = io.BytesIO()
                    Image.fromarray(arr, mode="L").save(buf, format="JPEG", quality=q)
                    with open(os.path.join(folder, f"{d}{i:02d}.jpg"), "wb") as fh:
                        fh.write(buf.getvalue())
                    if not masked:
                        save[f"jpg_{os.path.basename(folder)}_{d}{i:02d}"] = np.frombuffer(buf.getvalue(), np.uint8)
            folders, tf = {"rgb": rgb}, {"rgb": grey}
            if masked:
                folders["mask"], tf["mask"] = msk, grey
            sets.append(ImageFoldersDataset(folders, tf))
        m = Matcher(sp, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sg,
                    match_threshold=0.2, sinkhorn_iterations=20, super_glue_batch=2)
        frames, table, names, split = m.match_datasets(sets, None, 1)
        save[f"{tag}_names"] = np.array([names[i + 1] for i in range(len(names))])
        save[f"{tag}_split"] = np.array(split)
        for i, f in enumerate(frames):
            save[f"{tag}_kpts_{i}"], save[f"{tag}_scores_{i}"] = f.keypoints_poses, f.scores
            save[f"{tag}_image_shape_{i}"] = np.array(f.image.shape)
        for (a, b), tb in table.items():
            save[f"{tag}_table_{int(a)}_{int(b)}"] = tb
    npz("module_functions", **save)


def golden_sparse_depth():
    """Post-match consumer `generate_sparse_depth_from_colmap` (dataset/simple_sfm_dataset_utils.py:147-261): a tiny
    synthetic dataset (views_data.json + a COLMAP sparse model written with the reference's own writer) through the
    unmodified reference function; the input files are stored byte for byte, the outputs as arrays."""
    import json
    import shutil
    from simple_sfm.colmap_utils import read_write_colmap_data as RW
    from simple_sfm.dataset import simple_sfm_dataset_utils as U
    rng = np.random.RandomState(7)
    root = os.path.join(TMP, "depth_ds")
    os.makedirs(os.path.join(root, "colmap", "sparse"))
    n_cam, n_pts, Wd, Hd = 6, 400, 64, 48
    frames = []
    for i in range(n_cam):
        ang = 0.35 * (i - n_cam / 2)
        c, s_ = np.cos(ang), np.sin(ang)
        R = np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])          # camera-to-world rotation (before the json quirks)
        t = np.array([-2.5 * s_, 0.1 * i, -2.5 * c])
        c2w = np.eye(4)
        c2w[:3, :3], c2w[:3, 3] = R, t
        # undo what from_simple_sfm_json does (camera_multiple.py:242-247) so that the camera looks at the origin
        m = c2w.copy()
        m[0:3, 1] *= -1
        m[0:3, 2] *= -1
        m = m[[1, 0, 2, 3], :]
        frames.append({"id": (i * 5) % n_cam if False else i, "file_path": f"images/view_{i:02d}.png",
                       "transform_matrix": m.tolist(),
                       "intrinsic": {"f_x": 0.9 + 0.01 * i, "f_y": 1.2, "c_x": 0.5, "c_y": 0.47,
                                     "original_resolution_x": Wd, "original_resolution_y": Hd}})
    frames = [frames[k] for k in (3, 0, 5, 1, 4, 2)]                 # the reference sorts by id
    with open(os.path.join(root, "views_data.json"), "w") as fh:
        json.dump({"frames": frames}, fh)
    pts = {}
    for k in range(n_pts):
        xyz = rng.uniform(-0.9, 0.9, 3) * np.array([1.0, 0.7, 1.0])
        n_obs = rng.randint(1, 5)
        ids = rng.randint(1, n_cam + 1, size=n_obs)                 # may repeat an image (membership test in the reference)
        pts[k + 10] = RW.Point3D(id=k + 10, xyz=xyz, rgb=np.array([1, 2, 3]), error=float(rng.uniform(0.1, 3.5)),
                                 image_ids=ids, point2D_idxs=rng.randint(0, 50, size=n_obs))
    cams = {1: RW.Camera(id=1, model="SIMPLE_PINHOLE", width=Wd, height=Hd, params=np.array([50.0, 32.0, 24.0]))}
    ims = {i + 1: RW.BaseImage(id=i + 1, qvec=np.array([1.0, 0, 0, 0]), tvec=np.zeros(3), camera_id=1,
                               name=f"view_{i:02d}.png", xys=np.zeros((1, 2)), point3D_ids=np.array([-1]))
           for i in range(n_cam)}
    RW.write_model(cams, ims, pts, os.path.join(root, "colmap", "sparse"), ext="bin")
    save = {"views_json": np.frombuffer(open(os.path.join(root, "views_data.json"), "rb").read(), np.uint8)}
    for name in ("cameras.bin", "images.bin", "points3D.bin"):
        save["model_" + name.replace(".", "_")] = np.frombuffer(
            open(os.path.join(root, "colmap", "sparse", name), "rb").read(), np.uint8)
    for tag, scale, thr in (("s1", 1, 2.0), ("s2", 2, 3.0)):
        work = os.path.join(TMP, "depth_run_" + tag)
        shutil.copytree(root, work)
        U.generate_sparse_depth_from_colmap(work, scale=scale, error_threshold=thr)
        out = json.load(open(os.path.join(work, "views_data.json")))
        save[f"{tag}_paths"] = np.array([f["sparse_depth_path"] for f in out["frames"]])
        for f in out["frames"]:
            d = np.load(os.path.join(work, f["sparse_depth_path"]), allow_pickle=True).item()
            save[f"{tag}_depth_{f['id']}"] = d["sparse_depth"]
            save[f"{tag}_mask_{f['id']}"] = d["sparse_depth_mask"]
    npz("sparse_depth", **save)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (golden_superpoint, golden_superpoint_odd, golden_superglue, golden_matcher, golden_module_functions, golden_sparse_depth):
        if not only or fn.__name__ in only:
            fn()
