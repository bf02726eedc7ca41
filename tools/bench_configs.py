"""Short secondary measurements of the other BASELINE.json configs, appended to the bench line under `configs`
(bench.py).  Every entry states what ran; each is sized to finish in a few seconds.

  config1  one 640x480 image pair, 1024 keypoints, 100 Sinkhorn iterations: CPU oracle (SuperPoint x 2 + SuperGlue)
           beside the GPU fp32 and bf16 paths                                              [single-GPU runs only]
  config3  generate_sparse_colmap flow on 50 frames: image + mask files -> Matcher.match_simple_sfm_dataset
           (outdoor layout, batch 4, masks) -> ColmapDatabaseWriter; wall time including decode [single-GPU only]
  config4  4096 keypoints, outdoor layout: pairs/s per GPU on a 19-image scene (171 pairs)
  config5  log_optimal_transport + mutual matching, N = M = 4096, 100 iterations, 256 pairs split over the ranks:
           exact log-space form, scaling form with fp32 storage, scaling form with fp16 storage
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sync_ms(fn, dev, reps=1):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps, out


def _max_over_ranks(x, dev, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config1(dev):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sfm_oracle as O
    from simplesfm_b200 import synthetic
    from simplesfm_b200.models.superglue import SuperGlue
    from simplesfm_b200.models.superpoint import SuperPoint
    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0, "indoor")
    imgs = [synthetic.textured_image(11, shift=(0, 0)), synthetic.textured_image(11, shift=(5, -7))]
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        of = [O.superpoint_forward(spw, im, None, 4, 0.005, 1024, 4) for im in imgs]
        t_sp = time.perf_counter() - t0
        data = {"descriptors0": of[0]["descriptors"][None], "descriptors1": of[1]["descriptors"][None],
                "keypoints0": of[0]["keypoints"][None], "keypoints1": of[1]["keypoints"][None],
                "scores0": of[0]["scores"][None], "scores1": of[1]["scores"][None], "image0": imgs[0], "image1": imgs[1]}
        ref = O.superglue_forward(sgw, data, 100, 0.2, "train")
    t_cpu = time.perf_counter() - t0
    out = {"what": "SuperPoint x 2 (max_keypoints 1024) + SuperGlue (100 Sinkhorn iterations, thr 0.2, train-mode BN), "
                   "one 640x480 synthetic pair",
           "cpu_oracle_s_per_pair": t_cpu, "cpu_oracle_superpoint_s": t_sp, "cpu_cores": os.cpu_count(),
           "cpu_pairs_per_s": 1.0 / t_cpu}
    sp = SuperPoint(spw, max_keypoints=1024, device=dev)
    dimg = [im.to(dev) for im in imgs]
    for prec in ("fp32", "bf16"):
        sg = SuperGlue(sgw, sinkhorn_iterations=100, match_threshold=0.2, precision=prec, device=dev)

        def pair():
            f = [sp({"image": im}) for im in dimg]
            return sg({"descriptors0": f[0]["descriptors"][None], "descriptors1": f[1]["descriptors"][None],
                       "keypoints0": f[0]["keypoints"][None], "keypoints1": f[1]["keypoints"][None],
                       "scores0": f[0]["scores"][None], "scores1": f[1]["scores"][None],
                       "image0": dimg[0], "image1": dimg[1]}), f
        pair()
        ms, (res, f) = _sync_ms(pair, dev, reps=3)
        out[f"gpu_{prec}_ms_per_pair"] = ms
        out[f"gpu_{prec}_pairs_per_s"] = 1e3 / ms
        if prec == "fp32":
            # same keypoint sets and (through coordinates) the same matches as the oracle
            same_kp = all({tuple(x) for x in a["keypoints"].cpu().numpy().tolist()} ==
                          {tuple(x) for x in b["keypoints"].numpy().tolist()} for a, b in zip(f, of))
            k0, k1 = f[0]["keypoints"].cpu().numpy(), f[1]["keypoints"].cpu().numpy()
            o0, o1 = of[0]["keypoints"].numpy(), of[1]["keypoints"].numpy()
            m, mr = res["matches0"][0].cpu().numpy(), ref["matches0"][0].numpy()
            got = {(tuple(k0[i]), tuple(k1[j])) for i, j in enumerate(m) if j > -1}
            want = {(tuple(o0[i]), tuple(o1[j])) for i, j in enumerate(mr) if j > -1}
            out["gpu_fp32_keypoint_sets_equal_oracle"] = bool(same_kp)
            out["gpu_fp32_match_set_agreement_with_oracle"] = len(got & want) / max(len(got | want), 1)
            out["oracle_matches"] = len(want)
    return out


def config3(dev, precision, n_frames=50):
    from PIL import Image
    from simplesfm_b200 import synthetic
    from simplesfm_b200.colmap_db import ColmapDatabaseWriter
    from simplesfm_b200.matchers import Matcher
    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(1, "outdoor")
    tmp = tempfile.mkdtemp(prefix="sfm_cfg3_")
    meta = []
    for i in range(n_frames):
        im = (synthetic.textured_image(300 + i % 7, shift=(3 * (i % 5), -2 * (i % 9)))[0, 0] * 255).round().to(torch.uint8)
        mk = np.zeros((480, 640), np.uint8)
        mk[40:440, 60 + 2 * (i % 11):600] = 255
        Image.fromarray(im.numpy()).save(os.path.join(tmp, f"view_{i:03d}.png"), compress_level=1)
        Image.fromarray(mk).save(os.path.join(tmp, f"view_{i:03d}_mask.png"), compress_level=1)
        meta.append({"file_path": f"./view_{i:03d}.png", "object_of_interest_mask_path": f"./view_{i:03d}_mask.png"})
    with open(os.path.join(tmp, "views_data.json"), "w") as fh:
        json.dump({"frames": meta}, fh)
    # generate_sparse_colmap defaults (simple_sfm_dataset_utils.py:82-87, 112-121): nms 4, keypoint thr 1e-4, match thr
    # 0.9, 20 iterations, batch 4, outdoor weights, masks.  With synthetic weights a match threshold of 0.9 would keep
    # nothing, so 0.4 is used (it does not change the work done).
    m = Matcher(spw, 4, 1e-4, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                sinkhorn_iterations=20, super_glue_batch=4, precision=precision, device=dev, extract_batch=8,
                decode_threads=8)

    def run(sub):
        t0 = time.perf_counter()
        frames, table, names = m.match_simple_sfm_dataset(sub, None, "object_of_interest_mask_path")
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        db = ColmapDatabaseWriter(os.path.join(sub, "colmap"), images_folder_path=sub)
        db.replace_images_data(names)
        db.replace_keypoints(names, frames)
        db.replace_and_verificate_matches(table, names, run_verification=False)
        db.close_bd()
        return frames, table, t1 - t0, time.perf_counter() - t1

    # warm-up on a 6-frame copy of the folder layout (kernel load, allocator)
    warm = tempfile.mkdtemp(prefix="sfm_cfg3w_")
    for e in meta[:6]:
        for k in ("file_path", "object_of_interest_mask_path"):
            os.link(os.path.join(tmp, e[k]), os.path.join(warm, e[k]))
    with open(os.path.join(warm, "views_data.json"), "w") as fh:
        json.dump({"frames": meta[:6]}, fh)
    run(warm)
    frames, table, t_match, t_db = run(tmp)
    kp = [int(f.keypoints_poses.shape[0]) for f in frames]
    n_pairs = len(table)
    return {"what": f"{n_frames} frames (480x640 PNG + mask PNG) -> match_simple_sfm_dataset (outdoor layout, keypoint "
                    f"thr 1e-4, no keypoint cap, batch 4, 20 iterations, masks, SuperPoint fp32, SuperGlue {precision}) -> "
                    f"ColmapDatabaseWriter (sqlite); the reference call site is 200 frames / 19 900 pairs",
            "frames": n_frames, "pairs": n_pairs, "keypoints_per_image_min_max": [min(kp), max(kp)],
            "extract_and_match_s": t_match, "database_write_s": t_db, "pairs_per_s_wall": n_pairs / (t_match + t_db),
            "matches": int(sum(len(v) for v in table.values()))}


def config4(dev, precision, rank, world, kp=4096, n_img=19):
    from simplesfm_b200 import synthetic
    from simplesfm_b200.matchers import Matcher
    from simplesfm_b200.matchers.matcher import Frame
    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(1, "outdoor")
    m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                sinkhorn_iterations=20, super_glue_batch=5, precision=precision, device=dev, fuse_chunks=51)
    img = torch.zeros(1, 480, 640, device=dev)
    frames = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in synthetic.feature_scene(n_img, kp, seed=5 + rank)]
    names = {i + 1: str(i) for i in range(n_img)}
    run = lambda: m.match_frames(frames, names, None)[1]
    run()
    ms, table = _sync_ms(run, dev, reps=2)
    ms = _max_over_ranks(ms, dev, world)
    n_pairs = n_img * (n_img - 1) // 2
    return {"what": f"{n_img}-image scene per GPU ({n_pairs} pairs), {kp} keypoints, outdoor layout, batch 5, 20 "
                    f"iterations, train-mode BN, SuperGlue {precision}; the full config is 1000 images / 499 500 pairs "
                    f"over 8 GPUs",
            "pairs_per_s_per_gpu": n_pairs / ms * 1e3, "pairs_per_s_all_gpus": world * n_pairs / ms * 1e3,
            "gflop_per_pair": 823.2, "tflops_per_gpu": n_pairs / ms * 1e3 * 823.2e9 / 1e12,
            "matches": int(sum(len(v) for v in table.values()))}


def eval_mode(dev, precision, kp=2048, n_img=30):
    """The headline workload with `bn_mode='eval'` (running statistics folded into the weights; what the reference
    computes after `.eval()`): BatchNorm no longer couples the pairs of a chunk, so the keypoint encoder and GNN
    layer 0 run once per image (layer-0 hoisting) instead of once per pair."""
    from simplesfm_b200 import synthetic
    from simplesfm_b200.matchers import Matcher
    from simplesfm_b200.matchers.matcher import Frame
    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0, "indoor")
    img = torch.zeros(1, 480, 640, device=dev)
    frames = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in synthetic.feature_scene(n_img, kp, seed=3)]
    names = {i + 1: str(i) for i in range(n_img)}
    n_pairs = n_img * (n_img - 1) // 2
    out = {"what": f"{n_img}-image scene ({n_pairs} pairs), {kp} keypoints, indoor layout, batch 5, 20 iterations, "
                   f"SuperGlue {precision}: train-mode BatchNorm (the reference's behaviour) vs eval-mode BatchNorm "
                   f"with layer-0 hoisting"}
    for mode in ("train", "eval"):
        m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                    sinkhorn_iterations=20, super_glue_batch=5, precision=precision, device=dev, fuse_chunks=51,
                    bn_mode=mode)
        run = lambda: m.match_frames(frames, names, None)[1]
        run()
        ms, table = _sync_ms(run, dev, reps=2)
        out[f"pairs_per_s_{mode}"] = n_pairs / ms * 1e3
        out[f"matches_{mode}"] = int(sum(len(v) for v in table.values()))
        del m
        torch.cuda.empty_cache()
    return out


def config5(dev, rank, world, peaks, N=4096, iters=100, total=256):
    from simplesfm_b200 import _lib
    L = _lib.lib()
    B = max(1, total // world)
    g = torch.Generator(device=dev).manual_seed(rank)
    S = torch.randn(B, N, N, generator=g, device=dev) * 2.0
    idx = torch.arange(N // 2, device=dev)
    for b in range(B):
        S[b, idx, torch.randperm(N, generator=g, device=dev)[:N // 2]] += 20.0
    ws = torch.empty(L.sfm_sinkhorn_half_workspace_bytes(B, N, N), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    res = {"what": f"log_optimal_transport + mutual matching, N = M = {N}, {iters} iterations, {total} pairs split over "
                   f"{world} rank(s) ({B} per GPU); GB/s = algorithmic matrix bytes / time, against hbm_gbs = {hbm}",
           "batch_per_gpu": B}
    keep = {}
    for name, fn, bytes_per_el in (("exact_logspace_fp32", L.sfm_sinkhorn_match, (2 * iters + 3) * 4),
                                   ("scaling_form_fp32_storage", L.sfm_sinkhorn_match_scaled, (iters + 3) * 4),
                                   ("scaling_form_fp16_storage", L.sfm_sinkhorn_match_scaled_half, iters * 2 + 4 + 2 + 4)):
        m0 = torch.empty(B, N, dtype=torch.int64, device=dev)
        m1 = torch.empty(B, N, dtype=torch.int64, device=dev)
        s0, s1 = torch.empty(B, N, device=dev), torch.empty(B, N, device=dev)

        def run(Sx, nb):
            _lib.check(fn(Sx.data_ptr(), nb, N, N, 1.0, iters, 0.2, m0.data_ptr(), m1.data_ptr(), s0.data_ptr(),
                          s1.data_ptr(), ws.data_ptr(), ws.numel(), st))
        run(S[:1].clone(), 1)                                   # kernel load
        Sx = S.clone() if name == "scaling_form_fp32_storage" else S       # that entry point overwrites its input
        ms, _ = _sync_ms(lambda: run(Sx, B), dev)
        ms = _max_over_ranks(ms, dev, world)
        gbs = bytes_per_el * B * N * N / ms / 1e6
        keep[name] = m0.clone()
        res[name] = {"ms": ms, "pairs_per_s_all_gpus": world * B / ms * 1e3, "GBps_per_gpu": gbs, "frac_of_hbm_peak": gbs / hbm,
                     "matches_per_pair": float((m0 > -1).sum()) / B}
        del Sx
    a = keep["exact_logspace_fp32"]
    for name in ("scaling_form_fp32_storage", "scaling_form_fp16_storage"):
        b = keep[name]
        both = (a > -1) | (b > -1)
        res[name]["match_agreement_with_exact"] = float(((a == b) & both).sum()) / max(int(both.sum()), 1)
    return res


def run(a, dev, rank, world, peaks):
    out = {}
    steps = [("config4", lambda: config4(dev, a.precision, rank, world)),
             ("config5", lambda: config5(dev, rank, world, peaks))]
    if world == 1:
        steps = [("config1", lambda: config1(dev)), ("config3", lambda: config3(dev, a.precision)),
                 ("eval_mode", lambda: eval_mode(dev, a.precision))] + steps
    for name, fn in steps:
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:        # a failed secondary measurement must not lose the headline line
            if world > 1:
                raise                  # ... but ranks must not drift apart inside collectives
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        if isinstance(out[name], dict):
            out[name]["seconds"] = time.perf_counter() - t0
        torch.cuda.empty_cache()
    return out
