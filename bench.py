#!/usr/bin/env python
"""Benchmark of the hot path: SuperGlue exhaustive matching (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = matching every pair of ONE synthetic scene (2048 keypoints per image, 640x480, indoor-layout random
weights) through `Matcher.match_frames`, exactly as the reference's `Matcher.__super_glue_matcher` would (chunks of
`super_glue_batch` pairs, Matcher defaults: 20 Sinkhorn iterations, train-mode BatchNorm as the reference runs it).
  N = 1 : the 50-image scene of BASELINE config 2 (1225 pairs).
  N > 1 : the north_star partition on ONE scene -- images dealt to the ranks for SuperPoint, one
          all-gather of the feature records, the lexicographic pair list cut into chunk-aligned ranges, compact
          uint32 match tables gathered to rank 0.  The scene grows with N so that every GPU keeps about 1225 pairs
          (70 / 99 / 140 images at N = 2 / 4 / 8): "scaling": "weak", value = all pairs of the scene / time.
  value        : pairs/s with the feature records resident in HBM (device timing, barrier on both sides, max over
                 ranks; includes the table gather to rank 0 and its device->host copy)
  e2e          : pairs/s through the same public call with HOST (pinned) features -> H2D every step, tables D2H
  e2e_images   : pairs/s from HOST IMAGES: 8-bit pixels H2D -> SuperPoint (this rank's images) -> feature all-gather
                 -> SuperGlue (this rank's pairs) -> table gather, everything inside the timed region
  roofline     : dominant kernel (attention) -- algorithmic FLOPs / CUDA-event time measured inside the timed
                 region, against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port of the reference (oracle/sfm_oracle.py) on one 5-pair chunk of the workload
  table_sha256 : hash of the merged table of the 50-image scene matched with this run's rank count -- identical at
                 any N (chunk-aligned sharding keeps BatchNorm statistics and truncation of every chunk unchanged)
  configs      : short secondary measurements of the other BASELINE.json configs (1, 3, 4, 5)
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "superglue_pairs_per_sec_2048kp_640x480"
MATCH_THR = 0.4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SFM_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--images", type=int, default=0, help="images in the scene (0 = about 1225 pairs per GPU)")
    ap.add_argument("--kp", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=5, help="super_glue_batch (Matcher default 5)")
    ap.add_argument("--iters", type=int, default=20, help="sinkhorn_iterations (Matcher default 20)")
    ap.add_argument("--bn", default="train", choices=["train", "eval"])
    ap.add_argument("--cpu-pairs", type=int, default=5, help="pairs in the bounded CPU sample (one Matcher chunk)")
    ap.add_argument("--fuse-chunks", type=int, default=51, help="Matcher chunks launched together (BN stats stay per chunk)")
    ap.add_argument("--sp-precision", default="fp32", choices=["fp32", "bf16"], help="SuperPoint precision of e2e_images")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configs")
    ap.add_argument("--no-images", action="store_true", help="skip the e2e_images leg")
    return ap.parse_args()


def images_for(world):
    """Scene size with about 1225 pairs per GPU: n (n - 1) / 2 ~ 1225 * world."""
    if world <= 1:
        return 50
    return int(round((1 + math.sqrt(1 + 8 * 1225 * world)) / 2 - 0.5))


def n_images(a, world):
    return a.images if a.images > 0 else images_for(world)


def config_dict(a, world):
    n = n_images(a, world)
    return {"workload": (f"exhaustive matching of a {n}-image synthetic scene ({n * (n - 1) // 2} pairs), {a.kp} "
                         f"keypoints, 640x480, indoor-layout random weights"),
            "sinkhorn_iterations": a.iters, "bn_mode": a.bn, "super_glue_batch": a.batch, "match_threshold": MATCH_THR}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(a, steps, warmup):
    """Time the oracle port of the reference on the host cores: one Matcher chunk (`cpu_pairs` pairs at batch
    `cpu_pairs`) of the same workload per step.  Returns (cpu_baseline dict, seconds per step, oracle table)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sfm_oracle as O
    from simplesfm_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = synthetic.superglue_state_dict(0, "indoor")
    fr = synthetic.feature_scene(a.cpu_pairs + 1, a.kp, seed=0)
    frames = [(d, k, s, torch.zeros(1, 480, 640)) for d, k, s in fr]
    pairs = np.array([(1, j + 2) for j in range(a.cpu_pairs)])
    times, table = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            table = O.superglue_match_table(w, frames, pairs, a.cpu_pairs, a.iters, MATCH_THR, a.bn)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    dt = sorted(times)[len(times) // 2]
    return {"value": a.cpu_pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{a.cpu_pairs} pairs (one Matcher chunk at super_glue_batch {a.cpu_pairs}) of the same "
                      f"workload, median of {len(times)} runs, {a.iters} Sinkhorn iterations, bn={a.bn}, fp32, "
                      f"torch CPU ops with {cores} threads"}, dt, (table, fr, pairs)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    a.cpu_pairs = max(a.cpu_pairs, a.batch)          # one full chunk at the Matcher's own batch size
    steps, warmup = max(1, a.steps), min(a.warmup, 1)
    cb, dt, _ = cpu_sample(a, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(a, world),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5,
                                     stdin=subprocess.DEVNULL).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, n in enumerate(names):
                if len(s) > 2 + i and s[2 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def table_hash(table):
    """sha256 over (id1, id2, L, rows) of every pair in pair-list order."""
    import numpy as np
    h = hashlib.sha256()
    for (a, b), rows in table.items():
        h.update(np.array([int(a), int(b), len(rows)], dtype=np.int64).tobytes())
        h.update(np.ascontiguousarray(rows, dtype=np.uint32).tobytes())
    return h.hexdigest()


def set_agreement(ta, tb):
    inter = union = 0
    for k in ta:
        sa = {tuple(r) for r in ta[k].tolist()}
        sb = {tuple(r) for r in tb[k].tolist()}
        inter += len(sa & sb)
        union += len(sa | sb)
    return inter / max(union, 1)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(a):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from simplesfm_b200 import _lib, synthetic
    from simplesfm_b200.matchers import Matcher
    from simplesfm_b200.matchers.matcher import Frame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = _lib.lib()
    n_img = n_images(a, world)
    n_pairs = n_img * (n_img - 1) // 2
    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0, "indoor")

    def make_matcher(precision=a.precision, sp_precision=a.sp_precision, shard=world > 1):
        return Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw,
                       match_threshold=MATCH_THR, sinkhorn_iterations=a.iters, super_glue_batch=a.batch, bn_mode=a.bn,
                       precision=precision, device=dev, shard_pairs=shard, fuse_chunks=a.fuse_chunks,
                       max_keypoints=a.kp, superpoint_precision=sp_precision, extract_batch=8)

    m = make_matcher()
    scene = synthetic.feature_scene(n_img, a.kp, seed=0)            # the same scene on every rank
    host = [(d.pin_memory(), k.pin_memory(), s.pin_memory()) for d, k, s in scene]
    img = torch.zeros(1, 480, 640, device=dev)
    resident = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in host]
    names = {i + 1: f"{i:04d}.png" for i in range(n_img)}
    h2d = sum(d.numel() * 4 + k.numel() * 4 + s.numel() * 4 for d, k, s in host)

    def step_resident():
        return m.match_frames(resident, names, None)[1]

    def step_e2e():
        frames = [Frame(d.to(dev, non_blocking=True), k.to(dev, non_blocking=True), s.to(dev, non_blocking=True), img)
                  for d, k, s in host]
        return m.match_frames(frames, names, None)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, out

    def total(x):
        if world == 1:
            return x
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(tt)
        return tt.item()

    for _ in range(a.warmup):
        step_resident()
    # ---- timed region: resident inputs, profiling events on
    sampler = ClockSampler(local)
    sampler.start()
    check = _lib.check
    check(L.sfm_set_profiling(m.matcher.handle.h, 1))
    launches0 = L.sfm_launch_count()
    ms, table = timed(step_resident, a.steps)
    launches = int(total(L.sfm_launch_count() - launches0))
    prof_ms = (C.c_float * 4)()
    prof_n = (C.c_longlong * 4)()
    check(L.sfm_get_profile(m.matcher.handle.h, prof_ms, prof_n))
    check(L.sfm_set_profiling(m.matcher.handle.h, 0))
    clocks = sampler.stop()
    # ---- e2e: host features -> H2D -> match -> D2H tables
    step_e2e()
    ms_e2e, table_e2e = timed(step_e2e, a.steps)
    # bytes of match tables delivered to the host of rank 0 (every pair's table; rows travel as uint32)
    d2h = sum(v.nbytes for v in table_e2e.values()) if rank == 0 else 0
    value = n_pairs * a.steps / (ms / 1e3)
    e2e = n_pairs * a.steps / (ms_e2e / 1e3)

    # ---- e2e from host images: pixels H2D -> SuperPoint (own images) -> all-gather -> SuperGlue -> gather
    e2e_images = None
    if not a.no_images:
        own = [i for i in range(n_img) if i % world == rank]
        # overlapping views of ONE textured scene (windows of the same texture, shifted), so that the matches are real
        # correspondences and the detector-precision comparison below means something
        pix = {i: (synthetic.textured_image(1000, shift=(7 * (i % 15) - 49, 11 * ((i * 7) % 11) - 55))[0, 0] * 255)
               .round().to(torch.uint8).numpy() for i in own}
        items = [(pix.get(i), None) for i in range(n_img)]          # a rank only touches the images it extracts

        def step_images():
            frames = m._frames_from(iter(items))
            return frames, m.match_frames(frames, names, None)[1]

        step_images()
        ms_img, (frames_img, table_img) = timed(step_images, max(1, min(a.steps, 5)))
        counts = [int(f.keypoints_poses.shape[0]) for f in frames_img]
        e2e_images = {"value": n_pairs * max(1, min(a.steps, 5)) / (ms_img / 1e3), "unit": "pairs/s",
                      "ms_per_step": ms_img / max(1, min(a.steps, 5)), "superpoint_precision": a.sp_precision,
                      "h2d_bytes_per_step": n_img * 480 * 640, "keypoints_per_image_min_max": [min(counts), max(counts)],
                      "includes": "8-bit pixels H2D, SuperPoint on this rank's images, feature all-gather, "
                                  "SuperGlue on this rank's pairs, table gather to rank 0, D2H"}
        if rank == 0:
            e2e_images["matches_per_step"] = int(sum(len(v) for v in table_img.values()))
        # the same pipeline with the tcgen05 (bf16) SuperPoint: stated precision, a slightly different detector --
        # its end-to-end agreement with the exact-SuperPoint tables is measured here through keypoint coordinates
        other = "bf16" if a.sp_precision == "fp32" else "fp32"
        m_alt = make_matcher(sp_precision=other)

        def step_images_alt():
            frames = m_alt._frames_from(iter(items))
            return frames, m_alt.match_frames(frames, names, None)[1]

        step_images_alt()
        n_alt = max(1, min(a.steps, 5))
        ms_alt, (frames_alt, table_alt) = timed(step_images_alt, n_alt)
        alt = {"value": n_pairs * n_alt / (ms_alt / 1e3), "unit": "pairs/s", "ms_per_step": ms_alt / n_alt,
               "superpoint_precision": other}
        if rank == 0:
            kp_a = [f.keypoints_poses.cpu().numpy() for f in frames_img]
            kp_b = [f.keypoints_poses.cpu().numpy() for f in frames_alt]
            same_kp = sum(len({tuple(x) for x in ka.tolist()} & {tuple(x) for x in kb.tolist()}) for ka, kb in zip(kp_a, kp_b))
            inter = union = 0
            for k in list(table_img.keys())[:100]:
                i, j = int(k[0]) - 1, int(k[1]) - 1
                sa = {(tuple(kp_a[i][r[0]]), tuple(kp_a[j][r[1]])) for r in table_img[k].tolist()}
                sb = {(tuple(kp_b[i][r[0]]), tuple(kp_b[j][r[1]])) for r in table_alt[k].tolist()}
                inter += len(sa & sb)
                union += len(sa | sb)
            # descriptors of the keypoints both detectors found (first 8 images): cosine similarity
            cos = []
            for fa, fb, ka, kb in list(zip(frames_img, frames_alt, kp_a, kp_b))[:8]:
                ib = {tuple(x): j for j, x in enumerate(kb.tolist())}
                pairs_ab = [(i, ib[tuple(x)]) for i, x in enumerate(ka.tolist()) if tuple(x) in ib]
                if pairs_ab:
                    ia = torch.tensor([p[0] for p in pairs_ab], device=dev)
                    jb = torch.tensor([p[1] for p in pairs_ab], device=dev)
                    cos.append((fa.descriptors[:, ia] * fb.descriptors[:, jb]).sum(0))
            if cos:
                cos = torch.cat(cos)
                alt["descriptor_cosine_mean_min_on_common_keypoints"] = [float(cos.mean()), float(cos.min())]
            alt["note"] = ("random synthetic SuperPoint weights give few, near-threshold matches per image pair "
                           "(matches_per_step), so the table agreement of two detectors is dominated by threshold "
                           "flips; keypoint overlap and descriptor cosine are the informative figures")
            alt["keypoint_overlap_with_" + a.sp_precision] = same_kp / max(sum(len(k) for k in kp_a), 1)
            alt["match_table_agreement_with_" + a.sp_precision] = inter / max(union, 1)
            alt["agreement_pairs"] = min(100, len(table_img))
        e2e_images["alternative_superpoint"] = alt
        del frames_img, table_img, frames_alt, table_alt, m_alt

    # ---- determinism across rank counts: the 50-image scene matched with THIS world size
    sub_n = min(50, n_img)
    sub_names = {i + 1: names[i + 1] for i in range(sub_n)}
    sub_table = table if sub_n == n_img else m.match_frames(resident[:sub_n], sub_names, None)[1]
    sha = table_hash(sub_table) if rank == 0 else None
    sharded_equals_single = None
    if world > 1:
        # direct check on rank 0: the same 50-image scene matched WITHOUT sharding on this GPU
        if rank == 0:
            solo = make_matcher(shard=False).match_frames(resident[:sub_n], sub_names, None)[1]
            bad = [k for k in solo if k not in sub_table or not np.array_equal(solo[k], sub_table[k])]
            sharded_equals_single = len(bad) == 0 and len(solo) == len(sub_table)
            if bad:
                print(f"bench.py: {len(bad)} of {len(solo)} pairs differ between the sharded and the single-GPU run, "
                      f"first {bad[:3]} ({[(len(solo[k]), len(sub_table.get(k, []))) for k in bad[:3]]})", file=sys.stderr)
            sha_single = table_hash(solo)
        barrier()

    # ---- roofline of the dominant kernel (attention), from this rank's launches
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    pk = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    if "bf16_tflops_sustained" in pk:
        peak, peak_src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    else:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
    from simplesfm_b200.distributed import shard_range
    lo, hi = shard_range(n_pairs, a.batch, rank, world)
    my_pairs = hi - lo
    # dominant kernel = fused attention.  Algorithmic FLOPs: per pair and layer
    # 2 sides x (QK^T + PV) = 2 x 4*N*M*64 per head x 4 heads = 2048*N*M, 18 layers.
    att_groups = max(int(prof_n[1]), 1)
    att_total_ms = float(prof_ms[1])
    att_flop_total = float(a.steps) * my_pairs * 18 * 2048.0 * a.kp * a.kp
    achieved = att_flop_total / (att_total_ms * 1e-3) / 1e12 if att_total_ms > 0 else 0.0
    flop_group = att_flop_total / att_groups
    traffic, traffic_src = None, None
    for name in ("r2f_attention_traffic.json", "r2_attention_traffic.json", "r1_attention_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # dram bytes of one ncu-captured launch, scaled by FLOPs to this run's launch size
            traffic = tj["dram_bytes_per_launch"] * flop_group / tj["flop_per_launch"]
            traffic_src = f"profiles/{name} (ncu --set full capture of one launch, scaled by FLOPs to this launch size)"
            break
    roofline = {"bound": "tensor", "kernel": "tc_attention_kernel (QK^T, online softmax, PV; both sides of one layer)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "flop_per_launch": flop_group, "ms_per_launch": att_total_ms / att_groups,
                "launches": att_groups, "rank": rank,
                # the kernel's own bound at head dim 64: one ex2 per 256 tensor FLOPs, 16 ex2 / clk / SM on the MUFU
                # pipe (tools/micro/exp_mix_bench.cu) -> 148 SMs x 16 x clock x 256 FLOP
                "mufu_bound": (lambda mhz: None if not mhz else {
                    "peak_tflops_at_run_clock": 148 * 16 * mhz * 1e6 * 256 / 1e12, "sm_mhz": mhz,
                    "frac": achieved / (148 * 16 * mhz * 1e6 * 256 / 1e12)})(clocks.get("sm_mhz")),
                "share_of_step": {k: float(prof_ms[i]) / ms for i, k in
                                  enumerate(["linear", "attention", "sinkhorn_match", "other"])}}

    # ---- stated-precision checks (rank 0 of a single-GPU run): bf16 vs the exact fp32 path, bf16 vs the ORACLE
    parity = None
    cpu_baseline = None
    if world == 1 and a.precision == "bf16":
        ref_m = make_matcher(precision="fp32", shard=False)
        sub = {i + 1: names[i + 1] for i in range(min(4, n_img))}
        t32 = ref_m.match_frames(resident[:len(sub)], sub, None)[1]
        t16 = m.match_frames(resident[:len(sub)], sub, None)[1]
        parity = {"bf16_vs_fp32_match_set_agreement": set_agreement(t32, t16), "pairs": len(t32),
                  "reference_matches": int(sum(len(v) for v in t32.values()))}
        del ref_m
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        a.cpu_pairs = max(a.cpu_pairs, a.batch)
        cpu_baseline, _, (otable, ofr, opairs) = cpu_sample(a, 2, 1)
        # the same chunk through the GPU path (the oracle's features are the scene's first images)
        oframes = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in ofr]
        onames = {i + 1: str(i) for i in range(len(oframes))}
        mine = m.match_frames(oframes, onames, lambda p: p[0] == 1)[1]
        if parity is None:
            parity = {}
        parity["vs_oracle_match_set_agreement"] = set_agreement({(int(k[0]), int(k[1])): v for k, v in otable.items()},
                                                                {(int(k[0]), int(k[1])): v for k, v in mine.items()})
        parity["vs_oracle_pairs"] = len(otable)
        parity["vs_oracle_matches"] = int(sum(len(v) for v in otable.values()))
        parity["precision"] = a.precision

    cfg = config_dict(a, world)
    cfg.update({"pairs_per_step": n_pairs, "pairs_per_step_per_gpu": n_pairs / world, "precision": a.precision,
                "fused_chunks_per_launch": a.fuse_chunks,
                "partition": ("single GPU" if world == 1 else
                              f"one scene over {world} GPUs: pair list in chunk-aligned contiguous ranges, SuperPoint by "
                              f"image (i mod {world}), one feature all-gather, compact tables gathered to rank 0"),
                "l2": "working set exceeds L2 (feature records >= 105 MB + per-chunk activations)"})
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
            "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps},
            "e2e_images": e2e_images,
            "gpu_launches": launches, "roofline": roofline,
            "gflop_per_pair": 254.8 if a.kp == 2048 else None,
            "tensor_frac_whole_step": (value / world * 254.8e9 / 1e12) / peak if a.kp == 2048 else None,
            "matches_per_step": int(sum(len(v) for v in table.values())) if rank == 0 else None,
            "table_sha256_50_image_scene": sha, "sharded_table_equals_single_gpu": sharded_equals_single,
            "parity": parity}
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if not a.no_configs:
        del resident, host, scene
        torch.cuda.empty_cache()
        from tools import bench_configs
        line["configs"] = bench_configs.run(a, dev, rank, world, peaks=pk)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
