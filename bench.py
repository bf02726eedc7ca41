#!/usr/bin/env python
"""Benchmark of the hot path: SuperGlue exhaustive matching (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = matching every pair of a 50-image synthetic scene (1225 pairs, 2048 keypoints per
image, 640x480, indoor-layout random weights) through `Matcher.match_frames`, exactly as the
reference's `Matcher.__super_glue_matcher` would (chunks of `super_glue_batch` pairs, Matcher
defaults: 20 Sinkhorn iterations, train-mode BatchNorm as the reference runs it).
  value : pairs/s with the feature records already resident in HBM (device timing, max over ranks)
  e2e   : pairs/s through the same public call with HOST (pinned) features -> H2D every step and
          match tables D2H inside the timed region
  roofline     : dominant kernel (attention) -- algorithmic FLOPs / CUDA-event time measured
                 inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port of the reference (oracle/sfm_oracle.py) on a bounded sample
Multi-GPU: weak scaling -- every rank matches its own scene; no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "superglue_pairs_per_sec_2048kp_640x480"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SFM_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--images", type=int, default=50)
    ap.add_argument("--kp", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=5, help="super_glue_batch (Matcher default 5)")
    ap.add_argument("--iters", type=int, default=20, help="sinkhorn_iterations (Matcher default 20)")
    ap.add_argument("--bn", default="train", choices=["train", "eval"])
    ap.add_argument("--cpu-pairs", type=int, default=2, help="pairs in the bounded CPU sample")
    ap.add_argument("--fuse-chunks", type=int, default=32, help="Matcher chunks launched together (BN stats stay per chunk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    n_pairs = a.images * (a.images - 1) // 2
    return (f"exhaustive matching of a {a.images}-image synthetic scene ({n_pairs} pairs), {a.kp} keypoints, "
            f"640x480, indoor-layout random weights")


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(a, steps, warmup):
    """Time the oracle port of the reference on the host cores: one chunk of `cpu_pairs` pairs of the
    same workload per step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sfm_oracle as O
    from simplesfm_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = synthetic.superglue_state_dict(0, "indoor")
    fr = synthetic.feature_scene(a.cpu_pairs + 1, a.kp, seed=0)
    frames = [(d, k, s, torch.zeros(1, 480, 640)) for d, k, s in fr]
    import numpy as np
    pairs = np.array([(1, j + 2) for j in range(a.cpu_pairs)])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.superglue_match_table(w, frames, pairs, a.cpu_pairs, a.iters, 0.4, a.bn)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    dt = sorted(times)[len(times) // 2]
    return {"value": a.cpu_pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{a.cpu_pairs} pairs (one chunk) of the same workload, median of {len(times)} runs, "
                      f"{a.iters} Sinkhorn iterations, bn={a.bn}, fp32, torch CPU ops with {cores} threads"}, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), min(a.warmup, 1)
    cb, dt = cpu_sample(a, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "sinkhorn_iterations": a.iters, "bn_mode": a.bn,
                       "super_glue_batch": a.cpu_pairs},
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
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
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
    L.sfm_launch_count.restype = C.c_longlong
    L.sfm_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.sfm_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]

    spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0, "indoor")
    m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                sinkhorn_iterations=a.iters, super_glue_batch=a.batch, bn_mode=a.bn, precision=a.precision,
                device=dev, shard_pairs=False, fuse_chunks=a.fuse_chunks)
    scene = synthetic.feature_scene(a.images, a.kp, seed=rank)
    host = [(d.pin_memory(), k.pin_memory(), s.pin_memory()) for d, k, s in scene]
    img = torch.zeros(1, 480, 640, device=dev)
    resident = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in host]
    names = {i + 1: f"{i:04d}.png" for i in range(a.images)}
    n_pairs = a.images * (a.images - 1) // 2
    h2d = sum(d.numel() * 4 + k.numel() * 4 + s.numel() * 4 for d, k, s in host)

    def step_resident():
        return m.match_frames(resident, names, lambda p: True)[1]

    def step_e2e():
        frames = [Frame(d.to(dev, non_blocking=True), k.to(dev, non_blocking=True), s.to(dev, non_blocking=True), img)
                  for d, k, s in host]
        return m.match_frames(frames, names, lambda p: True)[1]

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

    for _ in range(a.warmup):
        step_resident()
    # ---- timed region: resident inputs, profiling events on
    sampler = ClockSampler(local)
    sampler.start()
    check = _lib.check
    check(L.sfm_set_profiling(m.matcher.handle.h, 1))
    launches0 = L.sfm_launch_count()
    ms, table = timed(step_resident, a.steps)
    launches = L.sfm_launch_count() - launches0
    prof_ms = (C.c_float * 4)()
    prof_n = (C.c_longlong * 4)()
    check(L.sfm_get_profile(m.matcher.handle.h, prof_ms, prof_n))
    check(L.sfm_set_profiling(m.matcher.handle.h, 0))
    clocks = sampler.stop()
    # ---- e2e: host features -> H2D -> match -> D2H tables
    step_e2e()
    ms_e2e, table_e2e = timed(step_e2e, a.steps)
    d2h = sum(v.nbytes for v in table_e2e.values())

    value = world * n_pairs * a.steps / (ms / 1e3)
    e2e = world * n_pairs * a.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (attention), per launch group
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    else:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
    # dominant kernel = fused attention.  Algorithmic FLOPs: per pair and layer
    # 2 sides x (QK^T + PV) = 2 x 4*N*M*64 per head x 4 heads = 2048*N*M, 18 layers.
    att_groups = max(int(prof_n[1]), 1)
    att_total_ms = float(prof_ms[1])
    att_flop_total = float(a.steps) * n_pairs * 18 * 2048.0 * a.kp * a.kp
    att_ms = att_total_ms / att_groups
    flop_group = att_flop_total / att_groups
    achieved = att_flop_total / (att_total_ms * 1e-3) / 1e12 if att_total_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_attention_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        # dram bytes of one ncu-captured launch, scaled by FLOPs to this run's launch size
        traffic = tj["dram_bytes_per_launch"] * flop_group / tj["flop_per_launch"]
    roofline = {"bound": "tensor", "kernel": "tc_attention_kernel (QK^T, online softmax, PV; both sides of one layer)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "flop_per_launch": flop_group, "ms_per_launch": att_ms,
                "launches": att_groups,
                "share_of_step": {k: float(prof_ms[i]) / ms for i, k in
                                  enumerate(["linear", "attention", "sinkhorn_match", "other"])}}

    # stated-precision check: bf16 vs the fp32 exact path on one chunk of this scene
    agreement = None
    if a.precision == "bf16":
        ref_m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                        sinkhorn_iterations=a.iters, super_glue_batch=a.batch, bn_mode=a.bn, precision="fp32",
                        device=dev, shard_pairs=False)
        sub = {i + 1: names[i + 1] for i in range(min(4, a.images))}
        t32 = ref_m.match_frames(resident[:len(sub)], sub, lambda p: True)[1]
        t16 = m.match_frames(resident[:len(sub)], sub, lambda p: True)[1]
        inter = union = 0
        for k in t32:
            s32 = {tuple(r) for r in t32[k].tolist()}
            s16 = {tuple(r) for r in t16[k].tolist()}
            inter += len(s32 & s16)
            union += len(s32 | s16)
        agreement = {"bf16_vs_fp32_match_set_agreement": inter / max(union, 1), "pairs": len(t32),
                     "reference_matches": int(sum(len(v) for v in t32.values()))}
        del ref_m

    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "pairs_per_step_per_gpu": n_pairs,
                       "sinkhorn_iterations": a.iters, "bn_mode": a.bn, "super_glue_batch": a.batch,
                       "precision": a.precision, "match_threshold": 0.4, "fused_chunks_per_launch": a.fuse_chunks,
                       "l2": "working set exceeds L2 (feature records 105 MB + per-chunk activations)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "gflop_per_pair": 254.8 if a.kp == 2048 else None,
            "tensor_frac_whole_step": (value / world * 254.8e9 / 1e12) / peak if a.kp == 2048 else None,
            "matches_per_step": int(sum(len(v) for v in table.values())), "parity": agreement}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_sample(a, 2, 1)
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
