"""Diagnose bf16-vs-fp32 match flips: where do disagreements sit relative to the threshold / top-2 margin?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from simplesfm_b200 import synthetic
from simplesfm_b200.models.superglue import SuperGlue
dev = torch.device("cuda:0")
sgw = synthetic.superglue_state_dict(0)
for (B, K, seed, thr) in ((4, 1024, 9, 0.2), (5, 2048, 0, 0.4), (4, 1024, 3, 0.2)):
    fr = synthetic.feature_scene(2 * B, K, seed=seed)
    data = {"descriptors0": torch.stack([fr[i][0] for i in range(B)]).to(dev),
            "descriptors1": torch.stack([fr[B + i][0] for i in range(B)]).to(dev),
            "keypoints0": torch.stack([fr[i][1] for i in range(B)]).to(dev),
            "keypoints1": torch.stack([fr[B + i][1] for i in range(B)]).to(dev),
            "scores0": torch.stack([fr[i][2] for i in range(B)]).to(dev),
            "scores1": torch.stack([fr[B + i][2] for i in range(B)]).to(dev),
            "image0": torch.zeros(B, 1, 480, 640), "image1": torch.zeros(B, 1, 480, 640)}
    for mode in ("train", "eval"):
        f = SuperGlue(sgw, sinkhorn_iterations=20, match_threshold=thr, bn_mode=mode, precision="fp32", device=dev)
        ref = f(data)
        ws = f.handle.workspace("sg", 0)
        S32 = f.tap(B, K, K, 2, ws).view(B, K, K).clone()
        h = SuperGlue(sgw, sinkhorn_iterations=20, match_threshold=thr, bn_mode=mode, precision="bf16", device=dev)
        out = h(data)
        S16 = h.tap(B, K, K, 2, h.handle.workspace("sg", 0)).view(B, K, K).clone()
        a, b = out["matches0"].cpu().numpy(), ref["matches0"].cpu().numpy()
        both = (a > -1) | (b > -1)
        flips = np.argwhere((a != b) & both)
        sc = ref["matching_scores0"].cpu().numpy()
        sc16 = out["matching_scores0"].cpu().numpy()
        near = sum(1 for (bi, i) in flips if abs(sc[bi, i] - thr) < 0.05 * thr or abs(sc16[bi, i] - thr) < 0.05 * thr)
        print(f"B={B} K={K} {mode}: matches {int((b>-1).sum())}, agreement {1 - len(flips)/max(both.sum(),1):.4f}, flips {len(flips)}, "
              f"near-threshold(5%) {near}; score-matrix rel err {float((S16-S32).abs().max()/S32.abs().max()):.4f} "
              f"rms {float((S16-S32).pow(2).mean().sqrt()/S32.pow(2).mean().sqrt()):.4f}")
        for (bi, i) in flips[:6]:
            print("   flip", bi, i, "ref", b[bi, i], f"{sc[bi,i]:.4f}", "bf16", a[bi, i], f"{sc16[bi,i]:.4f}")
