"""Smallest run that touches every kernel added or changed in round 2 (for compute-sanitizer memcheck):
bf16 SuperGlue forward (lagged-maximum attention + recomputation kernel, ragged keypoint counts), hoisted forward,
general log_sinkhorn_iterations, table packing, SimpleMatcher.match_pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplesfm_b200 import synthetic
from simplesfm_b200.matchers import Matcher
from simplesfm_b200.matchers.matcher import Frame
from simplesfm_b200.models import superglue as SG
dev = torch.device("cuda:0")
spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0)
fr = synthetic.feature_scene(4, 200, seed=1)
counts = [200, 173, 200, 96]
img = torch.zeros(1, 480, 640, device=dev)
frames = [Frame(d[:, :c].to(dev), k[:c].to(dev), s[:c].to(dev), img) for (d, k, s), c in zip(fr, counts)]
names = {i + 1: str(i) for i in range(4)}
for kw in (dict(precision="bf16"), dict(precision="bf16", bn_mode="eval"), dict(precision="fp32")):
    m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.2,
                sinkhorn_iterations=5, super_glue_batch=2, device=dev, **kw)
    t = m.match_frames(frames, names, None)[1]
    print(kw, sum(len(v) for v in t.values()), "matches")
same = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in fr]
m = Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.2,
            sinkhorn_iterations=5, super_glue_batch=2, device=dev, precision="bf16", bn_mode="eval")
print("hoisted", sum(len(v) for v in m.match_frames(same, names, None)[1].values()), "matches")
m = Matcher(spw, 4, 0.005, matcher_type="simple", match_threshold=0.9, device=dev)
print("simple", sum(len(v) for v in m.match_frames(frames, names, None)[1].values()), "matches")
g = torch.Generator().manual_seed(0)
Z = torch.randn(2, 37, 53, generator=g).to(dev)
out = SG.log_sinkhorn_iterations(Z, torch.log_softmax(torch.randn(2, 37, generator=g), -1).to(dev),
                                 torch.log_softmax(torch.randn(2, 53, generator=g), -1).to(dev), 7)
torch.cuda.synchronize()
print("log_sinkhorn_iterations", tuple(out.shape), bool(torch.isfinite(out).all()))
