"""One SuperGlue forward of a bench-shaped chunk (B pairs x K keypoints) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import synthetic
from simplesfm_b200.models.superglue import SuperGlue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
m = SuperGlue(synthetic.superglue_state_dict(0), sinkhorn_iterations=20, match_threshold=0.4, precision="bf16", device=dev)
fr = synthetic.feature_scene(2 * B, K, seed=0)
data = {"descriptors0": torch.stack([fr[i][0] for i in range(B)]).to(dev),
        "descriptors1": torch.stack([fr[B + i][0] for i in range(B)]).to(dev),
        "keypoints0": torch.stack([fr[i][1] for i in range(B)]).to(dev),
        "keypoints1": torch.stack([fr[B + i][1] for i in range(B)]).to(dev),
        "scores0": torch.stack([fr[i][2] for i in range(B)]).to(dev),
        "scores1": torch.stack([fr[B + i][2] for i in range(B)]).to(dev),
        "image0": torch.zeros(B, 1, 480, 640), "image1": torch.zeros(B, 1, 480, 640)}
for _ in range(reps):
    out = m(data)
torch.cuda.synchronize()
print("matches", int((out["matches0"] > -1).sum()))
