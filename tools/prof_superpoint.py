"""SuperPoint bf16 forward of a batch of images for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import synthetic
from simplesfm_b200.models.superpoint import SuperPoint
dev = torch.device("cuda:0")
m = SuperPoint(synthetic.superpoint_state_dict(0), max_keypoints=2048, device=dev, precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
batch = torch.cat([synthetic.textured_image(i)[:, 0] for i in range(4)], 0).to(dev).contiguous()
for _ in range(2):
    cand, ws = m.detect(batch)
    m.describe(4, 480, 640, 2048, ws)
torch.cuda.synchronize()
print("ok", cand.tolist())
