"""SuperPoint timing: images/s for 480x640 inputs (fp32 path), CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import synthetic
from simplesfm_b200.models.superpoint import SuperPoint
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
m = SuperPoint(synthetic.superpoint_state_dict(0), max_keypoints=2048, device=dev, precision=prec)
imgs = [synthetic.textured_image(i).to(dev) for i in range(4)]
for im in imgs:
    out = m({"image": im})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for i in range(n):
    out = m({"image": imgs[i % 4]})
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"SuperPoint.forward 480x640 {prec}: {ms:.3f} ms/image ({1000/ms:.1f} images/s), {out['keypoints'].shape[0]} keypoints; "
      f"52.1 GFLOP -> {52.1/ms:.1f} TFLOP/s")
# batched detect (no per-image host sync)
batch = torch.cat([im[:, 0] for im in imgs * 2], 0).contiguous()
for _ in range(2):
    cand, ws = m.detect(batch)
    m.describe(batch.shape[0], 480, 640, 2048, ws)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    cand, ws = m.detect(batch)
    m.describe(batch.shape[0], 480, 640, 2048, ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5 / batch.shape[0]
print(f"batched detect+describe (8 images per call): {ms:.3f} ms/image ({1000/ms:.1f} images/s)")
