"""Image ingestion throughput: PNG files on disk -> Frames.
(a) the reference's loop shape: decode, /255 on the host, one upload + one SuperPoint call + one read-back per image;
(b) simplesfm_b200.ingest: threaded decode, 8-bit upload, device /255, batched SuperPoint."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from simplesfm_b200 import synthetic
from simplesfm_b200.ingest import FrameExtractor, decode_ahead, load_grey_u8
from simplesfm_b200.models.superpoint import SuperPoint

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
tmp = tempfile.mkdtemp()
paths = []
for i in range(n):
    im = synthetic.textured_image(i % 8, shift=(i % 5, -(i % 7)))[0, 0].numpy()
    p = os.path.join(tmp, f"{i:04d}.png")
    Image.fromarray(np.clip(np.rint(im * 255), 0, 255).astype(np.uint8)).save(p)
    paths.append(p)
print(f"{n} PNG files, 640x480, host cores: {os.cpu_count()}")
for prec in ("fp32", "bf16"):
    sp = SuperPoint(synthetic.superpoint_state_dict(0), max_keypoints=2048, device=dev, precision=prec)
    for rep in range(2):      # second repetition is the measured one (page cache / workspace warm)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for p in paths:
            image = np.array(Image.open(p).convert('L')).astype('float32') / 255
            r = sp({"image": torch.from_numpy(image)[None, None].to(dev)})
        torch.cuda.synchronize(); ta = time.perf_counter() - t0
    for bs, th in ((8, 4), (16, 8)):
        ex = FrameExtractor(sp, bs)
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            fr = ex.extract(decode_ahead([(lambda p=p: (load_grey_u8(p), None)) for p in paths], th))
            torch.cuda.synchronize(); tb = time.perf_counter() - t0
        print(f"SuperPoint {prec}: per-image loop {n / ta:7.1f} images/s | batched x{bs}, {th} decode threads "
              f"{n / tb:7.1f} images/s ({ta / tb:.2f}x)")
