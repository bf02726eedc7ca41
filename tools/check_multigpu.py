"""Run under torchrun on N GPUs: the opt-in multi-GPU mode of the Matcher must hand rank 0 exactly the
table a single-GPU run produces -- chunk-aligned pair shards keep BatchNorm statistics and truncation identical --
both from given frames (`match_frames`) and from image files (`match_simple_sfm_dataset`: SuperPoint by image,
one feature all-gather)."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from simplesfm_b200 import synthetic
from simplesfm_b200.matchers import Matcher
from simplesfm_b200.matchers.matcher import Frame

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0)
fr = synthetic.feature_scene(9, 512, seed=3)
counts = [512, 512, 400, 512, 512, 512, 333, 512, 512]
img = torch.zeros(1, 480, 640, device=dev)
frames = [Frame(d[:, :c].to(dev), k[:c].to(dev), s[:c].to(dev), img) for (d, k, s), c in zip(fr, counts)]
names = {i + 1: f"{i}.png" for i in range(9)}
kw = dict(matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.2, sinkhorn_iterations=20,
          super_glue_batch=4, precision="bf16", device=dev)
same = lambda a, b: list(a.keys()) == list(b.keys()) and all(np.array_equal(a[k], b[k]) for k in a)
single = Matcher(spw, 4, 0.005, shard_pairs=False, **kw).match_frames(frames, names, None)[1]
sharded = Matcher(spw, 4, 0.005, shard_pairs=True, **kw).match_frames(frames, names, None)[1]
ok = same(single, sharded) if rank == 0 else (len(sharded) < len(single) and all(np.array_equal(single[k], sharded[k]) for k in sharded))

# image files -> image-sharded SuperPoint -> all-gather -> pair-sharded SuperGlue -> tables on rank 0
from PIL import Image
tmp = tempfile.mkdtemp() if rank == 0 else None
box = [tmp]
dist.broadcast_object_list(box, src=0)
tmp = box[0]
n = 7
if rank == 0:
    meta = []
    for i in range(n):
        im = (synthetic.textured_image(60, 240, 320, shift=(2 * i, -i))[0, 0] * 255).round().to(torch.uint8).numpy()
        Image.fromarray(im).save(os.path.join(tmp, f"v{i}.png"))
        meta.append({"file_path": f"./v{i}.png"})
    json.dump({"frames": meta}, open(os.path.join(tmp, "views_data.json"), "w"))
dist.barrier()
kw2 = dict(kw, max_keypoints=300)
f1, t1, n1 = Matcher(spw, 4, 0.005, shard_pairs=False, **kw2).match_simple_sfm_dataset(tmp)
f2, t2, n2 = Matcher(spw, 4, 0.005, shard_pairs=True, **kw2).match_simple_sfm_dataset(tmp)
ok_img = n1 == n2 and len(f1) == len(f2) and all(
    torch.equal(a.keypoints_poses, b.keypoints_poses) and torch.equal(a.descriptors, b.descriptors) and
    torch.equal(a.scores, b.scores) and a.image.shape == b.image.shape for a, b in zip(f1, f2))
ok_img = ok_img and (same(t1, t2) if rank == 0 else all(np.array_equal(t1[k], t2[k]) for k in t2))
t = torch.tensor([1 if ok else 0, 1 if ok_img else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multi-GPU ({world} ranks): match_frames {len(single)} pairs / {sum(len(v) for v in single.values())} matches "
          f"identical = {bool(t[0].item())}; image files -> tables {len(t1)} pairs / {sum(len(v) for v in t1.values())} "
          f"matches, frames and tables identical = {bool(t[1].item())}")
dist.destroy_process_group()
sys.exit(0 if t.min().item() else 1)
