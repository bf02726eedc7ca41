"""Run under torchrun on N GPUs: Matcher.match_frames with pair sharding must return, on every rank, exactly the
table a single-GPU run produces (chunk-aligned shards keep BatchNorm statistics and truncation identical)."""
import os, sys
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
single = Matcher(spw, 4, 0.005, shard_pairs=False, **kw).match_frames(frames, names, lambda p: True)[1]
sharded = Matcher(spw, 4, 0.005, shard_pairs=True, **kw).match_frames(frames, names, lambda p: True)[1]
ok = list(single.keys()) == list(sharded.keys()) and all(np.array_equal(single[k], sharded[k]) for k in single)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multi-GPU pair sharding ({world} ranks): {len(single)} pairs, "
          f"{sum(len(v) for v in single.values())} matches, identical on all ranks = {bool(t.item())}")
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
