"""Does the match table of the bench scene depend on how the pair list is cut into launches?  (It must not: the
multi-GPU mode relies on identical tables at any rank count.)  One GPU; compares the full run with the pair list cut
at the shard boundaries of 2 / 4 / 8 ranks and with chunk fusion off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplesfm_b200 import synthetic
from simplesfm_b200.distributed import concat_tables, shard_range
from simplesfm_b200.matchers import Matcher
from simplesfm_b200.matchers.matcher import Frame
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 50
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
spw, sgw = synthetic.superpoint_state_dict(0), synthetic.superglue_state_dict(0, "indoor")
mk = lambda fuse: Matcher(spw, 4, 0.005, matcher_type="super_glue", super_glue_weights_path=sgw, match_threshold=0.4,
                          sinkhorn_iterations=20, super_glue_batch=5, precision="bf16", device=dev, fuse_chunks=fuse)
m = mk(32)
img = torch.zeros(1, 480, 640, device=dev)
frames = [Frame(d.to(dev), k.to(dev), s.to(dev), img) for d, k, s in synthetic.feature_scene(n_img, kp, seed=0)]
wanted = m.pair_list(list(range(1, n_img + 1)), None)
full = m._super_glue_tables(frames, wanted)
again = m._super_glue_tables(frames, wanted)
def diff(a, b, tag):
    same = np.array_equal(a.lens, b.lens) and np.array_equal(a.rows, b.rows)
    nd = int((a.lens != b.lens).sum())
    if not same and nd == 0:
        cuts = np.concatenate([[0], np.cumsum(a.lens)])
        nd = sum(1 for i in range(len(a.lens)) if not np.array_equal(a.rows[cuts[i]:cuts[i+1]], b.rows[cuts[i]:cuts[i+1]]))
    print(f"{tag}: identical = {same}; pairs that differ: {nd} of {len(a.lens)}; matches {int(a.lens.sum())} vs {int(b.lens.sum())}")
diff(full, again, "same call twice")
for world in (2, 4, 8):
    parts = [m._super_glue_tables(frames, wanted[slice(*shard_range(len(wanted), 5, r, world))]) for r in range(world)]
    diff(full, concat_tables(parts), f"cut at the shard boundaries of {world} ranks")
diff(full, mk(1)._super_glue_tables(frames, wanted), "chunk fusion off")
diff(full, mk(8)._super_glue_tables(frames, wanted), "8 chunks per launch")
