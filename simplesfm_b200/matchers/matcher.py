"""Drop-in mirror of `simple_sfm/matchers/matcher.py` on libsfm_b200 (sm_100a CUDA).

Same class, constructor arguments, methods and return values as the reference `Matcher`
(matcher.py:31-400) and the same `Frame` tuple (matcher.py:24-28).  Differences are additive
only: `max_keypoints`, `bn_mode`, `precision`, `device` keyword arguments, and -- opt-in with
`shard_pairs=True` on an initialised `torch.distributed` process group -- the multi-GPU partition of
SURVEY.md 8e: images dealt to the ranks for SuperPoint, one all-gather of the feature records, the pair
list cut into chunk-aligned ranges, compact match tables gathered to rank 0 (`distributed.py`).

The SuperGlue hot loop keeps the reference semantics -- consecutive chunks of `super_glue_batch`
pairs, every side truncated to the chunk-minimum keypoint count (matcher.py:73-91), `(L,2)`
uint32 tables keyed by 1-based `(id1, id2)` (matcher.py:103-109) -- but runs without per-pair
host synchronisation: features live in one device record store, every chunk is one C call, and
match tables come back with a single device->host copy per super-chunk.
"""
__all__ = ['Matcher']

import json
import logging
import time
from pathlib import Path
from typing import Dict, List, NamedTuple, Tuple, Union

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from ..models.superglue import SuperGlue
from ..models.superpoint import SuperPoint
from .simple_matcher import SimpleMatcher

logger = logging.getLogger(__name__)


class Frame(NamedTuple):
    descriptors: torch.Tensor
    keypoints_poses: torch.Tensor
    scores: torch.Tensor
    image: torch.Tensor


class FeatureStore:
    """Fixed-capacity per-image feature records on one GPU (SURVEY.md 8e phase 1/2):
    desc (n,256,kcap), kpts (n,kcap,2), scores (n,kcap), counts (host list)."""

    def __init__(self, frames: List[Frame], device):
        n = len(frames)
        self.counts = [int(f.keypoints_poses.shape[0]) for f in frames]
        self.kcap = max(self.counts + [1])
        self.kcap = (self.kcap + 3) // 4 * 4
        self.desc = torch.zeros(n, 256, self.kcap, dtype=torch.float32, device=device)
        self.kpts = torch.zeros(n, self.kcap, 2, dtype=torch.float32, device=device)
        self.scores = torch.zeros(n, self.kcap, dtype=torch.float32, device=device)
        self.hw = [tuple(f.image.shape[-2:]) for f in frames]
        for i, f in enumerate(frames):
            k = self.counts[i]
            if k:
                self.desc[i, :, :k].copy_(f.descriptors, non_blocking=True)
                self.kpts[i, :k].copy_(f.keypoints_poses, non_blocking=True)
                self.scores[i, :k].copy_(f.scores, non_blocking=True)

    @classmethod
    def from_records(cls, rec) -> "FeatureStore":
        """Adopt gathered records (`distributed.Records`) without copying them again."""
        self = cls.__new__(cls)
        self.desc, self.kpts, self.scores = rec.desc, rec.kpts, rec.scores
        self.counts, self.hw, self.kcap = list(rec.counts), list(rec.hw), int(rec.desc.shape[-1])
        return self


class FrameList(list):
    """`List[Frame]` whose frames are views into one `FeatureStore` (what the image sources return when the
    features were all-gathered): `_super_glue_matcher` then matches straight from the store."""
    store: FeatureStore = None


class Matcher(object):

    def __init__(self,
                 super_point_extractor_weights_path: str,
                 nms_radius: float,
                 keypoint_threshold: float,
                 matcher_type: str = 'simple',
                 super_glue_weights_path: str = None,
                 match_threshold: float = 0.4,
                 sinkhorn_iterations: int = 20,
                 super_glue_batch: int = 5,
                 max_keypoints: int = -1,
                 bn_mode: str = 'train',
                 precision: str = 'fp32',
                 device: str = 'cuda',
                 shard_pairs: bool = False,
                 fuse_chunks: int = 51,
                 superpoint_precision: str = 'fp32',
                 extract_batch: int = 8,
                 decode_threads: int = 4,
                 process_group=None,
                 gather_tables: str = 'root',
                 hoist_layer0: bool = True):

        self.matcher_type = matcher_type
        self.device = torch.device(device)
        # multi-GPU (opt-in): every rank must call the image sources / match_frames with the SAME scene.
        # `gather_tables`: 'root' = rank 0 returns the table of every pair, the other ranks only their own pairs
        # (rank 0 is the database writer); 'none' = every rank keeps only its own pairs.
        self.shard_pairs = bool(shard_pairs)
        self.process_group = process_group
        if gather_tables not in ('root', 'none'):
            raise ValueError("gather_tables must be 'root' or 'none'")
        self.gather_tables = gather_tables
        # eval-mode BatchNorm + bf16 path only: GNN layer 0 (a self layer) once per image instead of once per pair
        # (SURVEY.md H9); bit-identical results, used when no chunk truncates its keypoints
        self.hoist_layer0 = bool(hoist_layer0)
        # consecutive chunks with identical keypoint counts are launched together (better GPU occupancy);
        # BatchNorm statistics stay per chunk, so results equal the chunk-by-chunk loop of the reference
        self.fuse_chunks = max(1, int(fuse_chunks)) if precision == 'bf16' else 1

        self.super_point_extractor = SuperPoint(weights_path=super_point_extractor_weights_path,
                                                nms_radius=nms_radius,
                                                keypoint_threshold=keypoint_threshold,
                                                max_keypoints=max_keypoints,
                                                device=device, precision=superpoint_precision)
        self.super_point_extractor.cuda()
        # image ingestion: images are decoded ahead on `decode_threads` threads, uploaded as 8-bit pixels and run
        # through SuperPoint `extract_batch` at a time (1 = the reference's one-image loop); see ingest.py
        self.extract_batch = max(1, int(extract_batch))
        self.decode_threads = max(1, int(decode_threads))

        if self.matcher_type == 'simple':
            self.matcher = SimpleMatcher(match_threshold)
        elif self.matcher_type == 'super_glue':
            assert super_glue_weights_path is not None, 'Please set weights path for SuperGlue.'
            self.matcher = SuperGlue(weights_path=super_glue_weights_path,
                                     match_threshold=match_threshold,
                                     sinkhorn_iterations=sinkhorn_iterations,
                                     bn_mode=bn_mode, precision=precision, device=device)
            self.super_glue_batch = super_glue_batch
            self.matcher.cuda()

    # ------------------------------------------------------------------------------------------
    def _super_glue_matcher(self,
                            frames: List[Frame],
                            match_list: np.ndarray,
                            super_chunk: int = 2048) -> Dict[Tuple[int, int], np.array]:
        """
        :param frames: Frames with filed keypoints_poses, images, scores and descriptors fields.
        :param match_list: (P,2) array of 1-based image ids which must be matched.
        :return: Dict where the key is a pair from the match_list,
                and the value is the points that match between images in pair.
        """
        return self._super_glue_tables(frames, match_list, super_chunk).to_dict()

    def _super_glue_tables(self, frames: List[Frame], match_list: np.ndarray, super_chunk: int = 2048,
                           exhaustive_first: int = None):
        """The hot loop (matcher.py:60-113) with compact output: `distributed.CompactTables` (pairs, per-pair
        lengths, all rows concatenated) -- no per-pair host object until `to_dict()`.  `exhaustive_first`: the
        pairs are entries [first, first + P) of the unfiltered lexicographic schedule of ids 1..n, so the index
        arrays are generated on the device (`sfm_pair_list`) instead of being uploaded."""
        from ..distributed import CompactTables, concat_tables
        match_list = np.asarray(match_list, dtype=np.int64).reshape(-1, 2)
        if len(match_list) == 0:
            return concat_tables([])
        sg: SuperGlue = self.matcher
        dev = sg.device
        store = getattr(frames, 'store', None) or FeatureStore(frames, dev)
        pairs0 = np.ascontiguousarray(match_list - 1)
        P = len(pairs0)
        L = lib()
        if exhaustive_first is not None:
            idx0_all = torch.empty(P, dtype=torch.int32, device=dev)
            idx1_all = torch.empty(P, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                check(L.sfm_pair_list(len(frames), int(exhaustive_first), P, ptr(idx0_all), ptr(idx1_all),
                                      stream_ptr(dev)))
        else:
            idx0_all = torch.from_numpy(np.ascontiguousarray(pairs0[:, 0]).astype(np.int32)).to(dev)
            idx1_all = torch.from_numpy(np.ascontiguousarray(pairs0[:, 1]).astype(np.int32)).to(dev)
        bs = int(self.super_glue_batch)
        counts = np.asarray(store.counts, dtype=np.int64)
        uniform_hw = len(set(store.hw)) == 1
        super_chunk = max(bs, super_chunk // bs * bs)
        # chunks per launch: bounded so that one launch holds at most 256 pairs (score matrices + activations of a
        # launch at 4096 keypoints stay around 35 GB)
        fuse_limit = max(1, min(self.fuse_chunks, 256 // bs))
        # layer-0 hoisting: valid when BatchNorm uses running statistics and no chunk truncates (equal counts)
        xstore = None
        if (self.hoist_layer0 and sg.can_hoist() and uniform_hw and len(counts) and counts.min() == counts.max()
                and counts[0] > 0):
            used = np.unique(pairs0)
            K = int(counts[0])
            xstore = torch.empty(len(counts), K, 256, dtype=torch.float32, device=dev)
            for i0 in range(0, len(used), 64):      # 64 images per pass bounds the workspace
                blk = used[i0:i0 + 64]
                if len(blk) == blk[-1] - blk[0] + 1:     # contiguous run of image indices: no gather needed
                    lo, hi = int(blk[0]), int(blk[-1]) + 1
                    xstore[lo:hi] = sg.hoist_layer0(store.desc[lo:hi], store.kpts[lo:hi], store.scores[lo:hi],
                                                    hi - lo, K, store.hw[0])
                else:
                    sel = torch.from_numpy(blk.astype(np.int64)).to(dev)
                    xstore[sel] = sg.hoist_layer0(store.desc[sel].contiguous(), store.kpts[sel].contiguous(),
                                                  store.scores[sel].contiguous(), len(blk), K, store.hw[0])
        parts = []
        for s0 in range(0, P, super_chunk):
            s1 = min(P, s0 + super_chunk)
            n_sc = s1 - s0
            tables = torch.empty(n_sc, store.kcap, 2, dtype=torch.int32, device=dev)
            lengths = torch.zeros(n_sc, dtype=torch.int32, device=dev)
            # chunk-minimum keypoint counts of both sides (matcher.py:82-83), all chunks of the super-chunk at once
            n_ch = (n_sc + bs - 1) // bs
            pad = n_ch * bs - n_sc
            c0s = counts[pairs0[s0:s1, 0]]
            c1s = counts[pairs0[s0:s1, 1]]
            if pad:
                c0s = np.concatenate([c0s, np.full(pad, c0s[-1])])
                c1s = np.concatenate([c1s, np.full(pad, c1s[-1])])
            k0s = c0s.reshape(n_ch, bs).min(1)
            k1s = c1s.reshape(n_ch, bs).min(1)
            # plan: (first pair, number of pairs, k0, k1, hw0, hw1, fused chunk count)
            plan = []
            for ci, c0 in enumerate(range(s0, s1, bs)):
                c1 = min(s1, c0 + bs)
                k0, k1 = int(k0s[ci]), int(k1s[ci])
                if uniform_hw:
                    h0 = h1 = store.hw[0]
                else:
                    hw0 = {store.hw[i] for i in pairs0[c0:c1, 0]}
                    hw1 = {store.hw[i] for i in pairs0[c0:c1, 1]}
                    if len(hw0) != 1 or len(hw1) != 1:
                        raise RuntimeError("stack expects each tensor to be equal size (images of one chunk differ)")
                    h0, h1 = next(iter(hw0)), next(iter(hw1))
                key = (k0, k1, h0, h1)
                full = (c1 - c0) == bs
                if (plan and full and tuple(plan[-1][2:6]) == key and plan[-1][1] == plan[-1][6] * bs):
                    plan[-1][1] += bs
                    plan[-1][6] += 1
                else:
                    plan.append([c0, c1 - c0, *key, 1])
            # a run of fusable chunks longer than the limit is cut into EQUAL launches (245 chunks at a limit of 51 ->
            # 5 x 49, not 4 x 51 + 41): every launch then fills the SMs equally well, no small tail launch
            runs, plan = plan, []
            for c0, B, k0, k1, h0, h1, groups in runs:
                for g in self.balanced_launches(groups, fuse_limit):
                    nb = min(g * bs, B)
                    plan.append([c0, nb, k0, k1, h0, h1, g])
                    c0 += nb
                    B -= nb
            for c0, B, k0, k1, hw0, hw1, groups in plan:
                c1 = c0 + B
                if k0 == 0 or k1 == 0:      # reference early-out: no matches (superglue.py:240-247)
                    continue
                if xstore is not None:
                    m0, _, _, _, _ = sg.match_hoisted(xstore, idx0_all[c0:c1], xstore, idx1_all[c0:c1], B, k0, k1)
                else:
                    m0, _, _, _, _ = sg.match_records(store.desc, store.kpts, store.scores, idx0_all[c0:c1],
                                                      store.desc, store.kpts, store.scores, idx1_all[c0:c1],
                                                      B, k0, k1, hw0, hw1, chunk_groups=groups)
                # compact into the super-chunk table; rows of pair p start at tables[p - s0]
                tview = tables[c0 - s0:c1 - s0]
                if k0 == store.kcap:
                    dst = tview
                else:
                    dst = torch.empty(B, k0, 2, dtype=torch.int32, device=dev)
                with torch.cuda.device(dev):
                    check(L.sfm_compact_matches(ptr(m0), B, k0, ptr(dst), ptr(lengths[c0 - s0:c1 - s0]),
                                                stream_ptr(dev)))
                if dst is not tview:
                    tview[:, :k0].copy_(dst)
            # flatten on the device, then ONE device->host copy of exactly the rows that exist
            ends = torch.cumsum(lengths, 0, dtype=torch.int64)
            offsets = ends - lengths
            lens_host = lengths.cpu().numpy().astype(np.int64)              # the one sync per super-chunk
            total = int(lens_host.sum())
            flat = torch.empty(max(total, 1), 2, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                check(L.sfm_pack_tables(ptr(tables), n_sc, store.kcap, ptr(lengths), ptr(offsets), ptr(flat),
                                        stream_ptr(dev)))
            rows = flat[:total].cpu().numpy().view(np.uint32)
            parts.append(CompactTables(match_list[s0:s1], lens_host, rows))
        return concat_tables(parts)

    def _simple_matcher(self, frames: List[Frame], match_list: np.ndarray) -> Dict[Tuple[int, int], np.array]:
        """Mutual nearest-neighbour tables (matcher.py:115-131): per pair the first two rows of
        `match_two_way_cuda` ([index0, index1, distance]) as an (L, 2) uint32 array."""
        return self._simple_tables(frames, match_list).to_dict()

    def _simple_tables(self, frames: List[Frame], match_list: np.ndarray):
        from ..distributed import CompactTables, concat_tables
        match_list = np.asarray(match_list, dtype=np.int64).reshape(-1, 2)
        if len(match_list) == 0:
            return concat_tables([])
        nn = self.matcher
        rows, lens = nn.match_pairs([f.descriptors for f in frames], match_list - 1)
        return CompactTables(match_list, lens, rows)

    def _group(self):
        """(process group, rank, world) of the opt-in multi-GPU mode, (None, 0, 1) otherwise."""
        from ..distributed import active_group
        return active_group(self.shard_pairs, self.process_group)

    def _shard_bounds(self, n_pairs: int) -> Tuple[int, int]:
        """[lo, hi) of this rank in a pair list of `n_pairs` entries: contiguous and chunk-aligned (SURVEY.md 8e
        phase 3), which keeps train-mode BatchNorm statistics identical to a single-GPU run."""
        group, rank, world = self._group()
        if group is None:
            return 0, n_pairs
        from ..distributed import shard_range
        return shard_range(n_pairs, getattr(self, 'super_glue_batch', 1), rank, world)

    def _shard(self, match_list: np.ndarray) -> np.ndarray:
        lo, hi = self._shard_bounds(len(match_list))
        return match_list[lo:hi]

    @staticmethod
    def balanced_launches(n_chunks: int, limit: int) -> List[int]:
        """Chunks per launch for a run of `n_chunks` fusable chunks: the fewest launches the limit allows, of equal
        size to within one chunk (245 at a limit of 51 -> [49] * 5)."""
        if n_chunks <= 0:
            return []
        n_launch = -(-n_chunks // max(1, limit))
        base, extra = divmod(n_chunks, n_launch)
        return [base + (1 if li < extra else 0) for li in range(n_launch)]

    @staticmethod
    def pair_list(ids, match_pairs_filter: callable = None) -> np.ndarray:
        """(P, 2) array of the id pairs to match: `itertools.combinations(ids, 2)` order, filtered
        (matcher.py:143-144).  Built without a Python loop over the pairs where possible (scenes with 10^5..10^6
        pairs): the upper triangle comes from `np.triu_indices`, and a filter written with array
        arithmetic -- e.g. the reference's `lambda x: np.abs(x[0] - x[1]) < radius`
        (scripts/generate_simple_sfm_dataset_from_scene.py:140-143) -- is evaluated once on all pairs; its answer is
        used only if it is a boolean vector that agrees with pair-by-pair evaluation (all pairs up to 20 000,
        a 2 048-pair sample beyond), otherwise the filter is applied pair by pair like the reference does."""
        ids = np.asarray(list(ids))
        a, b = np.triu_indices(len(ids), k=1)
        every = np.stack([ids[a], ids[b]], axis=1).reshape(-1, 2)
        if match_pairs_filter is None or len(every) == 0:
            return every
        keep = None
        try:
            ans = match_pairs_filter(every.T)
            if isinstance(ans, np.ndarray) and ans.dtype == np.bool_ and ans.shape == (len(every),):
                # The array answer is trusted only after comparison with pair-by-pair calls: EVERY pair up to 20 000
                # pairs; beyond that 2 048 pairs (evenly spaced + seeded random), unless the filter declares itself
                # element-wise with an attribute `vectorized = True`.
                if getattr(match_pairs_filter, 'vectorized', False):
                    probe = np.zeros(0, np.int64)
                elif len(every) <= 20000:
                    probe = np.arange(len(every))
                else:
                    probe = np.unique(np.concatenate([
                        np.linspace(0, len(every) - 1, num=1024).astype(np.int64),
                        np.random.RandomState(0).randint(0, len(every), size=1024)]))
                if all(bool(match_pairs_filter(every[i])) == bool(ans[i]) for i in probe):
                    keep = ans
        except Exception:       # not an array-style filter
            keep = None
        if keep is None:
            keep = np.fromiter((bool(match_pairs_filter(p)) for p in every), dtype=np.bool_, count=len(every))
        return every[keep].reshape(-1, 2)

    @staticmethod
    def _report(what: str, seconds: float, **counts):
        logger.info(f'Finished {what}\n Elapsed time: {seconds}\n'
                    + ''.join(f' {k.replace("_", " ").capitalize()}: {v}\n' for k, v in counts.items()))

    def match_frames(self,
                     frames: List[Frame],
                     images_names: dict,
                     match_pairs_filter: callable = None) -> Tuple[List[Frame],
                                                                   Dict[Tuple[int, int], np.array],
                                                                   Dict[int, str]]:
        """Match every id pair of `images_names` that passes `match_pairs_filter`, in the lexicographic order
        of `itertools.combinations` (matcher.py:133-165).  Returns `(frames, match_table, images_names)`."""
        t0 = time.time()
        logger.info('Starts matching descriptors.')
        wanted = self.pair_list(list(images_names.keys()), match_pairs_filter)
        group, rank, world = self._group()
        if group is not None:     # every rank must hold the same scene
            from ..distributed import assert_same_everywhere
            assert_same_everywhere([len(frames), len(wanted), sum(int(f.keypoints_poses.shape[0]) for f in frames)],
                                   self.device, group, what="frames / pair list")
        lo, hi = self._shard_bounds(len(wanted))
        mine = wanted[lo:hi]
        # unfiltered schedule over ids 1..n: the pair index arrays are generated on the device (8f-3)
        ids = list(images_names.keys())
        exhaustive = match_pairs_filter is None and len(ids) == len(frames) and ids == list(range(1, len(ids) + 1))
        if self.matcher_type == 'super_glue':
            tables = self._super_glue_tables(frames, mine, exhaustive_first=lo if exhaustive else None)
        elif self.matcher_type == 'simple':
            tables = self._simple_tables(frames, mine)
        else:
            tables = None
        if tables is None:
            match_table = None
        else:
            if group is not None and self.gather_tables == 'root':
                # compact uint32 rows + lengths to rank 0 (the database writer); other ranks keep their own pairs
                from ..distributed import gather_tables_to_root
                merged = gather_tables_to_root(tables, wanted, getattr(self, 'super_glue_batch', 1), self.device, group)
                tables = merged if merged is not None else tables
            match_table = tables.to_dict()
        self._report('matching', time.time() - t0, num_matches=len(wanted),
                     num_matched_points=0 if match_table is None else sum(len(v) for v in match_table.values()))
        return frames, match_table, images_names

    # ------------------------------------------------------------------------------------------
    # image sources (matcher.py:167-400).  All three feed one ingestion pipeline (ingest.py): host decode runs ahead
    # on `decode_threads` threads, pixels go up as they are, SuperPoint runs `extract_batch` images at a time;
    # image ids are 1-based in arrival order, exactly like the reference's per-image loops.
    def _frames_from(self, items, presharded: bool = False) -> List[Frame]:
        """items: iterable of (grey image (H, W) u8 | f32 host array, mask host array | None) over ALL images of the
        scene, in image-id order.  In the multi-GPU mode rank r extracts images r, r + G, r + 2G, ... only
        (`presharded=True`: `items` already holds just this rank's images) and one all-gather hands every rank the
        features of every image (SURVEY.md 8e phases 1-2); the frames are then views into the gathered records."""
        from ..ingest import FrameExtractor
        t0 = time.time()
        logger.info('Starts extracting descriptors.')
        group, rank, world = self._group()
        if group is not None and not presharded:
            items = (it for i, it in enumerate(items) if i % world == rank)
        ex = FrameExtractor(self.super_point_extractor, self.extract_batch)
        extracted = ex.extract(items)
        if group is None:
            frames = [Frame(e.descriptors, e.keypoints, e.scores, e.image) for e in extracted]
        else:
            from ..distributed import allgather_records
            rec = allgather_records([(e.descriptors, e.keypoints, e.scores, tuple(e.image.shape[-2:]))
                                     for e in extracted], None, self.device, group)
            frames = FrameList()
            frames.store = FeatureStore.from_records(rec)
            blank = torch.zeros((), dtype=torch.float32, device=self.device)
            for i, (k, hw) in enumerate(zip(rec.counts, rec.hw)):
                # the image itself stays on the rank that decoded it; elsewhere only its SHAPE is needed
                # (superglue.py:250-251), carried by a zero-stride view
                image = extracted[i // world].image if i % world == rank else blank.expand(1, hw[0], hw[1])
                frames.append(Frame(rec.desc[i, :, :k], rec.kpts[i, :k], rec.scores[i, :k], image))
        self._report('extraction', time.time() - t0, num_images=len(frames),
                     num_keypoints=sum(len(f.keypoints_poses) for f in frames))
        return frames

    def match_datasets(self,
                       datasets,
                       match_pairs_filter: callable = None,
                       image_name_path_shift: int = 1) -> Tuple[List[Frame],
                                                                Dict[Tuple[int, int], np.array],
                                                                Dict[int, str],
                                                                List[int]]:
        """One `ImageFoldersDataset` or a list of them (folder types 'rgb' and, optionally, 'mask'); the last
        return value holds, per dataset, the id following its final image (matcher.py:167-247)."""
        from torch.utils.data import DataLoader
        names, split_index = {}, []

        def samples():
            for ds in (datasets if isinstance(datasets, list) else [datasets]):
                for sample in DataLoader(ds, batch_size=1, shuffle=False):
                    tail = sample['path_rgb'][0].split('/')[-image_name_path_shift:]
                    names[len(names) + 1] = '_'.join(tail)
                    grey = sample['rgb'][0, 0].numpy().astype(np.float32, copy=False)
                    yield grey, ((sample['mask'][0, 0].numpy() > 0) if 'mask' in sample else None)
                split_index.append(len(names) + 1)

        frames = self._frames_from(samples())
        return (*self.match_frames(frames, names, match_pairs_filter), split_index)

    def match_video_stream(self,
                           vs,
                           match_pairs_filter: callable = None,
                           vs_mask=None) -> Tuple[List[Frame],
                                                  Dict[Tuple[int, int], np.array],
                                                  Dict[int, str]]:
        """Every frame of a `VideoStreamer` (`next_frame() -> (f32 image in [0, 1], ok, path)`), optionally with a
        second streamer delivering the masks; stops at the first failed read of either (matcher.py:249-321)."""
        names = {}

        def grabbed():
            while True:
                image, ok, path = vs.next_frame()
                mask = None
                if vs_mask is not None:
                    mask, ok_mask, _ = vs_mask.next_frame()
                    ok = ok and ok_mask
                if not ok:
                    return
                names[len(names) + 1] = path.split('/')[-1]
                yield np.asarray(image, dtype=np.float32), (None if mask is None else np.asarray(mask) > 0)

        frames = self._frames_from(grabbed())
        return self.match_frames(frames, names, match_pairs_filter)

    def match_simple_sfm_dataset(self,
                                 dataset_path: str,
                                 match_pairs_filter: callable = None,
                                 dataset_mask_name: str = None) -> Tuple[List[Frame],
                                                                         Dict[Tuple[int, int], np.array],
                                                                         Dict[int, str]]:
        """Every entry of `<dataset_path>/views_data.json['frames']`: `file_path` is the image, the optional key
        `dataset_mask_name` its mask; both are read as 8-bit grey (PIL mode 'L') (matcher.py:323-400)."""
        from ..ingest import decode_ahead, load_grey_u8
        root = Path(dataset_path)
        with open(root / 'views_data.json', encoding='UTF-8') as fh:
            entries = json.load(fh)['frames']

        def reader(entry):
            def read():
                mask = load_grey_u8(root / entry[dataset_mask_name]) if dataset_mask_name is not None else None
                return load_grey_u8(root / entry['file_path']), mask     # nonzero grey level == `mask / 255 > 0`
            return read

        names = {i + 1: e['file_path'].split('/')[-1] for i, e in enumerate(entries)}
        _, rank, world = self._group()
        loaders = [reader(e) for e in entries][rank::world]      # this rank decodes only the images it extracts
        frames = self._frames_from(decode_ahead(loaders, self.decode_threads), presharded=True)
        return self.match_frames(frames, names, match_pairs_filter)
