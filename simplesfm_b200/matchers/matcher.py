"""Drop-in mirror of `simple_sfm/matchers/matcher.py` on libsfm_b200 (sm_100a CUDA).

Same class, constructor arguments, methods and return values as the reference `Matcher`
(matcher.py:31-400) and the same `Frame` tuple (matcher.py:24-28).  Differences are additive
only: `max_keypoints`, `bn_mode`, `precision`, `device` keyword arguments, and pair sharding
across ranks when `torch.distributed` is initialised (`shard_pairs=True`).

The SuperGlue hot loop keeps the reference semantics -- consecutive chunks of `super_glue_batch`
pairs, every side truncated to the chunk-minimum keypoint count (matcher.py:73-91), `(L,2)`
uint32 tables keyed by 1-based `(id1, id2)` (matcher.py:103-109) -- but runs without per-pair
host synchronisation: features live in one device record store, every chunk is one C call, and
match tables come back with a single device->host copy per super-chunk.
"""
__all__ = ['Matcher']

import itertools
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
                 shard_pairs: bool = True,
                 fuse_chunks: int = 32,
                 superpoint_precision: str = 'fp32',
                 extract_batch: int = 8,
                 decode_threads: int = 4):

        self.matcher_type = matcher_type
        self.device = torch.device(device)
        self.shard_pairs = shard_pairs
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
        match_table = {}
        if len(match_list) == 0:
            return match_table
        sg: SuperGlue = self.matcher
        dev = sg.device
        store = FeatureStore(frames, dev)
        pairs0 = np.ascontiguousarray(match_list - 1)
        idx0_all = torch.from_numpy(np.ascontiguousarray(pairs0[:, 0]).astype(np.int32)).to(dev)
        idx1_all = torch.from_numpy(np.ascontiguousarray(pairs0[:, 1]).astype(np.int32)).to(dev)
        bs = int(self.super_glue_batch)
        P = len(pairs0)
        L = lib()
        super_chunk = max(bs, super_chunk // bs * bs)
        for s0 in range(0, P, super_chunk):
            s1 = min(P, s0 + super_chunk)
            n_sc = s1 - s0
            tables = torch.empty(n_sc, store.kcap, 2, dtype=torch.int32, device=dev)
            lengths = torch.zeros(n_sc, dtype=torch.int32, device=dev)
            # plan: (first pair, number of pairs, k0, k1, hw0, hw1, fused chunk count)
            plan = []
            for c0 in range(s0, s1, bs):
                c1 = min(s1, c0 + bs)
                k0 = min(store.counts[i] for i in pairs0[c0:c1, 0])      # matcher.py:82
                k1 = min(store.counts[i] for i in pairs0[c0:c1, 1])      # matcher.py:83
                hw0 = {store.hw[i] for i in pairs0[c0:c1, 0]}
                hw1 = {store.hw[i] for i in pairs0[c0:c1, 1]}
                if len(hw0) != 1 or len(hw1) != 1:
                    raise RuntimeError("stack expects each tensor to be equal size (images of one chunk differ)")
                key = (k0, k1, next(iter(hw0)), next(iter(hw1)))
                full = (c1 - c0) == bs
                if (plan and full and tuple(plan[-1][2:6]) == key and plan[-1][6] < self.fuse_chunks
                        and plan[-1][1] == plan[-1][6] * bs):
                    plan[-1][1] += bs
                    plan[-1][6] += 1
                else:
                    plan.append([c0, c1 - c0, *key, 1])
            for c0, B, k0, k1, hw0, hw1, groups in plan:
                c1 = c0 + B
                if k0 == 0 or k1 == 0:      # reference early-out: no matches (superglue.py:240-247)
                    continue
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
            lens = lengths.cpu().numpy()                                    # one sync per super-chunk
            maxlen = int(lens.max()) if n_sc else 0
            host = tables[:, :max(maxlen, 1)].cpu().numpy().view(np.uint32)
            for p in range(n_sc):
                a, b = pairs0[s0 + p]
                match_table[(a + 1, b + 1)] = host[p, :lens[p]].copy()
        return match_table

    def _simple_matcher(self,
                        frames: List[Frame],
                        match_list: np.ndarray) -> Dict[Tuple[int, int], np.array]:
        """
        :param frames: Frames with filed descriptors field.
        :param match_list: List with pairs which must be matched.
        :return: Dict where the key is a pair from the match_list,
                and the value is the points that match between images in pair.
        """
        match_table = {}
        for image_id_1, image_id_2 in match_list:
            matches = self.matcher.match_two_way_cuda(frames[image_id_1 - 1].descriptors,
                                                      frames[image_id_2 - 1].descriptors).t()

            matches = matches[:, :2].cpu().numpy().astype(np.uint32)
            match_table[(image_id_1, image_id_2)] = matches
        return match_table

    def _shard(self, match_list: np.ndarray) -> np.ndarray:
        """Contiguous, chunk-aligned slice of the pair list for this rank (SURVEY.md 8e phase 3).
        Chunk alignment keeps train-mode BatchNorm statistics identical to a single-GPU run."""
        import torch.distributed as dist
        if not (self.shard_pairs and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return match_list
        from ..distributed import shard_range
        bs = getattr(self, 'super_glue_batch', 1)
        lo, hi = shard_range(len(match_list), bs, dist.get_rank(), dist.get_world_size())
        return match_list[lo:hi]

    def match_frames(self,
                     frames: List[Frame],
                     images_names: dict,
                     match_pairs_filter: callable = None) -> Tuple[List[Frame],
                                                                   Dict[Tuple[int, int], np.array],
                                                                   Dict[int, str]]:

        logger.info('Starts matching descriptors.')
        start = time.time()

        match_list = np.array(list(itertools.combinations(images_names.keys(), r=2)))
        match_list = np.array(list(filter(match_pairs_filter, match_list))).reshape(-1, 2)
        local_list = self._shard(match_list)

        match_table = None

        if self.matcher_type == 'simple':
            match_table = self._simple_matcher(frames, local_list)
        elif self.matcher_type == 'super_glue':
            match_table = self._super_glue_matcher(frames, local_list)

        if local_list is not match_list:
            from ..distributed import gather_match_tables
            match_table = gather_match_tables(match_table, match_list)

        num_matched_points = 0
        for keys, value in match_table.items():
            num_matched_points += len(value)

        end = time.time()
        logger.info(f'Finished \n'
                    f'Elapsed time: {float(end - start)} \n'
                    f'Num matches: {len(match_list)} \n'
                    f'Num matched points: {num_matched_points} \n')

        return frames, match_table, images_names

    # ------------------------------------------------------------------------------------------
    def _extract(self, image: torch.Tensor, mask: torch.Tensor = None) -> Dict[str, torch.Tensor]:
        data = {'image': image}
        if mask is not None:
            data['mask'] = mask
        return self.super_point_extractor(data)

    def _extract_stream(self, items) -> List[Frame]:
        """Frames for an iterable of (grey image (H,W) u8 | f32, mask | None) host arrays, in order."""
        from ..ingest import FrameExtractor
        ex = FrameExtractor(self.super_point_extractor, self.extract_batch)
        return [Frame(e.descriptors, e.keypoints, e.scores, e.image) for e in ex.extract(items)]

    def match_datasets(self,
                       datasets,
                       match_pairs_filter: callable = None,
                       image_name_path_shift: int = 1) -> Tuple[List[Frame],
                                                                Dict[Tuple[int, int], np.array],
                                                                Dict[int, str],
                                                                List[int]]:
        """
        Get ImageFoldersDataset dataset or list of ImageFoldersDatasets and processed all frames from it,
        image folders type must be set as 'rgb' and 'mask'(if masks are used).  (matcher.py:167-247)
        """
        from torch.utils.data import DataLoader

        if match_pairs_filter is None:
            match_pairs_filter = lambda x: True

        logger.info('Starts extracting descriptors.')
        start = time.time()
        image_bd_index = 1
        images_bd_index_names = {}
        processed_frames = []
        num_keypoints = 0

        if not isinstance(datasets, list):
            datasets = [datasets]

        dataset_split_bd_index = []
        if self.extract_batch > 1:
            def stream():
                nonlocal image_bd_index
                for dataset in datasets:
                    for data in DataLoader(dataset, batch_size=1, shuffle=False):
                        images_names = '_'.join(data['path_rgb'][0].split('/')[-image_name_path_shift:])
                        images_bd_index_names[image_bd_index] = images_names
                        image_bd_index += 1
                        img = data['rgb'][0, 0].numpy().astype(np.float32, copy=False)
                        msk = (data['mask'][0, 0].numpy() > 0) if 'mask' in data else None
                        yield img, msk
                    dataset_split_bd_index.append(image_bd_index)
            processed_frames = self._extract_stream(stream())
            num_keypoints = sum(len(f.keypoints_poses) for f in processed_frames)
            datasets = []
        for dataset in datasets:
            image_dataloader = DataLoader(dataset, batch_size=1, shuffle=False)
            for data in image_dataloader:
                images = data['rgb'].to(self.device)
                images_names = '_'.join(data['path_rgb'][0].split('/')[-image_name_path_shift:])
                masks = data['mask'].to(self.device) > 0 if 'mask' in data else None
                result = self._extract(images, masks)
                processed_frames.append(Frame(result['descriptors'], result['keypoints'], result['scores'], images[0]))
                images_bd_index_names[image_bd_index] = images_names
                image_bd_index += 1
                num_keypoints += len(result['keypoints'])
            dataset_split_bd_index.append(image_bd_index)

        end = time.time()
        logger.info(f'Finished \n '
                    f'Elapsed time: {float(end - start)} \n'
                    f'Num images: {len(processed_frames)} \n'
                    f'Num keypoints: {num_keypoints} \n')

        return *self.match_frames(processed_frames, images_bd_index_names, match_pairs_filter), dataset_split_bd_index

    def match_video_stream(self,
                           vs,
                           match_pairs_filter: callable = None,
                           vs_mask=None) -> Tuple[List[Frame],
                                                  Dict[Tuple[int, int], np.array],
                                                  Dict[int, str]]:
        """
        Get VideoStreamer and processed all frames from it (matcher.py:249-321).
        `vs.next_frame()` must return (image float32 (H,W) in [0,1], status, path).
        """
        if match_pairs_filter is None:
            match_pairs_filter = lambda x: True

        logger.info('Starts extracting descriptors.')
        start = time.time()
        image_bd_index = 1
        images_bd_index_names = {}
        processed_frames = []
        num_keypoints = 0
        if self.extract_batch > 1:
            def stream():
                nonlocal image_bd_index
                while True:
                    image, status, img_path = vs.next_frame()
                    image_mask = None
                    if vs_mask is not None:
                        image_mask, status_mask, img_mask_path = vs_mask.next_frame()
                        status &= status_mask
                    if status is False:
                        return
                    images_bd_index_names[image_bd_index] = img_path.split('/')[-1]
                    image_bd_index += 1
                    yield np.asarray(image, dtype=np.float32), (None if image_mask is None else image_mask > 0)
            processed_frames = self._extract_stream(stream())
            num_keypoints = sum(len(f.keypoints_poses) for f in processed_frames)
        while self.extract_batch <= 1:
            image, status, img_path = vs.next_frame()
            if vs_mask is not None:
                image_mask, status_mask, img_mask_path = vs_mask.next_frame()
                status &= status_mask
            if status is False:
                break
            image = torch.from_numpy(image)[None, None, :, :].to(self.device)
            mask = None
            if vs_mask:
                mask = torch.from_numpy(image_mask > 0)[None, None, :, :].to(self.device)
            result = self._extract(image, mask)
            processed_frames.append(Frame(result['descriptors'], result['keypoints'], result['scores'],
                                          image.squeeze(1)))
            images_bd_index_names[image_bd_index] = img_path.split('/')[-1]
            image_bd_index += 1
            num_keypoints += len(result['keypoints'])

        end = time.time()
        logger.info(f'Finished \n '
                    f'Elapsed time: {float(end - start)} \n'
                    f'Num images: {len(processed_frames)} \n'
                    f'Num keypoints: {num_keypoints} \n')

        return self.match_frames(processed_frames, images_bd_index_names, match_pairs_filter)

    def match_simple_sfm_dataset(self,
                                 dataset_path: str,
                                 match_pairs_filter: callable = None,
                                 dataset_mask_name: str = None) -> Tuple[List[Frame],
                                                                         Dict[Tuple[int, int], np.array],
                                                                         Dict[int, str]]:
        """
        Process every frame listed in `<dataset_path>/views_data.json` (matcher.py:323-400).
        """
        from PIL import Image

        if match_pairs_filter is None:
            match_pairs_filter = lambda x: True

        logger.info('Starts extracting descriptors.')
        start = time.time()
        image_bd_index = 1
        images_bd_index_names = {}
        processed_frames = []
        num_keypoints = 0

        views_data_json_path = Path(dataset_path, "views_data.json")
        with open(views_data_json_path, encoding="UTF-8") as file:
            views_data = json.load(file)

        if self.extract_batch > 1:
            from ..ingest import decode_ahead, load_grey_u8

            def loader(item):
                def load():
                    img = load_grey_u8(Path(dataset_path, item['file_path']))
                    msk = load_grey_u8(Path(dataset_path, item[dataset_mask_name])) if dataset_mask_name is not None else None
                    return img, msk          # mask: nonzero grey level == `image_mask / 255 > 0`
                return load

            frames_meta = views_data['frames']
            processed_frames = self._extract_stream(decode_ahead([loader(it) for it in frames_meta],
                                                                 self.decode_threads))
            for item in frames_meta:
                images_bd_index_names[image_bd_index] = item['file_path'].split('/')[-1]
                image_bd_index += 1
            num_keypoints = sum(len(f.keypoints_poses) for f in processed_frames)
            views_data = {'frames': []}
        for item in views_data['frames']:
            image = np.array(Image.open(Path(dataset_path, item['file_path'])).convert('L')).astype('float32') / 255
            image_name = item['file_path'].split('/')[-1]
            mask = None
            if dataset_mask_name is not None:
                image_mask = np.array(Image.open(Path(dataset_path, item[dataset_mask_name])).convert('L')
                                      ).astype('float32') / 255
                mask = torch.from_numpy(image_mask > 0)[None, None, :, :].to(self.device)
            image = torch.from_numpy(image)[None, None, :, :].to(self.device)
            result = self._extract(image, mask)
            processed_frames.append(Frame(result['descriptors'], result['keypoints'], result['scores'],
                                          image.squeeze(1)))
            images_bd_index_names[image_bd_index] = image_name
            image_bd_index += 1
            num_keypoints += len(result['keypoints'])

        end = time.time()
        logger.info(f'Finished \n '
                    f'Elapsed time: {float(end - start)} \n'
                    f'Num images: {len(processed_frames)} \n'
                    f'Num keypoints: {num_keypoints} \n')

        return self.match_frames(processed_frames, images_bd_index_names, match_pairs_filter)
