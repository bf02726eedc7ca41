"""Image ingestion for the matching path.

The reference decodes one image at a time on the host, converts it to float32 there and runs SuperPoint
with batch 1 and a blocking read-back per image (`matcher.py:208-236, 280-312, 359-391`).  Once SuperPoint
costs a fraction of a millisecond that loop is the bottleneck, so this module

* decodes ahead of the GPU on a small thread pool (PIL releases the GIL while decoding),
* stages the raw 8-bit pixels (and masks) in pinned host buffers, two batches deep, and converts
  `u8 -> float32 / 255` on the device (`sfm_u8_to_unit_f32`, bit-identical to the host conversion),
* runs SuperPoint detect + describe on whole batches (`sfm_superpoint_detect / _describe`), with no host
  synchronisation at all when `max_keypoints` bounds the output size.

Every image gets exactly the keypoints / scores / descriptors the one-image `SuperPoint.forward` returns.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterable, Iterator, List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

Array = np.ndarray
# an item is (image, mask): image (H, W) uint8 (0..255) or float32 in [0, 1]; mask (H, W) any dtype, nonzero = keep
Item = Tuple[Array, Optional[Array]]


class Extracted(NamedTuple):
    """Per-image SuperPoint output (device tensors; views into the batch they were computed in)."""
    descriptors: torch.Tensor   # (256, K) f32
    keypoints: torch.Tensor     # (K, 2) f32 (x, y)
    scores: torch.Tensor        # (K,) f32
    image: torch.Tensor         # (1, H, W) f32 in [0, 1]


def load_grey_u8(path) -> Array:
    """`Image.open(path).convert('L')` as a (H, W) uint8 array (matcher.py:360 without the host-side /255)."""
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert('L'), dtype=np.uint8)


def decode_ahead(loaders: Sequence[Callable[[], Item]], threads: int = 4, depth: int = 16) -> Iterator[Item]:
    """Run the loader callables on a thread pool, at most `depth` ahead, yielding results in order."""
    if threads <= 1:
        for f in loaders:
            yield f()
        return
    with ThreadPoolExecutor(max_workers=threads) as pool:
        pending = []
        it = iter(loaders)
        for f in it:
            pending.append(pool.submit(f))
            if len(pending) >= depth:
                break
        while pending:
            fut = pending.pop(0)
            nxt = next(it, None)
            if nxt is not None:
                pending.append(pool.submit(nxt))
            yield fut.result()


class _Staging:
    """Two pinned host buffers per (shape, dtype) so that filling batch k+1 overlaps the upload of batch k."""

    def __init__(self):
        self.bufs = {}
        self.turn = {}

    def get(self, key, shape, dtype):
        if key not in self.bufs:
            self.bufs[key] = [torch.empty(shape, dtype=dtype, pin_memory=True) for _ in range(2)]
            self.turn[key] = 0
        events = self.bufs[key]
        i = self.turn[key]
        self.turn[key] = i ^ 1
        return events[i], (key, i)


class FrameExtractor:
    """Batched SuperPoint over a stream of grey images."""

    def __init__(self, super_point, batch_size: int = 8):
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        self.sp = super_point
        self.device = super_point.device
        self.batch_size = int(batch_size)
        self._staging = _Staging()
        self._busy = {}          # staging slot -> CUDA event recorded after its upload

    # ------------------------------------------------------------------------------------------
    def _upload(self, arrays: List[Array], tag: str) -> torch.Tensor:
        """Stack same-shape host arrays into a pinned buffer and start the H2D copy."""
        n = len(arrays)
        H, W = arrays[0].shape
        tdtype = torch.uint8 if arrays[0].dtype == np.uint8 else torch.float32
        host, slot = self._staging.get((tag, self.batch_size, H, W, tdtype), (self.batch_size, H, W), tdtype)
        ev = self._busy.get(slot)
        if ev is not None:
            ev.synchronize()     # the copy that last used this slot has finished
        view = host.numpy()
        for i, a in enumerate(arrays):
            np.copyto(view[i], a, casting="unsafe")
        dev = host[:n].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._busy[slot] = ev
        return dev

    def _to_unit_f32(self, dev: torch.Tensor) -> torch.Tensor:
        if dev.dtype == torch.float32:
            return dev
        out = torch.empty(dev.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().sfm_u8_to_unit_f32(ptr(dev), ptr(out), dev.numel(), stream_ptr(self.device)))
        return out

    def _run_batch(self, items: List[Item]):
        imgs = [np.ascontiguousarray(im) for im, _ in items]
        if imgs[0].dtype != np.uint8:
            imgs = [im.astype(np.float32, copy=False) for im in imgs]
        n = len(imgs)
        H, W = imgs[0].shape
        images = self._to_unit_f32(self._upload(imgs, "img"))
        mask = None
        if items[0][1] is not None:
            masks = [np.ascontiguousarray(m if m.dtype == np.uint8 else (m > 0).astype(np.uint8)) for _, m in items]
            mask = self._upload(masks, "mask")       # nonzero = keep (`mask > 0`, matcher.py:220,293,373)
        sp = self.sp
        cand, ws = sp.detect(images, mask)
        mk = sp.max_keypoints
        if mk > 0:
            cap = mk                                  # output size known: no host synchronisation
        else:
            cap = max(int(cand.max().item()), 1)      # unbounded: one read-back per batch
        kpts, scores, desc, counts = sp.describe(n, H, W, cap, ws)
        return images, kpts, scores, desc, counts

    # ------------------------------------------------------------------------------------------
    def extract(self, items: Iterable[Item]) -> List[Extracted]:
        """SuperPoint on every (image, mask) item, in order.  Consecutive items of equal shape / dtype /
        mask presence share a batch."""
        with torch.cuda.device(self.device):
            done = []           # (images, kpts, scores, desc, counts) per batch, counts still on the device
            batch: List[Item] = []
            key = None
            for im, m in items:
                if im.ndim != 2:
                    raise _lib.SfmError(-1, f"ingest: expected a (H, W) grey image, got shape {im.shape}")
                if m is not None and m.shape != im.shape:
                    raise _lib.SfmError(-1, f"ingest: mask shape {m.shape} != image shape {im.shape}")
                k = (im.shape, im.dtype == np.uint8, m is not None)
                if batch and (k != key or len(batch) == self.batch_size):
                    done.append(self._run_batch(batch))
                    batch = []
                key = k
                batch.append((im, m))
            if batch:
                done.append(self._run_batch(batch))
            out: List[Extracted] = []
            for images, kpts, scores, desc, counts in done:
                cnt = counts.tolist()           # one small read-back per batch, after everything was queued
                for i, k in enumerate(cnt):
                    out.append(Extracted(desc[i, :, :k], kpts[i, :k], scores[i, :k], images[i:i + 1]))
            return out
