"""Drop-in mirror of `simple_sfm/models/superpoint.py` running on libsfm_b200 (sm_100a CUDA).

Same names, argument meaning and error behaviour as the reference module: `SuperPoint`
(superpoint.py:57-171) and the free functions `simple_nms` (:9-24), `remove_borders` (:27-32),
`top_k_keypoints` (:35-39), `sample_descriptors` (:42-54).  Additive kwargs only (`device`,
`precision`, `align_corners`).  All tensors must live on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Dict, Optional

import torch

from .. import _lib
from .._lib import Handle, check, lib, ptr, require_cuda, stream_ptr
from ..weights import load_state_dict_file, pack_superpoint

logger = logging.getLogger(__name__)


def default_align_corners() -> bool:
    """The reference decides on the 3rd character of torch.__version__ (superpoint.py:49);
    reproduce whatever it would do under the installed torch ('2.11.0' -> False)."""
    return int(torch.__version__[2]) > 2


def simple_nms(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    """ Fast Non-maximum suppression to remove nearby points.  scores: (b, H, W) CUDA f32. """
    assert (nms_radius >= 0)
    s = require_cuda(scores, "scores", torch.float32)
    b, H, W = s.shape
    out = torch.empty_like(s)
    with torch.cuda.device(s.device):
        check(lib().sfm_simple_nms(ptr(s), ptr(out), b, H, W, int(nms_radius), stream_ptr(s.device)))
    return out


def remove_borders(keypoints, scores, border: int, height: int, width: int):
    """ Removes keypoints too close to the border (index bookkeeping on (K,2) (row,col) tensors;
    the fused device path lives inside `SuperPoint.forward`). """
    mask_h = (keypoints[:, 0] >= border) & (keypoints[:, 0] < (height - border))
    mask_w = (keypoints[:, 1] >= border) & (keypoints[:, 1] < (width - border))
    mask = mask_h & mask_w
    return keypoints[mask], scores[mask]


def top_k_keypoints(keypoints, scores, k: int):
    if k >= len(keypoints):
        return keypoints, scores
    scores, indices = torch.topk(scores, k, dim=0)
    return keypoints[indices], scores


def sample_descriptors(keypoints: torch.Tensor, descriptors: torch.Tensor, s: int = 8,
                       align_corners: Optional[bool] = None) -> torch.Tensor:
    """ Interpolate descriptors at keypoint locations.
    keypoints (b, K, 2) (x, y); descriptors (b, 256, h, w) already L2-normalised per cell
    (re-normalising is idempotent up to rounding); returns (b, 256, K). """
    if s != 8:
        raise _lib.SfmError(-5, "sample_descriptors: only s == 8 is supported")
    if align_corners is None:
        align_corners = default_align_corners()
    kp = require_cuda(keypoints, "keypoints", torch.float32)
    d = require_cuda(descriptors, "descriptors", torch.float32)
    b, c, h, w = d.shape
    if c != 256:
        raise _lib.SfmError(-5, "sample_descriptors: descriptor_dim must be 256")
    K = kp.shape[1]
    dn = d.permute(0, 2, 3, 1).contiguous()          # NHWC view of the caller's map
    out = torch.empty(b, 256, K, device=d.device, dtype=torch.float32)
    counts = torch.full((b,), K, dtype=torch.int32, device=d.device)
    if K > 0:
        with torch.cuda.device(d.device):
            check(lib().sfm_sample_descriptors(ptr(dn), b, h, w, ptr(kp), ptr(counts), K, int(align_corners),
                                               ptr(out), stream_ptr(d.device)))
    return out


class SuperPoint:
    """SuperPoint Convolutional Detector and Descriptor (B200-native).

    Mirrors `simple_sfm.models.superpoint.SuperPoint`: same constructor arguments and the same
    `forward(data) -> {'keypoints' (K,2) f32 (x,y), 'scores' (K,), 'descriptors' (256,K)}`
    contract with `data = {'image': (1,1,H,W) f32, ['mask': (1,1,H,W) bool]}`.
    Equal scores are ordered by row-major pixel index (the reference leaves tie order unspecified).
    """

    def __init__(self, weights_path, descriptor_dim=256, nms_radius=4, keypoint_threshold=0.005,
                 max_keypoints=-1, remove_borders=4, device="cuda", precision="fp32",
                 align_corners: Optional[bool] = None):
        if descriptor_dim != 256:
            raise _lib.SfmError(-5, "descriptor_dim must be 256 (superpoint_v1 layout)")
        self.descriptor_dim = descriptor_dim
        self.nms_radius = int(nms_radius)
        self.keypoint_threshold = float(keypoint_threshold)
        self.max_keypoints = int(max_keypoints)
        self.remove_borders = int(remove_borders)
        self.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        self.align_corners = default_align_corners() if align_corners is None else bool(align_corners)
        self.training = True

        packed = pack_superpoint(load_state_dict_file(weights_path))      # strict, like load_state_dict
        mk = self.max_keypoints
        if mk == 0 or mk < -1:
            raise ValueError('\"max_keypoints\" must be positive or \"-1\"')
        self.handle = Handle(device)
        self.device = self.handle.device
        check(lib().sfm_load_superpoint(self.handle.h, packed.data_ptr(), packed.numel()))
        logger.info('Loaded SuperPoint model')

    # nn.Module-style conveniences used by the reference callers (matcher.py:48)
    def cuda(self, *a, **k):
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, data):
        return self.forward(data)

    # ------------------------------------------------------------------------------------------
    def detect(self, images: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """images (n,H,W) f32 -> per-image candidate counts (device int32); keeps dense maps in
        the handle workspace for `describe`."""
        n, H, W = images.shape
        L = lib()
        nbytes = L.sfm_superpoint_workspace_bytes(n, H, W, self.precision)
        if nbytes == 0:
            raise _lib.SfmError(-1, L.sfm_last_error().decode())
        ws = self.handle.workspace("sp", nbytes)
        cand = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(L.sfm_superpoint_detect(self.handle.h, ptr(images), n, H, W, ptr(mask), self.keypoint_threshold,
                                          self.nms_radius, self.remove_borders, self.precision, ptr(cand),
                                          ptr(ws), ws.numel(), stream_ptr(self.device)))
        return cand, ws

    def describe(self, n, H, W, kp_cap, ws, max_keypoints=None):
        mk = self.max_keypoints if max_keypoints is None else max_keypoints
        kpts = torch.empty(n, kp_cap, 2, dtype=torch.float32, device=self.device)
        scores = torch.empty(n, kp_cap, dtype=torch.float32, device=self.device)
        desc = torch.empty(n, 256, kp_cap, dtype=torch.float32, device=self.device)
        counts = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().sfm_superpoint_describe(self.handle.h, n, H, W, mk, int(self.align_corners), self.precision,
                                                kp_cap, ptr(kpts), ptr(scores), ptr(desc), ptr(counts),
                                                ptr(ws), ws.numel(), stream_ptr(self.device)))
        return kpts, scores, desc, counts

    def tap(self, n, H, W, which, ws):
        p, cnt = C.c_void_p(), C.c_size_t()
        check(lib().sfm_superpoint_tap(n, H, W, self.precision, which, ptr(ws), ws.numel(), C.byref(p), C.byref(cnt)))
        off = (p.value - ws.data_ptr())
        return ws[off:off + 4 * cnt.value].view(torch.float32)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """ Compute keypoints, scores, descriptors for image """
        image = require_cuda(data['image'], "data['image']", torch.float32)
        if image.dim() != 4 or image.shape[1] != 1:
            raise _lib.SfmError(-1, "data['image'] must be (b, 1, H, W)")
        b, _, H, W = image.shape
        mask = None
        if 'mask' in data:
            # the reference indexes `mask[0]` against the (b,H,W) score tensor (superpoint.py:134):
            # a (1,1,H,W) mask selects the whole of image 0
            m = data['mask']
            if not m.is_cuda:
                raise _lib.SfmError(-1, "data['mask'] must be a CUDA tensor")
            mask = (m[0].reshape(-1, H, W)[:1] != 0).to(torch.uint8).contiguous()
        # only batch element 0 is returned by the reference (superpoint.py:165-171)
        img0 = image[:1, 0].contiguous()
        cand, ws = self.detect(img0, mask)
        count = int(cand[0].item())                       # the single host sync of this call
        keep = count if self.max_keypoints < 0 else min(count, self.max_keypoints)
        kpts, scores, desc, _ = self.describe(1, H, W, max(keep, 1), ws)
        return {
            'keypoints': kpts[0, :keep],
            'scores': scores[0, :keep],
            'descriptors': desc[0, :, :keep],
        }
