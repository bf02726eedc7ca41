import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.fixture(scope="session")
def sp_weights():
    from simplesfm_b200 import synthetic
    return synthetic.superpoint_state_dict(0)


@pytest.fixture(scope="session")
def sg_weights():
    from simplesfm_b200 import synthetic
    return synthetic.superglue_state_dict(0, "indoor")


@pytest.fixture(scope="session")
def sg_weights_outdoor():
    from simplesfm_b200 import synthetic
    return synthetic.superglue_state_dict(1, "outdoor")


def canon_keypoints(kpts, scores, W=100000):
    """Order-independent view of a keypoint set: rows sorted by (score desc, pixel index asc)."""
    kpts = np.asarray(kpts)
    scores = np.asarray(scores)
    pix = kpts[:, 1].astype(np.int64) * W + kpts[:, 0].astype(np.int64)
    order = np.lexsort((pix, -scores.astype(np.float64)))
    return kpts[order], scores[order], order
