"""simplesfm_b200 -- B200-native (sm_100a) SuperPoint -> SuperGlue matching path behind the
`simple_sfm.matchers` API of palsol/SimpleSfm.

Public surface (mirrors the reference module paths):
    simplesfm_b200.matchers.Matcher / matchers.matcher.Frame / matchers.simple_matcher.SimpleMatcher
    simplesfm_b200.models.superpoint.SuperPoint, simple_nms, remove_borders, top_k_keypoints, sample_descriptors
    simplesfm_b200.models.superglue.SuperGlue, normalize_keypoints, log_optimal_transport, ...
All compute runs in libsfm_b200.so (hand-written CUDA, C ABI in include/sfm_b200.h); there is
no CPU fallback and no multi-backend dispatch.
"""
__version__ = "0.1.0"


def install_as_simple_sfm() -> None:
    """Alias the mirrored modules under the reference's import paths
    (`simple_sfm.matchers`, `simple_sfm.models.superpoint`, `simple_sfm.models.superglue`) so that
    callers such as `scripts/generate_simple_sfm_dataset_from_scene.py:111` pick up this
    implementation without edits.  See INTEGRATION.md."""
    import importlib
    import sys
    import types
    from . import matchers, models
    from .matchers import matcher, simple_matcher
    from .models import superglue, superpoint
    try:
        pkg = importlib.import_module("simple_sfm")
    except ImportError:
        pkg = types.ModuleType("simple_sfm")
        pkg.__path__ = []
        sys.modules["simple_sfm"] = pkg
    for name, mod in (("simple_sfm.matchers", matchers), ("simple_sfm.matchers.matcher", matcher),
                      ("simple_sfm.matchers.simple_matcher", simple_matcher),
                      ("simple_sfm.models.superpoint", superpoint), ("simple_sfm.models.superglue", superglue)):
        sys.modules[name] = mod
    if "simple_sfm.models" not in sys.modules:
        sys.modules["simple_sfm.models"] = models
    pkg.matchers = matchers
    # post-match consumer: only the one function is replaced -- the reference module holds other,
    # out-of-scope generators that must stay reachable
    from .dataset import simple_sfm_dataset_utils as depth_utils
    try:
        ref_utils = importlib.import_module("simple_sfm.dataset.simple_sfm_dataset_utils")
        ref_utils.generate_sparse_depth_from_colmap = depth_utils.generate_sparse_depth_from_colmap
    except Exception:
        sys.modules.setdefault("simple_sfm.dataset.simple_sfm_dataset_utils", depth_utils)
