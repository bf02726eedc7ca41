"""CPU tests: the C-ABI library loads and exports every symbol of include/sfm_b200.h (no compute
without a GPU), weight packing is strict, host-side Matcher logic (pair lists, shard ranges) and the
multi-process gather run under gloo with world_size 2."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from simplesfm_b200 import _lib
    lib = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 25 and len(set(syms)) == len(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sfm_b200.h but not exported"
    assert sorted(syms) == sorted(_lib._SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.sfm_abi_version() == 1
    assert lib.sfm_superpoint_packed_floats() == 1300865
    assert lib.sfm_superglue_packed_floats() == 12042689


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never fall back."""
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from simplesfm_b200 import _lib, synthetic
    from simplesfm_b200.models.superglue import SuperGlue
    from simplesfm_b200.models.superpoint import SuperPoint
    h = C.c_void_p()
    assert _lib.lib().sfm_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in _lib.lib().sfm_last_error()
    with pytest.raises(_lib.SfmError):
        SuperPoint(synthetic.superpoint_state_dict(0))
    with pytest.raises(_lib.SfmError):
        SuperGlue(synthetic.superglue_state_dict(0))


def test_workspace_queries_are_shape_functions():
    from simplesfm_b200 import _lib
    L = _lib.lib()
    a = L.sfm_superglue_workspace_bytes(5, 2048, 2048, 1, 1)
    b = L.sfm_superglue_workspace_bytes(40, 2048, 2048, 1, 8)
    assert 0 < a < b
    assert L.sfm_superglue_workspace_bytes(5, 2048, 2048, 0, 1) > a          # fp32 path materialises attention
    assert L.sfm_superglue_workspace_bytes(5, 0, 10, 0, 1) == 0               # invalid shape -> 0 + error text
    assert b"K0, K1 > 0" in L.sfm_last_error()
    assert L.sfm_superpoint_workspace_bytes(1, 480, 640, 0) > 2 * 480 * 640 * 64 * 4
    assert L.sfm_superpoint_workspace_bytes(1, 481, 640, 0) > 0               # fp32 path: any size >= 8, like the reference
    assert L.sfm_superpoint_workspace_bytes(1, 481, 640, 1) == 0              # tcgen05 path: multiples of 8
    assert L.sfm_superpoint_workspace_bytes(1, 7, 640, 0) == 0
    assert L.sfm_sinkhorn_workspace_bytes(2, 100, 90) > 0


def test_weight_packing_strict(sp_weights, sg_weights):
    from simplesfm_b200 import weights
    p = weights.pack_superpoint(sp_weights)
    assert p.numel() == 1300865 and p.dtype == torch.float32
    assert torch.equal(p[:576], sp_weights["conv1a.weight"].reshape(-1))
    q = weights.pack_superglue(sg_weights)
    assert q.numel() == 12042689 and float(q[0]) == float(sg_weights["bin_score"])
    assert len(weights.superglue_keys()) == 339 and len(weights.superpoint_keys()) == 24
    bad = dict(sg_weights)
    bad.pop("final_proj.bias")
    with pytest.raises(RuntimeError):
        weights.pack_superglue(bad)
    bad = dict(sp_weights)
    bad["extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        weights.pack_superpoint(bad)


def test_shard_ranges_cover_and_align():
    from simplesfm_b200.distributed import image_block, shard_range
    for n, chunk, world in ((1225, 5, 8), (1225, 5, 3), (7, 4, 2), (0, 5, 4), (499500, 5, 8), (3, 5, 8)):
        spans = [shard_range(n, chunk, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        for a, b in spans:
            assert a % chunk == 0 or a == n           # chunk aligned => identical chunk composition
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= chunk
    blocks = [image_block(50, r, 8) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 50 and all(b == c for (_, b), (c, _) in zip(blocks, blocks[1:]))


def test_install_as_simple_sfm_aliases():
    import simplesfm_b200
    simplesfm_b200.install_as_simple_sfm()
    from simple_sfm.matchers import Matcher
    from simple_sfm.matchers.matcher import Frame
    from simple_sfm.models.superglue import SuperGlue, log_optimal_transport
    from simple_sfm.models.superpoint import SuperPoint, simple_nms
    import simplesfm_b200.matchers.matcher as mm
    assert Matcher is mm.Matcher and Frame is mm.Frame
    assert Frame._fields == ("descriptors", "keypoints_poses", "scores", "image")
    import inspect
    params = list(inspect.signature(Matcher.__init__).parameters)
    assert params[1:9] == ["super_point_extractor_weights_path", "nms_radius", "keypoint_threshold", "matcher_type",
                           "super_glue_weights_path", "match_threshold", "sinkhorn_iterations", "super_glue_batch"]
    d = inspect.signature(Matcher.__init__).parameters
    assert (d["matcher_type"].default, d["match_threshold"].default, d["sinkhorn_iterations"].default,
            d["super_glue_batch"].default) == ("simple", 0.4, 20, 5)
    s = inspect.signature(SuperGlue.__init__).parameters
    assert (s["sinkhorn_iterations"].default, s["match_threshold"].default) == (100, 0.2)
    p = inspect.signature(SuperPoint.__init__).parameters
    assert (p["nms_radius"].default, p["keypoint_threshold"].default, p["max_keypoints"].default,
            p["remove_borders"].default) == (4, 0.005, -1, 4)
    for k in ("simple_sfm.matchers", "simple_sfm.matchers.matcher", "simple_sfm.matchers.simple_matcher",
              "simple_sfm.models.superpoint", "simple_sfm.models.superglue", "simple_sfm.models", "simple_sfm"):
        sys.modules.pop(k, None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from simplesfm_b200.distributed import (CompactTables, active_group, allgather_records, assert_same_everywhere,
                                            concat_tables, gather_tables_to_root, shard_range)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cpu = torch.device("cpu")
        group, r, w = active_group(True)
        ok = (r, w) == (rank, world) and active_group(False) == (None, 0, 1)
        # ---- phase 4: compact tables of this rank's chunk-aligned pair range -> rank 0 only
        ids = list(range(1, 8))
        pairs = np.array([(a, b) for i, a in enumerate(ids) for b in ids[i + 1:]], dtype=np.int64)
        lo, hi = shard_range(len(pairs), 4, rank, world)
        rng = np.random.RandomState(0)
        full = {(np.int64(a), np.int64(b)): rng.randint(0, 500, size=(int(a * 3 + b) % 9, 2)).astype(np.uint32)
                for a, b in pairs}
        keys = list(full)
        local = CompactTables(pairs[lo:hi], np.array([len(full[k]) for k in keys[lo:hi]], dtype=np.int64),
                              np.concatenate([full[k] for k in keys[lo:hi]]).reshape(-1, 2))
        merged = gather_tables_to_root(local, pairs, 4, cpu, group)
        if rank == 0:
            table = merged.to_dict()
            ok = ok and list(table.keys()) == keys and all(np.array_equal(table[k], full[k]) and
                                                           table[k].dtype == np.uint32 for k in full)
            ok = ok and all(isinstance(k[0], np.int64) for k in table)
        else:
            ok = ok and merged is None
        # an empty shard (fewer chunks than ranks) still takes part in the collectives
        tiny = pairs[:3]
        lo, hi = shard_range(len(tiny), 4, rank, world)
        loc = CompactTables(tiny[lo:hi], np.array([len(full[k]) for k in keys[:3]][lo:hi], dtype=np.int64),
                            (np.concatenate([full[k] for k in keys[:3]]) if hi > lo else np.zeros((0, 2), np.uint32)
                             ).reshape(-1, 2))
        merged = gather_tables_to_root(loc, tiny, 4, cpu, group)
        if rank == 0:
            t3 = merged.to_dict()
            ok = ok and list(t3.keys()) == keys[:3] and all(np.array_equal(t3[k], full[k]) for k in keys[:3])
        ok = ok and concat_tables([]).to_dict() == {}
        # ---- phase 2: rank r owns images r, r + world, ...; ragged keypoint counts, mixed image sizes
        n_img = 5
        g = torch.Generator().manual_seed(1)
        counts = [6, 0, 3, 5, 1]
        hw = [(480, 640), (480, 640), (240, 320), (480, 640), (120, 160)]
        desc = [torch.randn(256, k, generator=g) for k in counts]
        kp = [torch.randn(k, 2, generator=g) for k in counts]
        sc = [torch.rand(k, generator=g) for k in counts]
        mine = [(desc[i], kp[i], sc[i], hw[i]) for i in range(rank, n_img, world)]
        rec = allgather_records(mine, None, cpu, group)
        ok = ok and rec.counts == counts and rec.hw == hw and rec.desc.shape == (n_img, 256, 8)
        for i, k in enumerate(counts):
            ok = ok and torch.equal(rec.desc[i, :, :k], desc[i]) and torch.equal(rec.kpts[i, :k], kp[i]) \
                and torch.equal(rec.scores[i, :k], sc[i])
        # ---- ranks handed different scenes are detected
        assert_same_everywhere([n_img, len(pairs)], cpu, group)
        try:
            assert_same_everywhere([n_img + rank], cpu, group)
            ok = False
        except RuntimeError:
            pass
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_tables_and_features():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_compact_tables_roundtrip():
    from simplesfm_b200.distributed import CompactTables, concat_tables
    a = CompactTables(np.array([[1, 2], [1, 3]], np.int64), np.array([2, 0], np.int64),
                      np.array([[0, 5], [3, 1]], np.uint32))
    b = CompactTables(np.array([[2, 3]], np.int64), np.array([1], np.int64), np.array([[7, 7]], np.uint32))
    d = concat_tables([a, b]).to_dict()
    assert list(d.keys()) == [(1, 2), (1, 3), (2, 3)]
    assert d[(1, 2)].tolist() == [[0, 5], [3, 1]] and d[(1, 3)].shape == (0, 2) and d[(2, 3)].tolist() == [[7, 7]]
    assert all(v.dtype == np.uint32 for v in d.values())


def test_pair_list_matches_itertools_and_vectorised_filters():
    """Matcher.pair_list == filter(f, itertools.combinations(ids, 2)) (matcher.py:143-144) for array-style filters
    (evaluated once), per-pair filters, constant filters and filters that cannot take arrays."""
    import itertools
    import numpy as np
    from simplesfm_b200.matchers.matcher import Matcher
    ids = [np.int64(i) for i in range(1, 41)]
    allowed = {(1, 2), (3, 9), (39, 40)}
    filters = [None,
               lambda x: np.abs(x[0] - x[1]) < 5,                     # generate_..._from_scene.py:141
               lambda x: True,
               lambda x: False,
               lambda x: x[0] == 1 and x[1] < 7,                      # raises on arrays -> per pair
               lambda x: (int(x[0]), int(x[1])) in allowed,           # scalar-only -> per pair
               lambda x: (x[0] + x[1]) % 3 == 0]
    for f in filters:
        every = list(itertools.combinations(ids, 2))
        ref = np.array([p for p in every if (f is None or f(np.array(p)))]).reshape(-1, 2)
        got = Matcher.pair_list(ids, f)
        assert got.shape == ref.shape and np.array_equal(got, ref)
    assert Matcher.pair_list([np.int64(1)], lambda x: True).shape == (0, 2)
    # a filter whose array answer disagrees with its per-pair answer must not be trusted
    tricky = lambda x: (np.asarray(x[0]) > 0) if np.ndim(x[0]) else (x[0] % 2 == 0)
    ref = np.array([p for p in itertools.combinations(ids, 2) if tricky(np.array(p))]).reshape(-1, 2)
    assert np.array_equal(Matcher.pair_list(ids, tricky), ref)
    # a filter with a GLOBAL reduction: its array form differs from its per-pair form on pairs a
    # small sample would miss; up to 20 000 pairs every pair is compared, so the per-pair answer wins
    glob = lambda x: np.abs(x[0] - x[1]) < (np.max(x[0]) if np.ndim(x[0]) else 1000) - 36
    ref = np.array([p for p in itertools.combinations(ids, 2) if glob(np.array(p))]).reshape(-1, 2)
    assert np.array_equal(Matcher.pair_list(ids, glob), ref)
    # a filter that declares itself element-wise is evaluated once and not re-checked
    calls = []
    def banded(x):
        calls.append(np.ndim(x[0]))
        return np.abs(x[0] - x[1]) < 3
    banded.vectorized = True
    got = Matcher.pair_list(ids, banded)
    assert calls == [1] and len(got) == sum(1 for a, b in itertools.combinations(range(40), 2) if b - a < 3)


def test_balanced_launches():
    """Fused chunk runs are cut into the fewest launches the limit allows, equal to within one chunk."""
    from simplesfm_b200.matchers.matcher import Matcher
    assert Matcher.balanced_launches(245, 51) == [49] * 5
    assert Matcher.balanced_launches(245, 32) == [31] * 5 + [30] * 3
    assert Matcher.balanced_launches(51, 51) == [51] and Matcher.balanced_launches(52, 51) == [26, 26]
    assert Matcher.balanced_launches(1, 51) == [1] and Matcher.balanced_launches(0, 51) == []
    for n in range(1, 400):
        for lim in (1, 7, 51):
            parts = Matcher.balanced_launches(n, lim)
            assert sum(parts) == n and max(parts) <= lim and max(parts) - min(parts) <= 1
            assert len(parts) == -(-n // lim)


def test_compact_tables_to_dict_views_and_empties():
    """CompactTables.to_dict: pair order, np.int64 keys like the reference's table, zero-length tables, empty set."""
    import numpy as np
    from simplesfm_b200.distributed import CompactTables, concat_tables
    pairs = np.array([[1, 2], [1, 3], [2, 3]], np.int64)
    lens = np.array([2, 0, 3], np.int64)
    rows = np.arange(10, dtype=np.uint32).reshape(5, 2)
    d = CompactTables(pairs, lens, rows).to_dict()
    assert list(d.keys()) == [(1, 2), (1, 3), (2, 3)] and all(isinstance(k[0], np.int64) for k in d)
    assert d[(1, 2)].tolist() == [[0, 1], [2, 3]] and d[(1, 3)].shape == (0, 2) and d[(2, 3)].tolist() == [[4, 5], [6, 7], [8, 9]]
    assert d[(2, 3)].dtype == np.uint32 and np.shares_memory(d[(2, 3)], rows)
    assert concat_tables([]).to_dict() == {}
    both = concat_tables([CompactTables(pairs[:1], lens[:1], rows[:2]), CompactTables(pairs[1:], lens[1:], rows[2:])])
    assert {k: v.tolist() for k, v in both.to_dict().items()} == {k: v.tolist() for k, v in d.items()}


def test_sparse_depth_entries_edge_cases():
    """camera_point_entries: image ids outside the camera set are ignored, nothing selected -> empty arrays,
    a point observed twice by the same camera counts once."""
    import numpy as np
    from simplesfm_b200.dataset import simple_sfm_dataset_utils as U
    err = np.array([0.5, 0.5, 5.0])
    tp = np.array([0, 0, 0, 1, 2, 2])
    ti = np.array([1, 1, 9, 2, 1, 2])
    cam, pt = U.camera_point_entries(err, tp, ti, [0, 1], 2.0)
    assert cam.tolist() == [0, 1] and pt.tolist() == [0, 1]          # point 2 fails the error test, image 9 is unknown
    cam, pt = U.camera_point_entries(err, tp, ti, [5, 6], 2.0)
    assert len(cam) == 0 and len(pt) == 0
    cam, pt = U.camera_point_entries(np.zeros(0), np.zeros(0, np.int64), np.zeros(0, np.int64), [0], 2.0)
    assert len(cam) == 0
