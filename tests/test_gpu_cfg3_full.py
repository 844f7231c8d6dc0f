"""BASELINE config 3 at FULL size on the GPU: 50x50x20 m clutter map at 0.05 m voxels (41 118 blocks,
1.3 GB mirror + 1.5 GB dense meta volume — larger than L2, the regime where the HBM roofline is
physical), 8 views x {VoxelType, Frontier, Frontier accurate}, scan-order and wavefront kernels, against
tests/golden/cfg3_goldens.npz (VoxelType rows and every ordered-list digest from oracle/_ref, the
reference's own sources; see tests/golden/make_cfg3_goldens.py)."""
import hashlib
import os

import numpy as np
import pytest

from mav_active_3d_planning_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg3_goldens.npz")
CAM = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
EVALS = {"vt": ("VoxelTypeEvaluator", {}), "fr": ("FrontierEvaluator", {}),
         "fr26": ("FrontierEvaluator", dict(accurate_frontiers=True))}


@pytest.fixture(scope="module")
def cfg3(built):
    from mav_active_3d_planning_b200 import capi
    gold = np.load(GOLD)
    m = synth.clutter_map(0.05)
    h = hashlib.sha256()
    for a in (m.block_idx, m.dist, m.observed):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest().encode() == gold["map_sha256"].tobytes(), "synthetic map differs from the one the goldens saw"
    assert m.n_blocks == int(gold["n_blocks"][0]) == 41118
    ctx = capi.Context(0)
    ctx.upload_map(m)
    pos, quat = synth.candidate_views(m, 8, seed=42)
    yield ctx, gold, pos, quat
    ctx.close()


@pytest.mark.parametrize("ek", list(EVALS))
def test_cfg3_full_size_against_reference_goldens(cfg3, ek):
    ctx, gold, pos, quat = cfg3
    caster = ctx.caster("IterativeRayCaster", **CAM)
    i = caster.info
    assert [i.c_res_x, i.c_res_y, i.c_n_sections] == gold["info"].tolist() == [158, 129, 7]
    etype, ekw = EVALS[ek]
    ev = ctx.evaluator(etype, **ekw)
    for digests in (True, False):    # scan-order kernel / wavefront kernel
        r = ctx.compute_gains(caster, ev, pos, quat, digests=digests)
        assert ctx.last_cast_kernel == (1 if digests else 2)
        for k in ("n_visible", "n_samples", "set_digest") + (("digest",) if digests else ()):
            assert np.array_equal(r[k], gold[f"{ek}/{k}"]), (ek, k, digests)
        assert np.allclose(r["gain"], gold[f"{ek}/gain"], rtol=1e-5, atol=0), (ek, digests)
    # gains and counts only: the kernel chain (a3d_split.cu)
    r = ctx.compute_gains(caster, ev, pos, quat, digests=False, set_digests=False)
    assert ctx.last_cast_kernel == 3
    for k in ("n_visible", "n_samples"):
        assert np.array_equal(r[k], gold[f"{ek}/{k}"]), (ek, k, "chain")
    assert np.allclose(r["gain"], gold[f"{ek}/gain"], rtol=1e-5, atol=0), (ek, "chain")
    assert int(gold[f"{ek}/n_samples"].max()) > 1_000_000   # full-size views: > 10^6 samples each
