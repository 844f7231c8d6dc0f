"""The multi-GPU group of the C ABI (a3d_group_*, single process, ncclCommInitAll): results must not
depend on how many GPUs share the work.  N = 1 runs on any box; N >= 2 needs as many devices."""
import numpy as np
import pytest

from conftest import CAM_BASE
from mav_active_3d_planning_b200 import synth

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count()


def run_sequence(capi, devices, m, pos, quat, vseg, nseg):
    """cfg5 in small: full map, then two block deltas (one with removals), gains after each."""
    g = capi.Group(devices)
    g.upload_map(m)
    g.set_option("deal_block", 3)
    caster = g.caster("IterativeRayCaster", **CAM_BASE)
    out = []
    rng = np.random.default_rng(17)
    for step in range(3):
        if step:
            sel = rng.choice(m.n_blocks, m.n_blocks // 10, replace=False)
            nd = (m.dist[sel] + rng.normal(0, 0.03, m.dist[sel].shape)).astype(np.float32)
            no = np.maximum(m.observed[sel], rng.random(m.observed[sel].shape) < 0.3).astype(np.uint8)
            g.upload_blocks(m.block_idx[sel], nd, no)          # ESDF-only delta: weights stay
            if step == 2:
                g.remove_blocks(m.block_idx[rng.choice(m.n_blocks, 4, replace=False)])
        for ek, kw in (("VoxelTypeEvaluator", dict(gain_free=0.5)), ("FrontierEvaluator", {}), ("VoxelWeightEvaluator", {})):
            ev = g.evaluator(ek, **kw)
            dig = ek != "VoxelWeightEvaluator"
            out.append(g.compute_gains(caster, ev, pos, quat, vseg, nseg, digests=dig, set_digests=dig))
    launches = g.kernel_launches
    g.close()
    return out, launches


def reference_sequence(capi, m, pos, quat, vseg, nseg):
    return run_sequence(capi, [0], m, pos, quat, vseg, nseg)


@pytest.fixture(scope="module")
def workload(room02):
    pos, quat = synth.candidate_views(room02, 41, seed=8)
    vseg = np.sort(np.random.default_rng(2).integers(0, 29, 41)).astype(np.int32)   # ragged, some empty
    return room02, pos, quat, vseg, 30


def test_group_of_one_matches_context(built, workload):
    from mav_active_3d_planning_b200 import capi
    m, pos, quat, vseg, nseg = workload
    got, launches = run_sequence(capi, [0], m, pos, quat, vseg, nseg)
    assert launches > 0
    ctx = capi.Context(0)
    ctx.upload_map(m)
    caster, ev = ctx.caster("IterativeRayCaster", **CAM_BASE), ctx.evaluator("VoxelTypeEvaluator", gain_free=0.5)
    want = ctx.compute_gains(caster, ev, pos, quat, vseg, nseg)
    for k in ("gain", "n_visible", "n_samples", "digest", "set_digest"):
        assert np.array_equal(got[0][k], want[k]), k
    ctx.close()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_group_results_do_not_depend_on_gpu_count(built, workload, n):
    from mav_active_3d_planning_b200 import capi
    if n_devices() < n:
        pytest.skip(f"needs {n} GPUs")
    m, pos, quat, vseg, nseg = workload
    one, _ = reference_sequence(capi, m, pos, quat, vseg, nseg)
    many, _ = run_sequence(capi, list(range(n)), m, pos, quat, vseg, nseg)
    assert len(one) == len(many) == 9
    for a, b in zip(one, many):
        for k in ("n_visible", "n_samples", "digest", "set_digest"):
            if a[k] is not None:
                assert np.array_equal(a[k], b[k]), k
        # VoxelType / Frontier gains are sums of integers times constants: identical; VoxelWeight: 1e-5
        assert np.allclose(a["gain"], b["gain"], rtol=1e-5, atol=0)
