"""The wavefront kernel's "certified sample" claim (a3d_wave.cu sample_prepare), checked on the CPU:
whenever frac(pf / voxel_size) keeps more than delta from 0, 1/2 and 1 on an axis, voxblox's own
fp32 index arithmetic lands on block/voxel floor(T), getVoxelCenter's voxel
(double subtraction) is the same one, and the sign of pf - centre is the one frac(T) predicts.
numpy float32 arithmetic is IEEE round-to-nearest without contraction, like the device intrinsics."""
import numpy as np
import pytest

F = np.float32
EPS = F(1e-6)
MAGIC = F(12582912.0)


def literal(p64, vs, vps):
    """(block*vps + voxel from getVoxelState's arithmetic, same from getVoxelCenter's, below-centre flag,
    in-range flag) per axis, exactly as the reference computes them."""
    vs = F(vs)
    vs_inv = F(1.0 / float(vs))
    bs = F(vs * F(vps))
    bs_inv = F(1.0 / float(bs))
    pf = p64.astype(F)
    bi = np.floor(pf * bs_inv + EPS).astype(np.int64)
    origin = (bi.astype(F) * bs).astype(F)
    v_state = np.floor((pf - origin).astype(F) * vs_inv + EPS).astype(np.int64)
    v_centre = np.floor((p64 - origin.astype(np.float64)).astype(F) * vs_inv + EPS).astype(np.int64)
    vc = np.clip(v_state, 0, vps - 1)
    centre = (origin + ((vc.astype(F).astype(np.float64) + 0.5) * float(vs)).astype(F)).astype(F)
    below = (pf - centre).astype(F) < 0
    in_range = (v_state >= 0) & (v_state < vps)
    return bi * vps + v_state, bi * vps + v_centre, below, in_range


def certified(p64, vs, vps, t_max):
    vs = F(vs)
    vs_inv = F(1.0 / float(vs))
    pf = p64.astype(F)
    T = (pf * vs_inv).astype(F)
    s = (T + MAGIC).astype(F)
    fr = (T - (s - MAGIC).astype(F)).astype(F)
    delta = F(F(t_max * 5e-7) + F((vps + 1) * (16.0 * 5.97e-8 + 1.1e-6)))
    af = np.abs(fr)
    safe = (af > delta) & (af < F(0.5) - delta)
    below_int = (fr < 0)
    g = (s.view(np.int32).astype(np.int64) - 0x4B400000) - below_int.astype(np.int64)
    return safe, g, ~below_int, delta


def certified_fma(P0, direction, d, vs, vps, t_max):
    """The kernel's form of the test: half-voxel coordinate z = fma(float(d), dz, z0) in fp32, without
    the reference's point.  (fp32 fma emulated in float64: the product of two floats is exact there.)"""
    vs_inv = float(F(1.0 / float(F(vs))))
    dz = (direction * (2.0 * vs_inv)).astype(F)
    z0 = (P0 * (2.0 * vs_inv) - 0.5).astype(F)
    z = (d.astype(F).astype(np.float64)[:, None] * dz.astype(np.float64) + z0.astype(np.float64)).astype(F)
    s = (z + MAGIC).astype(F)
    fr = (z - (s - MAGIC).astype(F)).astype(F)
    lim = F(0.5) - F(2.0) * F(F(t_max * 9e-7) + F((vps + 1) * (16.0 * 5.97e-8 + 1.1e-6)))
    safe = np.all(np.abs(fr) < lim, axis=1)
    h = s.view(np.int32).astype(np.int64) - 0x4B400000
    return safe, h >> 1, (h & 1) == 0, lim


@pytest.mark.parametrize("vs,vps", [(0.1, 16), (0.2, 16), (0.05, 16), (0.2, 8), (0.1, 32), (0.37, 4)])
@pytest.mark.parametrize("extent", [8.0, 60.0, 700.0, 9000.0])
def test_fma_certified_samples_agree_with_the_literal_index_arithmetic(vs, vps, extent):
    """Samples generated the way the caster generates them (camera position + distance * direction,
    distance accumulated in steps of the voxel size); a share of the rays is aimed so that samples hug
    voxel faces / centres."""
    rng = np.random.default_rng(int(vs * 1000) + 7 * vps + int(extent))
    n = 600_000
    vsd = float(F(vs))
    ray_length = 8.0
    P0 = (rng.random((n, 3)) * 2.0 - 1.0) * extent
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1)[:, None]
    # axis-aligned and nearly axis-aligned rays from positions on / next to faces and centres
    k = n // 3
    direction[:k] = np.eye(3)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], k)[:, None]
    direction[:k] += rng.normal(size=(k, 3)) * rng.choice([0.0, 1e-7, 1e-4], k)[:, None]
    snap = np.round(P0[:k] / vsd * 2.0) / 2.0 * vsd
    P0[:k] = snap + rng.integers(-40, 41, (k, 3)) * np.spacing(np.abs(snap).astype(F)).astype(np.float64) * 0.25
    d = rng.integers(0, int(ray_length / vsd), n) * vsd + rng.integers(-3, 4, n) * 1e-15
    pts = P0 + d[:, None] * direction  # fp64 multiply, then add: the reference's point
    t_max = (extent + ray_length) * float(F(1.0 / vsd)) * 1.001 + 2.0
    safe, g, below, lim = certified_fma(P0, direction, d, vs, vps, t_max)
    g_state, g_centre, below_lit, in_range = literal(pts, vs, vps)
    if lim > 0:
        assert safe.any()
    assert (~safe).any()
    assert np.all(in_range[safe]), "certified sample with an artefact voxel index"
    assert np.array_equal(g[safe], g_state[safe]), "block/voxel of getVoxelState differs"
    assert np.array_equal(g[safe], g_centre[safe]), "voxel of getVoxelCenter differs"
    assert np.array_equal(below[safe], below_lit[safe]), "interpolation octant differs"
    if extent <= 60.0:
        assert safe[k:].mean() > 0.97
    # how much of the margin the worst observed disagreement uses (must stay well below 1)
    bad = np.any((g != g_state) | (g != g_centre) | (below != below_lit), axis=1)
    if bad.any():
        vs_inv = float(F(1.0 / vsd))
        y = pts * (2.0 * vs_inv)
        dist = np.abs(y - np.round(y)).min(axis=1) / 2.0  # distance of T from 0, 1/2, 1 (voxel units)
        assert dist[bad].max() < 0.5 * (0.5 - float(lim)) / 2.0 + 1e-12


@pytest.mark.parametrize("vs,vps", [(0.1, 16), (0.2, 16), (0.05, 16), (0.2, 8), (0.1, 32), (0.37, 4)])
@pytest.mark.parametrize("extent", [8.0, 60.0, 700.0, 9000.0])
def test_certified_samples_agree_with_the_literal_index_arithmetic(vs, vps, extent):
    rng = np.random.default_rng(int(vs * 1000) + vps + int(extent))
    n = 400_000
    vsd = float(F(vs))
    # uniform points, plus points hugging voxel faces and voxel centres within a few fp32 ulps
    p = (rng.random(n) * 2.0 - 1.0) * extent
    k = rng.integers(-int(extent / vsd), int(extent / vsd), n)
    ulp = np.spacing(np.abs(k * vsd).astype(F)).astype(np.float64)
    faces = k * vsd + rng.integers(-40, 41, n) * ulp * 0.25
    centres = (k + 0.5) * vsd + rng.integers(-40, 41, n) * ulp * 0.25
    near = k * vsd + (rng.random(n) - 0.5) * 4e-3 * vsd
    pts = np.concatenate([p, faces, centres, near])
    t_max = (extent + 5.0) * float(F(1.0 / vsd)) * 1.001 + 2.0
    safe, g, below, delta = certified(pts, vs, vps, t_max)
    g_state, g_centre, below_lit, in_range = literal(pts, vs, vps)
    assert safe.any() and (~safe).any()
    ok = safe
    assert np.all(in_range[ok]), "certified sample with an artefact voxel index"
    assert np.array_equal(g[ok], g_state[ok]), "block/voxel of getVoxelState differs"
    assert np.array_equal(g[ok], g_centre[ok]), "voxel of getVoxelCenter differs"
    assert np.array_equal(below[ok], below_lit[ok]), "interpolation octant differs"
    # the margin is not vacuous: it certifies the bulk of ordinary samples at room scale
    if extent <= 60.0:
        assert safe[:n].mean() > 0.99
    # and it is needed: without it the shortcut does disagree near faces / centres
    bad = (g != g_state) | (g != g_centre) | (below != below_lit)
    assert not np.any(bad & ok)
    if extent >= 60.0:
        assert np.any(bad & ~ok)
