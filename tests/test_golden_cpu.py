"""The CPU oracle (the restatement the GPU suite checks against) reproduces the committed
reference outputs tests/golden/ref_goldens.npz, which oracle/_ref — the reference's own sources —
generated (tests/golden/make_ref_goldens.py).  Runs without /root/reference."""
import os

import numpy as np
import pytest

import golden_cases as G
from mav_active_3d_planning_b200 import synth
from oracle import oracle_py as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_goldens.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", G.cases(), ids=[c[0] for c in G.cases()])
def test_oracle_reproduces_reference_goldens(built, gold, case, room02, room01):
    name, mk, kind, cam, spec, evs, variant = case
    m = {"room02": room02, "room01": room01}[mk]
    pos, quat = G.poses(synth, m, spec)
    c = O.OracleMap.from_synth(m).caster(kind, **cam)
    assert [c.c_res_x, c.c_res_y, c.c_n_sections] == gold[f"{name}/info"].tolist()
    nv0 = int(np.sum(G.MULTI_VSEG == 0)) if variant == "multi" else 1
    lst, ns = c.visible_voxels(pos[:nv0], quat[:nv0])
    assert ns == int(gold[f"{name}/list0_samples"][0])
    assert O.digest(lst) == int(gold[f"{name}/list0_digest"][0])
    if name in G.LIST_CASES:
        assert np.array_equal(lst, gold[f"{name}/list0"])
    for ek in evs:
        etype, ekw = G.EVALS[ek]
        e = O.eval_params(etype, clear_from_parents=(variant == "chain"), **ekw)
        r = c.compute_gains(e, pos, quat, **G.call_kwargs(variant))
        for k in ("gain", "n_visible", "n_samples", "digest", "set_digest"):
            assert np.array_equal(r[k], gold[f"{name}/{ek}/{k}"]), (name, ek, k)
