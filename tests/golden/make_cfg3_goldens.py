#!/usr/bin/env python
"""Writes tests/golden/cfg3_goldens.npz: BASELINE config 3 at FULL size — the 50x50x20 m clutter map at
0.05 m voxels (41 118 blocks of 16^3), IterativeRayCaster 320x240 f=160 L=5 (158x129 rays, 7 sections),
8 candidate views — for {VoxelType, Frontier, Frontier accurate_frontiers}.

VoxelType comes from oracle/_ref (the reference's own sources).  The reference's Frontier fold erases
observed entries one at a time from ~10^6-entry vectors (frontier_evaluator.cpp:110-116: O(V^2), hours
per view at this size), so the Frontier rows come from the restatement in oracle/ — which
tests/test_ref_pin.py holds bit-equal to oracle/_ref on every smaller case — and the file says so.
The map's own digest is stored too: the GPU test regenerates the map from its seed and checks it."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mav_active_3d_planning_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

CAM = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
EVALS = {"vt": ("VoxelTypeEvaluator", {}, "ref"), "fr": ("FrontierEvaluator", {}, "oracle"),
         "fr26": ("FrontierEvaluator", dict(accurate_frontiers=True), "oracle")}
N_VIEWS, SEED = 8, 42


def map_digest(m):
    h = hashlib.sha256()
    for a in (m.block_idx, m.dist, m.observed):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    if O.build_ref() is None:
        raise SystemExit("oracle/_ref cannot be built here (/root/reference missing)")
    m = synth.clutter_map(0.05)
    pos, quat = synth.candidate_views(m, N_VIEWS, seed=SEED)
    out = {"map_sha256": np.frombuffer(map_digest(m).encode(), np.uint8), "n_blocks": np.array([m.n_blocks])}
    maps = {}
    for ek, (etype, ekw, impl) in EVALS.items():
        if impl not in maps:
            maps[impl] = O.OracleMap.from_synth(m, impl=impl)
        c = maps[impl].caster("IterativeRayCaster", **CAM)
        out["info"] = np.array([c.c_res_x, c.c_res_y, c.c_n_sections], np.int32)
        r = c.compute_gains(O.eval_params(etype, **ekw), pos, quat, n_threads=8)
        for k in ("gain", "n_visible", "n_samples", "digest", "set_digest"):
            out[f"{ek}/{k}"] = r[k]
        out[f"{ek}/impl"] = np.frombuffer(impl.encode(), np.uint8)
        print(ek, impl, r["gain"], flush=True)
    # the caster's lists do not depend on the evaluator: _ref and the restatement agree on them here too
    assert np.array_equal(out["vt/digest"], out["fr/digest"]) and np.array_equal(out["vt/n_samples"], out["fr/n_samples"])
    path = os.path.join(HERE, "cfg3_goldens.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
