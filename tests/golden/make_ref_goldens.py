#!/usr/bin/env python
"""Writes tests/golden/ref_goldens.npz from oracle/_ref — the reference's own UNMODIFIED sources
compiled behind the orc_* API (oracle/Makefile target `_ref`).  Needs /root/reference, so it runs
in the build container only; the .npz travels to the GPU box, where tests/test_gpu_golden.py
demands the CUDA path reproduce it.

    python tests/golden/make_ref_goldens.py

Per case and evaluator: gain, n_visible, n_samples, ordered digest and set digest per segment; for
the first segment of each case also the ordered visible-voxel list itself (float64 centres).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_cases as G  # noqa: E402
from mav_active_3d_planning_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def main():
    if O.build_ref() is None:
        raise SystemExit("oracle/_ref cannot be built here (/root/reference missing)")
    assert O.lib("ref").orc_impl() == b"reference"
    maps = {"room02": synth.room_map(0.2), "room01": synth.room_map(0.1)}
    ref = {k: O.OracleMap.from_synth(m, impl="ref") for k, m in maps.items()}
    out = {}
    for name, mk, kind, cam, spec, evs, variant in G.cases():
        pos, quat = G.poses(synth, maps[mk], spec)
        c = ref[mk].caster(kind, **cam)
        out[f"{name}/info"] = np.array([c.c_res_x, c.c_res_y, c.c_n_sections], np.int32)
        nv0 = int(np.sum(G.MULTI_VSEG == 0)) if variant == "multi" else 1
        lst, ns = c.visible_voxels(pos[:nv0], quat[:nv0])
        if name in G.LIST_CASES:  # the digests pin every list; a few are kept in full
            out[f"{name}/list0"] = lst
        out[f"{name}/list0_digest"] = np.array([O.digest(lst)], np.uint64)
        out[f"{name}/list0_samples"] = np.array([ns], np.int64)
        for ek in evs:
            etype, ekw = G.EVALS[ek]
            e = O.eval_params(etype, clear_from_parents=(variant == "chain"), **ekw)
            r = c.compute_gains(e, pos, quat, **G.call_kwargs(variant))
            for k in ("gain", "n_visible", "n_samples", "digest", "set_digest"):
                out[f"{name}/{ek}/{k}"] = r[k]
        print(name, "ok", flush=True)
    path = os.path.join(HERE, "ref_goldens.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
