"""Randomised configurations (tools/fuzz_wave_vs_scan.py): voxel size, voxels per side, camera, mounting,
caster, evaluator, bounding volume, map offset, segment layout.  The scan-order kernel against the CPU
oracle, and the wavefront kernel (both CTA widths, gains and ordered lists) against the scan-order
kernel."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_random_configurations_against_oracle_and_between_kernels(built):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_wave_vs_scan as fz
    from oracle import oracle_py as O
    rng = np.random.default_rng(2026)
    for case in range(14):
        r = fz.one_case(rng, case, oracle=O if case < 8 else None, max_views=5 if case < 8 else None)
        assert not r["bad"], r
