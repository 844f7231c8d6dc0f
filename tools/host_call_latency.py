#!/usr/bin/env python
"""Wall time of the C++ host's per-segment SimulatedSensorEvaluator::computeGain (the reference's own call
pattern: one segment per call, voxel list into the segment's info) on the bench workload's map and camera,
from the host test driver's own timing lines.
    python tools/host_call_latency.py [yaml under tests/golden, default planner_naive_iterative.yaml]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mav_active_3d_planning_b200 import synth  # noqa: E402
from test_host_cpp import EXE, GOLD, write_inputs  # noqa: E402


def main():
    yaml = sys.argv[1] if len(sys.argv) > 1 else "planner_naive_iterative.yaml"
    m = synth.room_map(0.1)
    pos, quat = synth.candidate_views(m, 16, seed=42)
    with tempfile.TemporaryDirectory() as tmp:
        mp, sp = write_inputs(tmp, m, pos, quat, 1, 0)
        out = os.path.join(tmp, "out.txt")
        subprocess.run([EXE, "gains", os.path.join(GOLD, yaml), mp, sp, out], check=True, timeout=600)
        ms = [float(l.split()[3]) for l in open(out) if l.startswith("# computeGain")]
        sizes = [int(l.split()[4]) for l in open(out) if l.startswith("single")]
    print({"yaml": yaml, "calls": len(ms), "first_ms": round(ms[0], 3), "median_ms": round(sorted(ms[1:])[len(ms[1:]) // 2], 3),
           "min_ms": round(min(ms[1:]), 3), "median_list_entries": sorted(sizes)[len(sizes) // 2]})


if __name__ == "__main__":
    main()
