#!/bin/bash
# A/B of variant_<name>.so builds on the evaluators whose leaf kernel has its own register budget:
# cfg2 with VoxelWeightEvaluator (1 / 148 / 4096 views) and FrontierEvaluator (4096 views).
# One line per point: variant, workload, evaluator, views, cast ms, per-kernel ms, sum of gains.
fmt='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        j = json.loads(l); print(sys.argv[1], j["workload"], j["evaluator"], j["views"], round(j["kernel_ms"], 3), j["stage_ms"], repr(j["gain_sum"]), flush=True)'
for f in variant_*.so; do
  v=${f#variant_}; v=${v%.so}
  A3D_LIB_PATH=$PWD/$f python tools/sweep.py --workload cfg2 --evaluator VoxelWeight --views 1,148,4096 2>/dev/null | python -c "$fmt" $v
  A3D_LIB_PATH=$PWD/$f python tools/sweep.py --workload cfg2 --evaluator Frontier --views 4096 2>/dev/null | python -c "$fmt" $v
done
