#!/bin/bash
# A/B on one box: every variant_<name>.so of the repo root (tools/build_variant.sh) on cfg2 at 1 / 148 / 4096
# views, and the variants named in $AB_WIDE on the other evaluators / configurations.  One line per point:
# variant, workload, evaluator, views, kernel ms, per-stage ms, sum of gains (must agree between variants).
fmt='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        j = json.loads(l); print(sys.argv[1], j["workload"], j["evaluator"], j["views"], round(j["kernel_ms"], 3), j["stage_ms"], repr(j["gain_sum"]), flush=True)'
for f in variant_*.so; do
  v=${f#variant_}; v=${v%.so}
  A3D_LIB_PATH=$PWD/$f python tools/sweep.py --workload cfg2 --views 1,148,4096 2>/dev/null | python -c "$fmt" $v
done
for v in $AB_WIDE; do
  f=$PWD/variant_$v.so
  A3D_LIB_PATH=$f python tools/sweep.py --workload cfg2 --evaluator Frontier --views 4096 2>/dev/null | python -c "$fmt" $v
  A3D_LIB_PATH=$f python tools/sweep.py --workload cfg2 --evaluator VoxelWeight --views 4096 2>/dev/null | python -c "$fmt" $v
  A3D_LIB_PATH=$f python tools/sweep.py --workload cfg1 --views 4096 2>/dev/null | python -c "$fmt" $v
  A3D_LIB_PATH=$f python tools/sweep.py --workload cfg4 --views 1024 2>/dev/null | python -c "$fmt" $v
done
