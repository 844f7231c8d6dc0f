#!/bin/bash
# A/B of kernel tuning macros on one box: benchmarks every variant_<name>.so in the repo root (each
# built from the csrc/*.cu with different -DA3D_WAVE_* values, same flags as __graft_entry__.build_cuda)
# on the bench workload at 1 / 148 / 4096 views per launch.  Prints kernel milliseconds.
for f in variant_*.so; do
  v=${f#variant_}; v=${v%.so}
  A3D_LIB_PATH=$PWD/$f python tools/sweep.py --workload cfg2 --views 1,148,4096 2>/dev/null | grep "^{" | python -c "
import sys, json
print('$v', [(json.loads(l)['views'], round(json.loads(l)['kernel_ms'], 3)) for l in sys.stdin])"
done
