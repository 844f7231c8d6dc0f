for f in variant_*.so; do v=${f#variant_}; v=${v%.so}; A3D_LIB_PATH=$PWD/$f python tools/sweep.py --workload cfg2 --views 1,148,4096 2>/dev/null | grep "^{" | python -c "
import sys,json
print('$v', [ (json.loads(l)['views'], round(json.loads(l)['kernel_ms'],3)) for l in sys.stdin])"; done
