#!/bin/bash
# tools/run_scaling.sh N — bench.py under torchrun on N GPUs of this box (weak: 4096 views per GPU;
# strong: 4096 views in total) + the single-process group (a3d_group_*) on 1..N GPUs.  Results under
# gpurun_out/r2_scale_*.json(l).
N=$1
for mode in weak strong; do
  extra=""; [ $mode = strong ] && extra="--global-views 4096"
  if [ "$N" = 1 ]; then
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline $extra > gpurun_out/r2_scale_${mode}_n1.json 2> gpurun_out/r2_scale_${mode}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 $extra > gpurun_out/r2_scale_${mode}_n$N.json 2> gpurun_out/r2_scale_${mode}_n$N.err
  fi
  python - <<PY
import json
for l in open("gpurun_out/r2_scale_${mode}_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$mode", d["n_gpus"], round(d["ms_per_step"],3), round(d["kernel_ms"],3), round(d["value"]), round(d["e2e"]["value"]), d["gain_digest"])
PY
done
