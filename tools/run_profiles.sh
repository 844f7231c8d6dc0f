#!/bin/bash
# tools/run_profiles.sh [part] — the measurements behind profiles/r2_*.  Everything lands in gpurun_out/r2p_*;
# ncu reports are summarised on the spot (tools/ncu_summary.py) and only the summaries are kept, except
# for the report of the bench workload's chain.
#   part 1: bench + reference arm, stored-list / list-latency tools, ncu launch list of the bench command
#   part 2: ncu --set full of the kernel chain on cfg2 (VoxelType, Frontier, VoxelWeight), cfg1, cfg3
set -x
O=gpurun_out
if [ "${1:-1}" = 1 ]; then
  python bench.py --steps 40 --warmup 3 > $O/r2p_bench.json 2> $O/r2p_bench.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/r2p_bench_reference.json 2>/dev/null
  python tools/refold_bench.py > $O/r2p_refold.json 2> $O/r2p_refold.err
  python tools/list_latency.py > $O/r2p_list_latency.json 2> $O/r2p_list_latency.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
else
  NCU="ncu --set full --clock-control none --import-source on -k regex:k_split -s 10 -c 5 -f"
  summarise() {  # name: per-kernel metrics + SASS view of the leaf kernel (4th launch; 2nd for SimpleRayCaster)
    python tools/ncu_summary.py $O/$1.ncu-rep --kernel ${2:-3} --blocks > $O/$1_summary.txt 2>&1
  }
  $NCU -o $O/r2p_chain python tools/sweep.py --workload cfg2 --views 4096 --reps 2 > /dev/null 2>&1
  summarise r2p_chain
  python tools/ncu_summary.py $O/r2p_chain.ncu-rep --kernel 1 --blocks > $O/r2p_chain_mark_summary.txt 2>&1
  for e in Frontier VoxelWeight; do
    $NCU -o $O/r2p_$e python tools/sweep.py --workload cfg2 --views 4096 --reps 2 --evaluator $e > /dev/null 2>&1
    summarise r2p_$e; rm -f $O/r2p_$e.ncu-rep
  done
  $NCU -o $O/r2p_cfg1 python tools/sweep.py --workload cfg1 --views 4096 --reps 2 > /dev/null 2>&1
  summarise r2p_cfg1; rm -f $O/r2p_cfg1.ncu-rep
  $NCU -o $O/r2p_cfg3 python tools/sweep.py --workload cfg3 --views 1024 --reps 2 > /dev/null 2>&1
  summarise r2p_cfg3; rm -f $O/r2p_cfg3.ncu-rep
fi
ls -la $O/r2p_*
