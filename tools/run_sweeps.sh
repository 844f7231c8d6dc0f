set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 600 gpurun_out/r2_bench_b.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_b.json 2>/dev/null
: > gpurun_out/r2_sweep.jsonl
python tools/sweep.py --workload cfg2 --views 1,12,148,1024,4096,16384,65536 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl
python tools/sweep.py --workload cfg2 --views 1048576 --reps 1 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl
python tools/sweep.py --workload cfg1 --views 64,4096 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl
python tools/sweep.py --workload cfg4 --views 64,1024 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl
for e in Frontier VoxelWeight Naive; do python tools/sweep.py --workload cfg2 --views 4096 --evaluator $e 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl; done
python tools/sweep.py --workload cfg3 --views 64,1024,4096 2>/dev/null | grep "^{" >> gpurun_out/r2_sweep.jsonl
python - <<PY
import json
for l in open("gpurun_out/r2_sweep.jsonl"):
    d=json.loads(l); print(d["workload"],d["evaluator"],d["views"],round(d["kernel_ms"],3),round(d["frac_of_hbm"],3),d["kernel"],d.get("stage_ms"))
d=json.load(open("gpurun_out/r2_bench_b.json")); print(d["value"], d["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"]["kernel_ms_by_stage"], d.get("cpu_baseline"))
PY
