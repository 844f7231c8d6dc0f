#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics of every kernel in the report, then
the SASS opcode mix, stall mix and (--blocks) hot blocks of ONE kernel (--kernel K: the K-th launch in
the report, default 0).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--blocks] [--kernel K]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_bytes.sum ",
        "l1tex__t_bytes.sum ", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum ", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg ",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ", "lts__t_sectors_op_read.sum ",
        "sm__inst_executed_pipe_xu_realtime", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        print("== kernel:", row[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, row):
            if any(h.startswith(k.strip()) and (not k.endswith(" ") or h == k.strip()) for k in KEYS):
                print(f"  {h:75s} {v:>18s} {u}")
    k = int(sys.argv[sys.argv.index("--kernel") + 1]) if "--kernel" in sys.argv else 0
    src = page(rep, "source", ("--launch-skip", str(k), "--launch-count", "1"))
    print("\n== SASS of launch", k, ":", src[0][1][:110] if len(src[0]) > 1 else "")
    hdr = src[1]
    # (the page repeats the kernel once per view; keep the first copy)
    seen_addr = set()
    body = []
    for r in src[2:]:
        if len(r) >= len(hdr) and r[0].startswith("0x"):
            if r[0] in seen_addr:
                break
            seen_addr.add(r[0])
        body.append(r)
    src = src[:2] + body
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samples = collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = tots = thr = 0
    data = []
    for r in src[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[ix["Instructions Executed"]])
            s = int(r[ix["# Samples"]])
            t = int(r[ix["Thread Instructions Executed"]])
        except ValueError:
            continue
        sass = r[ix["Source"]]
        toks = sass.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = ".".join(op.split(".")[:3])
        ops[op] += n
        samples[op] += s
        tot += n
        tots += s
        thr += t
        data.append((sass, n, s, t))
        for c in stall_cols:
            try:
                stalls[c] += int(r[ix[c]])
            except ValueError:
                pass
    print(f"\nwarp instructions {tot:,}  thread instr {thr:,}  avg active lanes {thr / max(tot, 1):.1f}  samples {tots:,}")
    for op, n in ops.most_common(40):
        print(f"  {op:28s} {n:14,d} {100 * n / tot:5.1f}%   stall-samples {100 * samples[op] / max(tots, 1):5.1f}%")
    print("stalls:", ", ".join(f"{k}={100 * v / max(tots, 1):.1f}%" for k, v in stalls.most_common(10)))
    if "--blocks" in sys.argv:
        prev, start = None, 0
        for i, (sass, n, s, t) in enumerate(data + [("", -1, 0, 0)]):
            if prev is None or abs(n - prev) > 0.02 * max(n, prev, 1):
                if prev is not None and prev > 0:
                    blk = data[start:i]
                    bt = sum(d[3] for d in blk) / max(sum(d[1] for d in blk), 1)
                    print(f"  [{start:5d}-{i - 1:5d}] n={i - start:4d} exec={prev / 1e6:9.2f}M lanes={bt:5.1f} "
                          f"samples={sum(d[2] for d in blk):8d}  {blk[0][0][:60]}")
                start = i
            prev = n


if __name__ == "__main__":
    main()
