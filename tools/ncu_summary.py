"""Condenses an .ncu-rep (one kernel launch, --set full --import-source on) into a text summary:
key raw metrics, warp-stall breakdown, opcode mix and the hottest SASS instructions.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print("# ncu summary of", rep)
print("kernel:", M.get("Kernel Name", ("?",))[0], " grid:", M.get("Grid Size", ("?",))[0], " block:", M.get("Block Size", ("?",))[0])
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second"]
print("\n## key metrics")
for k in keys:
    if k in M:
        print(f"{k:70s} {M[k][0]:>20s} {M[k][1]}")
print("\n## warp stall reasons (average warps stalled per issue-active cycle)")
st = [(float(v[0]), h) for h, v in M.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[0]]
for v, h in sorted(st, reverse=True):
    print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), int(r[isamp] or 0), int(r[iex] or 0), r[isrc]))
    except Exception:
        pass
base = min(d[0] for d in data)
tot, totex = sum(d[1] for d in data), sum(d[2] for d in data)
op, ops = collections.Counter(), collections.Counter()
for a, s, ex, sc in data:
    o = sc.split()[0] if not sc.startswith("@") else sc.split()[1]
    o = o.split(".")[0]
    op[o] += ex
    ops[o] += s
print(f"\n## opcode mix ({totex} warp instructions, {tot} stall samples)")
for o, c in op.most_common(16):
    print(f"{o:10s} executed {100 * c / totex:5.1f}%   samples {100 * ops[o] / tot:5.1f}%")
print("\n## hottest SASS instructions by stall samples")
for a, s, ex, sc in sorted(data, key=lambda d: -d[1])[:25]:
    print(f"+0x{a - base:05x} {100 * s / tot:5.1f}%  executed {ex:12d}  {sc[:90]}")
