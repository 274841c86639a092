"""Per-kernel summary of an `ncu --set full` report (CPU only): duration, DRAM bytes, occupancy, stall ratios; with a kernel
regex also the hottest SASS instructions of that kernel.   python tools/ncu_summary.py REPORT [KERNEL_REGEX [N]]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rdMB", "dram__bytes_write.sum": "wrMB",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bankc",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "long", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "short",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "bar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "membar",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "wait", "smsp__inst_executed.sum": "inst"}
idx = {v: hdr.index(k) for k, v in want.items() if k in hdr}
ki = hdr.index("Kernel Name")
seen = set()
for r in rows[2:]:
    name = r[ki].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")[:38]
    if name in seen:
        continue
    seen.add(name)
    def f(k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except Exception:
            return float("nan")
    print(f"{name:38s} {f('us'):7.1f}us  rd {f('rdMB'):7.2f} wr {f('wrMB'):6.2f} MB  sm {f('sm%'):4.0f}% dram {f('dram%'):4.0f}%  regs {f('regs'):4.0f} grid {f('grid'):5.0f} "
          f"warps {f('warps%'):4.0f}%  stall long {f('long'):5.1f} short {f('short'):4.1f} bar {f('bar'):4.1f} membar {f('membar'):4.1f} wait {f('wait'):4.1f}  inst {f('inst'):9.0f} bankc {f('bankc'):8.0f}")
if len(sys.argv) > 2:
    kern, n = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for bi, b in enumerate(blocks[:1]):
        end = blocks[bi + 1] if bi + 1 < len(blocks) else len(rows)
        h = rows[b + 1]
        si, ie = h.index("# Samples"), h.index("Instructions Executed")
        data = [r for r in rows[b + 2:end] if len(r) > ie and r[si].isdigit()]
        tot = sum(int(r[si]) for r in data)
        print(rows[b][1][:80], "samples", tot)
        for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
            print(f"  {100 * int(r[si]) / max(tot, 1):5.1f}%  exec {r[ie]:>7}  {r[1].strip()[:110]}")
        ops = collections.Counter()
        for r in data:
            t = r[1].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += int(r[si])
        print("  by opcode:", [(k, round(100 * v / max(tot, 1))) for k, v in ops.most_common(8)])
