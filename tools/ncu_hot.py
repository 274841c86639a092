"""Top stall-sampled SASS instructions of one kernel in an .ncu-rep (needs ncu on PATH; CPU only)."""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) > ie and r[si].isdigit():
        if r[0] in seen:
            break          # second launch of the same kernel
        seen.add(r[0]); data.append(r)
tot = sum(int(r[si]) for r in data)
print(f"{kern}: {len(data)} SASS instr, {tot} samples, {sum(int(r[ie]) for r in data)} warp-instr executed")
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    print(f"{100*int(r[si])/max(tot,1):5.1f}%  exec {r[ie]:>7}  {r[1].strip()[:90]}")
ops = collections.Counter()
for r in data:
    t = r[1].split(); op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[si])
print("by opcode:", [(k, round(100*v/max(tot,1))) for k, v in ops.most_common(8)])
