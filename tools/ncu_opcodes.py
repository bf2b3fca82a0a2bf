"""ncu report → dynamic opcode histogram (warp-instructions executed per opcode) + per-source-line totals.
usage: ncu_opcodes.py <rep> [out.txt]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None; data = []
for x in rows:
    if x and x[0] == "Address": h = x; continue
    if h and len(x) == len(h): data.append(dict(zip(h, x)))
ops = collections.Counter(); tot = 0
for d in data:
    n = int(d.get("Instructions Executed") or 0)
    s = d["Source"].strip()
    toks = s.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    ops[op] += n; tot += n
out = [f"total warp-instructions {tot}"]
for op, n in ops.most_common(45):
    out.append(f"{n:12d} {100*n/tot:5.1f}%  {op}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        fh.write(txt + "\n\n")
        for d in data:
            fh.write(f'{int(d.get("Instructions Executed") or 0):10d} {int(d.get("# Samples") or 0):6d}  {d["Source"]}\n')
