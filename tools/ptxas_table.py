"""Print (kernel, registers, spill bytes, smem) from pm-rl_b200/build/ptxas.log."""
import re, subprocess, sys, os
log = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pm-rl_b200", "build", "ptxas.log")).read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)")
names = [m.group(1) for m in pat.finditer(log)]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
for m, d in zip(pat.finditer(log), dem):
    d = d.replace("pmrl::", "").replace("(StepParams)", "").replace("void ", "")
    print(f"{d[:70]:70s} regs={m.group(5):>3s} spill={m.group(3):>4s}/{m.group(4):<4s} {m.group(6).strip()[:40]}")
