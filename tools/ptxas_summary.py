"""Summarises volume-renderer_b200/lib/ptxas.log: registers, stack frame and spills per kernel instantiation."""
import collections
import re
import subprocess
import sys

log = open(sys.argv[1] if len(sys.argv) > 1 else "volume-renderer_b200/lib/ptxas.log").read()
ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log)
if not ents:
    print("no kernels found in the log (build error?)"); sys.exit(1)
dem = subprocess.run(["cu++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
rows = []
for (n, st, ss, sl, regs), d in zip(ents, dem):
    m = re.match(r"void vr::(\w+)<(.*)>\(", d)
    rows.append((m.group(1) if m else d[:40], m.group(2) if m else "", int(regs), int(st), int(ss), int(sl)))
print(len(rows), "kernels;", sum(1 for r in rows if r[3] > 0), "with a stack frame")
for r in rows:
    if r[3] > 0 or "-v" in sys.argv:
        print(f"  {r[0]}<{r[1]}>: {r[2]} regs, stack {r[3]} B, spill st/ld {r[4]}/{r[5]} B")
c = collections.Counter((r[0], r[2]) for r in rows)
for (k, regs), n in sorted(c.items()):
    print(f"{k}: {n} x {regs} regs")
