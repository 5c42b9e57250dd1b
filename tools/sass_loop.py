"""Dumps the SASS of one kernel instantiation of libvolren_b200.so (by demangled-substring match) and reports the
instruction mix of its hot loop = the code between the first backward branch target and that branch, widest loop
containing a TLD4/TLD.  Usage: python tools/sass_loop.py '<substring of demangled name>' [--dump]"""
import collections
import re
import subprocess
import sys

import os
so = os.environ.get("SO", "volume-renderer_b200/lib/libvolren_b200.so")
pat = sys.argv[1]
names = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
syms = sorted(set(re.findall(r"\.text\.(_Z\w+)", names)))
dem = subprocess.run(["cu++filt"] + syms, capture_output=True, text=True).stdout.splitlines()
cands = [(s, d) for s, d in zip(syms, dem) if pat in d.replace("(int)", "").replace("(bool)", "").replace(" ", "")]
if len(cands) != 1:
    print(len(cands), "matches:")
    for s, d in cands[:40]:
        print("  ", d.replace("(int)", "").replace("(bool)", "")[:160])
    sys.exit(1)
sym, d = cands[0]
print(d[:200])
sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, so], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            body = ins[addr_index[tgt]:i + 1]
            if any(("TLD" in x[1] or "TEX" in x[1]) for x in body):
                loops.append(body)
print("total instructions", len(ins), "; loops with texture fetches:", [len(b) for b in loops])
if loops:
    body = max(loops, key=len)
    ops = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    ntex = sum(v for k, v in ops.items() if k in ("TLD4", "TLD", "TEX"))
    print(f"hot loop: {len(body)} instructions, {ntex} texture fetches -> {len(body) / max(ntex, 1):.1f} instr/sample")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
    print("  local-memory ops in loop:", sum(v for k, v in ops.items() if k in ("LDL", "STL")))
if "--dump" in sys.argv:
    for a, t in ins:
        print(f"{a:05x}  {t}")
