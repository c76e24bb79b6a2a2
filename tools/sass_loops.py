"""Static look at a kernel's SASS: list loops (backward branches) with their opcode histograms.
usage: python tools/sass_loops.py <lib.so> <mangled-name-substring>"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.check_output(["cuobjdump", "-sass", lib]).decode()
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for ln in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]:k + 1]
                ops = collections.Counter()
                for _, t in body:
                    t = re.sub(r"^@!?U?P\d+\s+", "", t)
                    ops[t.split()[0].split(".")[0]] += 1
                fp64 = sum(v for o, v in ops.items() if o in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"))
                print("  loop %#x..%#x: %d instr, %d fp64/mufu  %s" % (tgt, a, len(body), fp64,
                                                                   dict(ops.most_common(14))))
