#!/usr/bin/env python3
"""Static instruction count of the innermost hot loop(s) of a kernel: dumps SASS of functions matching a
substring from an object file and reports, for every backward branch, the span length and opcode mix."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name[:100], len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, txt) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", txt)
        if m:
            t = int(m.group(1), 16)
            if t < a and t in addr and i - addr[t] >= minlen:
                body = ins[addr[t]:i + 1]
                mix = collections.Counter()
                for _, x in body:
                    op = x.split()[1] if x.startswith("@") else x.split()[0]
                    mix[op.split(".")[0]] += 1
                print(f"  loop {t:#x}..{a:#x}: {len(body)} instructions;", ", ".join(f"{k} {v}" for k, v in mix.most_common(16)))
