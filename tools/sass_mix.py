#!/usr/bin/env python
"""Static instruction mix of the innermost hot loops of a kernel, from `cuobjdump -sass`.

    python tools/sass_mix.py victor_b200/libvictor_b200.so 'K1CfgILb1ELb0ELi4ELi6ELb0' [nodes_per_trip]

Finds every backward branch in the matching function, takes the loop body it closes and prints
the opcode histogram of the bodies holding the most FP64 instructions (per velocity node when
`nodes_per_trip` is given).  No GPU needed.
"""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, body = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if cur and m:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        print("==", name, len(ins), "instructions")
        loops = []
        for idx, (addr, text) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                tgt = int(m.group(1), 16)
                start = next(i for i, (a, _) in enumerate(ins) if a >= tgt)
                loops.append((start, idx))
        scored = []
        inner = [(a, b) for a, b in loops if not any((c, d) != (a, b) and a <= c and d <= b for c, d in loops)]
        for start, end in inner:
            ops = collections.Counter()
            for _, text in ins[start:end + 1]:
                parts = text.split()
                op = parts[1] if parts[0].startswith("@") else parts[0]
                ops[op.split(".")[0]] += 1
            f64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            scored.append((f64, start, end, ops))
        for f64, start, end, ops in sorted(scored, key=lambda t: -t[0])[:2]:
            tot = sum(ops.values())
            print(f"  loop @{ins[start][0]:#x}..{ins[end][0]:#x}: {tot} instr, {f64} fp64  "
                  f"-> per node: total {tot / per:.1f}, fp64 {f64 / per:.1f}")
            print("   ", {k: round(v / per, 2) for k, v in ops.most_common()})


if __name__ == "__main__":
    main()
