#!/usr/bin/env python
"""Register-file read traffic of a kernel's hot loop (from cuobjdump -sass): 32-bit register source
operands per velocity node, counting 64-bit operands of FP64 / 64-bit instructions twice and
operands served by the reuse cache (same register, same slot, flagged .reuse on the previous
instruction of the same kind) as free.

    python tools/sass_regreads.py victor_b200/libvictor_b200.so 'K1CfgILb1ELb0ELi4ELi5ELi3ELb0E' 4
"""
import collections
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from sass_mix import functions  # noqa: E402

WIDE = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RSQ64H", "MUFU.RCP64H", "F2I.F64", "I2F.F64")


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    for name, ins in functions(lib).items():
        if pat not in name:
            continue
        loops = []
        for idx, (addr, text) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                tgt = int(m.group(1), 16)
                loops.append((next(i for i, (a, _) in enumerate(ins) if a >= tgt), idx))
        inner = [(a, b) for a, b in loops if not any((c, d) != (a, b) and a <= c and d <= b for c, d in loops)]
        best = max(inner, key=lambda ab: sum(1 for _, t in ins[ab[0]:ab[1] + 1] if "DFMA" in t))
        reads = collections.Counter()
        prev_reuse = {}
        for _, t in ins[best[0]:best[1] + 1]:
            parts = t.split(None, 1)
            if parts[0].startswith("@"):
                parts = parts[1].split(None, 1)
            op = parts[0]
            base = op.split(".")[0]
            ops = [o.strip() for o in parts[1].rstrip(";").split(",")] if len(parts) > 1 else []
            is_store = base in ("STS", "STG", "ST")
            srcs = ops if is_store else ops[1:]
            width = 2 if any(op.startswith(w) for w in WIDE) else 1
            if base == "MUFU":
                width = 1          # the 64H variants read the high word only
            now_reuse = {}
            for slot, o in enumerate(srcs):
                regs = re.findall(r"R(\d+)(\.reuse)?", o)
                for reg, ru in regs:
                    if reg == "Z":
                        continue
                    if ru:
                        now_reuse[slot] = reg
                    if prev_reuse.get(slot) == reg:
                        continue
                    w = width
                    if base in ("LDS", "LDG", "LDC", "STS") or "[" in o:
                        w = 1
                    if base == "STS" and slot == 1:
                        w = 2 if ".64" in op else (4 if ".128" in op else 1)
                    reads[base] += w
            prev_reuse = now_reuse
        tot = sum(reads.values())
        print(f"== {name}\n  32-bit register reads per node: {tot / per:.1f}   by opcode: "
              f"{ {k: round(v / per, 1) for k, v in reads.most_common()} }")


if __name__ == "__main__":
    main()
