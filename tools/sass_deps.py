#!/usr/bin/env python
"""Static dependency distances in the hot loop of a kernel (from cuobjdump -sass): for every FP64
instruction, how many instructions (and how many FP64 instructions) earlier its closest FP64
producer was issued.  A DFMA result is ready 8.5 cycles (~4 FP64 issue slots) after issue, so
distances below 4 FP64 slots mean the warp stalls there.

    python tools/sass_deps.py victor_b200/libvictor_b200.so 'K1CfgILb1ELb0ELi4ELi5E'
"""
import collections
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from sass_mix import functions  # noqa: E402

FP64 = ("DFMA", "DMUL", "DADD")


def regs_of(text):
    """(dest registers, source registers) of an FP64 arithmetic instruction (64-bit pairs)."""
    parts = text.split(None, 1)
    if parts[0].startswith("@"):
        parts = parts[1].split(None, 1)
    ops = [o.strip() for o in parts[1].rstrip(";").split(",")]
    def pair(o):
        m = re.match(r"-?\|?R(\d+)", o)
        if not m:
            return set()
        n = int(m.group(1))
        return {n, n + 1}
    dst = pair(ops[0])
    src = set()
    for o in ops[1:]:
        src |= pair(o)
    return dst, src


def main():
    lib, pat = sys.argv[1], sys.argv[2]
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
        best = max(inner, key=lambda ab: sum(1 for _, t in ins[ab[0]:ab[1] + 1] if any(f in t for f in FP64)))
        body = [t for _, t in ins[best[0]:best[1] + 1]]
        body2 = body + body          # wrap around once for loop-carried dependencies
        last_writer = {}              # reg -> (index in body2, fp64 count at that point)
        fp = 0
        hist = collections.Counter()
        short = []
        for i, t in enumerate(body2):
            op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
            if op in FP64:
                dst, src = regs_of(t)
                prod = [last_writer[r] for r in src if r in last_writer]
                if prod and i >= len(body):
                    pi, pf = max(prod)
                    d = fp - pf
                    hist[min(d, 8)] += 1
                    if d < 4:
                        short.append((d, i - pi, t[:60]))
                for r in dst:
                    last_writer[r] = (i, fp)
                fp += 1
            else:
                # any other instruction overwriting a register ends the FP64 dependence through it
                m = re.match(r"(?:@!?U?P\d+\s+)?\S+\s+R(\d+)", t)
                if m:
                    last_writer.pop(int(m.group(1)), None)
        n = sum(hist.values())
        print("==", name)
        print("  FP64 instr in loop:", n, " distance to closest FP64 producer (in FP64 slots): ",
              {k: v for k, v in sorted(hist.items())})
        print("  stalling (< 4 slots):", sum(v for k, v in hist.items() if k < 4), "of", n)
        for d, di, t in short[:12]:
            print(f"    d={d} ({di} instr)  {t}")


if __name__ == "__main__":
    main()
