#!/usr/bin/env python
"""Print (or summarise) the SASS of one kernel of a library: tools/sass_fn.py LIB SUBSTRING [--count]"""
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
for part in parts[1:]:
    name = part.split("\n", 1)[0]
    if key in name:
        if "--count" in sys.argv:
            ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part, flags=re.M)
            from collections import Counter
            c = Counter(o.split(".")[0] for o in ops)
            print(name, len(ops), dict(c.most_common(25)))
        else:
            print(part)
