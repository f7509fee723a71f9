#!/usr/bin/env python
"""Print the SASS of one kernel of the built library (substring match on the mangled name)."""
import re
import subprocess
import sys

so = "dnmf_b200/_C/libdnmf_b200.so"
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
for part in parts[1:]:
    name = part.split("\n", 1)[0].strip()
    if pat in name:
        lines = []
        for ln in part.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", ln)
            if m:
                lines.append("%s  %s" % (m.group(1), m.group(2)))
        print(name)
        print("\n".join(lines))
        break
