"""Loop structure of one kernel's SASS: every backward branch [target, branch] whose body holds F2I.FLOOR (the z march of
the fused kernel), with instruction counts by class.   python tools/measure/loops.py lib.so kernel_substring"""
import re
import subprocess
import sys
from collections import Counter

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for ln in part.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;?\s*/\*", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr_ix = {a: i for i, (a, _) in enumerate(ins)}
    print(name, len(ins), "instructions")
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?|BRA.*0x([0-9a-f]+)", t)
        m2 = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= a and tgt in addr_ix:
                loops.append((addr_ix[tgt], i))
    for a, b in loops:
        body = [t for _, t in ins[a:b + 1]]
        nf2i = sum("F2I" in t for t in body)
        if nf2i >= 3 and b - a < 700:
            c = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
            print("  loop %5d-%5d (%3d instr)  %s" % (a, b, b - a + 1, " ".join("%s=%d" % kv for kv in c.most_common(14))))
