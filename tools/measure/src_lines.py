"""Map the per-instruction samples of an `ncu --page source --csv` export to source lines through
`nvdisasm --print-line-info` of the same kernel (instruction order is identical), and print samples per source line.
   python tools/measure/src_lines.py file_src.csv cubin mangled_kernel_name [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
txt = subprocess.run(["nvdisasm", "--print-line-info", sys.argv[2]], capture_output=True, text=True).stdout
sec = txt.split(".text." + sys.argv[3] + ":", 1)[1]
sec = sec.split("//--------------------- .", 1)[0]
line, lines = None, []
for ln in sec.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(line)
assert len(lines) == len(data), (len(lines), len(data))
agg, exa = defaultdict(int), defaultdict(int)
tot = 0
for r, l in zip(data, lines):
    s = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    agg[l] += s
    exa[l] += int(r[ix["Instructions Executed"]] or 0)
    tot += s
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
byfile = sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0)))
print("total samples", tot)
for l, s in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print("%-22s %6d %5.2f%%  exec %.3g" % ("%s:%d" % l if l else "?", s, 100.0 * s / tot, exa[l]))
# coarse: cumulative by line ranges of dnmf_fit.cuh
print("--- by line (dnmf_fit.cuh), cumulative in source order")
acc = 0
for l, s in byfile:
    if l and l[0] == "dnmf_fit.cuh" and s > 0.002 * tot:
        print("%5d %6.2f%%" % (l[1], 100.0 * s / tot))
# regions given as a:b index ranges -> the dnmf_fit.cuh lines they come from
for arg in sys.argv[5:]:
    a, b = (int(x) for x in arg.split(":"))
    c = defaultdict(int)
    last = None
    for i in range(a, b):
        l = lines[i]
        if l and l[0] != "sm_100_rt.hpp":
            last = l
        c[last] += int(data[i][ix["Warp Stall Sampling (All Samples)"]] or 0)
    s = sum(c.values())
    print("region %d:%d samples %.2f%%: " % (a, b, 100.0 * s / tot) + ", ".join("%s:%d=%.1f%%" % (k[0][:8], k[1], 100.0 * v / tot) for k, v in sorted(c.items(), key=lambda kv: -kv[1])[:8] if k))
