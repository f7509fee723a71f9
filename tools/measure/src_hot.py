"""Summarise an `ncu --page source --csv` export: executed-instruction weighted regions and per-instruction stall
samples of the hottest code.   python tools/measure/src_hot.py file.csv [min_exec_fraction] [--dump]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
ex = [int(r[ix["Instructions Executed"]] or 0) for r in data]
smp = [int(r[ix["Warp Stall Sampling (All Samples)"]] or 0) for r in data]
tot_ex, tot_s = sum(ex), sum(smp)
print("instructions %d  executed %.4g  samples %d" % (len(data), tot_ex, tot_s))
thr = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 0.0005
# regions: maximal runs where executed >= thr * total
regions, cur = [], None
for i, e in enumerate(ex):
    hot = e >= thr * tot_ex
    if hot and cur is None:
        cur = i
    if not hot and cur is not None:
        regions.append((cur, i))
        cur = None
if cur is not None:
    regions.append((cur, len(ex)))
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for a, b in regions:
    e, s = sum(ex[a:b]), sum(smp[a:b])
    if s < 0.005 * tot_s:
        continue
    st = {n: sum(int(r[ix[n]] or 0) for r in data[a:b]) for n in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print("region %5d-%5d (%4d instr) exec %5.1f%%  samples %5.1f%%  ipc-ish %.2f  %s" %
          (a, b, b - a, 100 * e / tot_ex, 100 * s / tot_s, e / max(s, 1), " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(s, 1)) for k, v in top)))
if "--dump" in sys.argv:
    a0, b0 = (int(x) for x in sys.argv[sys.argv.index("--dump") + 1].split("-"))
    for i in range(a0, b0):
        r = data[i]
        st = {n: int(r[ix[n]] or 0) for n in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print("%5d %-70s ex %9d smp %5d  %s" % (i, r[1].strip()[:70], ex[i], smp[i], " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))
