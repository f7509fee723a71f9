"""Table of the scaling sweep (tools/sweep.sh): frame-iterations/s per (T_global, N) and the strong-scaling efficiency
value(N) / (N * value(1)) at equal T_global, read from gpurun_out/sweep/sweep_N*.jsonl."""
import glob
import json
import os
import sys

root = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep"
val = {}
for f in sorted(glob.glob(os.path.join(root, "sweep_N*.jsonl"))):
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        n = d["n_gpus"]
        tg = d["config"]["frames_per_gpu"] * n
        val[(tg, n)] = (d["value"], d["ms_per_step"])
ns = sorted({n for _, n in val})
tgs = sorted({t for t, _ in val})
print("| T_global | " + " | ".join("N=%d frame-iters/s (ms/step)" % n for n in ns) + " | " +
      " | ".join("strong eff N=%d" % n for n in ns if n > 1) + " |")
print("|---|" + "---|" * (len(ns) + len([n for n in ns if n > 1])))
for t in tgs:
    cells = ["%.4g (%.2f)" % val[(t, n)] if (t, n) in val else "-" for n in ns]
    eff = ["%.3f" % (val[(t, n)][0] / (n * val[(t, 1)][0])) if (t, n) in val and (t, 1) in val else "-" for n in ns if n > 1]
    print("| %d | " % t + " | ".join(cells + eff) + " |")
