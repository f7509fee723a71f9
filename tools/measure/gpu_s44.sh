#!/bin/bash
# one-off: instruction-cache footprint variants of the fit kernel (run-time tail kind, always-clamped loop)
mkdir -p gpurun_out/s44; cd /root/repo
for v in base dyn dyn_safe; do
  if [ $v = base ]; then lib=/root/repo/dnmf_b200/_C/libdnmf_b200.so; else lib=/root/repo/variants/lib_$v.so; fi
  DNMF_B200_LIB=$lib timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s44/bench_cfg2_$v.json 2> gpurun_out/s44/bench_cfg2_$v.err
done
DNMF_B200_LIB=/root/repo/variants/lib_dyn.so timeout 200 python bench.py --config cfg3 --frames 300 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s44/bench_cfg3_dyn.json 2> gpurun_out/s44/bench_cfg3_dyn.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s44/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["deformed_beta"]["kernel_ms_per_launch"], d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
