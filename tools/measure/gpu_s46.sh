#!/bin/bash
# one-off: specialised main loops with even / odd lists merged (4 bodies instead of 6)
mkdir -p gpurun_out/s46; cd /root/repo
for v in base m01; do
  if [ $v = base ]; then lib=/root/repo/dnmf_b200/_C/libdnmf_b200.so; else lib=/root/repo/variants/lib_$v.so; fi
  DNMF_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s46/bench_cfg2_$v.json 2> gpurun_out/s46/bench_cfg2_$v.err
  DNMF_DYN_TAIL=0 DNMF_B200_LIB=$lib timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/s46/bench_cfg2_${v}_static.json 2> gpurun_out/s46/bench_cfg2_${v}_static.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s46/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, d["value"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["deformed_beta"]["kernel_ms_per_launch"], d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
