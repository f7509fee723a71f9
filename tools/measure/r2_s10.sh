#!/bin/bash
# round 2, session 10: new bench.py (all legs), reference arm, smoke, whole GPU suite
mkdir -p gpurun_out/r2s10; cd /root/repo
timeout 900 python bench.py > gpurun_out/r2s10/bench.json 2> gpurun_out/r2s10/bench.err; echo bench $?; tail -3 gpurun_out/r2s10/bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2s10/bench.json").read().strip().splitlines()[-1])
print("value %.4g frac %.3f ms/step %.3f"%(d["value"], d["roofline"]["frac"], d["ms_per_step"]))
print("e2e", {k:d["e2e"][k] for k in ("value","h2d_gbps_per_gpu","h2d_link_ceiling_gbps_per_gpu","host")})
print("e2e_b4", d["e2e_reference_batch"]); print("e2e_res", {k:d["e2e_resident"][k] for k in ("value","steps")})
print("deformed", d["deformed_beta"]["value"], d["deformed_beta"]["frac"]); print("ref_batch", d["reference_batch"]["value"])
print("mu", {k:d["trace_update"][k] for k in ("value","stats_ms","sweeps_ms","stats_kernel")})
for n,l in (d["configs"] or {}).items():
    print(n, {k:l.get(k) for k in ("value","ms_per_step","frames_per_gpu")}, "frac", l.get("roofline",{}).get("frac"), "deformed frac", l.get("deformed_beta",{}).get("frac"), "mu", {k:l.get("trace_update",{}).get(k) for k in ("stats_ms","sweeps_ms","stats_kernel")}, l.get("error"))
print("cpu", d["cpu_baseline"])
P
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2s10/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/r2s10/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
