#!/bin/bash
# one-off measurement script (round 1, session 3): local sweeps kernel (thread per row), per-kernel times of the trace update
mkdir -p gpurun_out/s39; cd /root/repo
timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -k "mu_ or update_footprints or demo_trajectory or shuffled" > gpurun_out/s39/pytest.log 2>&1; echo pytest exit $?; tail -15 gpurun_out/s39/pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s39/bench_cfg2.json 2> gpurun_out/s39/bench_cfg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mu_|fit_tile" -c 3000 --csv --log-file gpurun_out/s39/launches_mu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s39/ncu_list.log 2>&1; echo list $?
python - <<'PY'
import json,glob,csv,collections
for f in sorted(glob.glob("gpurun_out/s39/bench_*.json")):
    try:
        d=json.load(open(f)); mu=d.get("trace_update") or {}; print(f, d["value"], d["roofline"]["frac"], mu.get("stats_ms"), mu.get("sweeps_ms"), d["reference_batch"]["value"])
    except Exception as e: print(f, "ERR", e)
rows=list(csv.reader(open('gpurun_out/s39/launches_mu.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]
    if 'mu_' in k or '3, 1>' in k:
        a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=v; a[2]=max(a[2],v)
for k,a in agg.items(): print(k, a)
PY
