#!/bin/bash
# round-end check of the final state: GPU suite, default bench, reference arm, smoke, launch list, ncu --set full of the fit kernel
mkdir -p gpurun_out/final; cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest.log 2>&1; echo pytest exit $?; tail -3 gpurun_out/final/pytest.log
timeout 400 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo bench $?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo ref $?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 3 --warmup 3 --frames 200 --no-cpu-baseline > gpurun_out/final/ncu_list.log 2>&1; echo list $?
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fit_tile_kernelILi1ELi1ELi2ELi0E --launch-skip 5 -c 1 -o gpurun_out/final/prof_fit_v13 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-mu > gpurun_out/final/ncu_fit.log 2>&1; echo full $?
cut -c1-300 gpurun_out/final/bench.json
