#!/bin/bash
# The 1-GPU record of a round: GPU test suite, smoke, the default bench line, the reference arm and the ncu launch
# list of the same bench command.       gpurun --timeout 2400 -- 'bash tools/measure/record_1gpu.sh <name>'
# Outputs land in gpurun_out/<name>/; copy what is to be judged into profiles/.
out=gpurun_out/${1:-record}; mkdir -p $out; cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err ) 2>&1 | grep real; echo bench $?; tail -2 $out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo ref $?
if [ "${NCU:-1}" = 1 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-legs --no-cpu-baseline > $out/ncu_bench.log 2>&1; echo ncu $?
fi
