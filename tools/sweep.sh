#!/bin/bash
# Strong / weak scaling sweep of BASELINE configuration 5: T_global = 1k .. 40k frames of the cfg2 volume on N GPUs
# of one node (N = number of visible GPUs unless given).  One JSON line per (N, T_global) into
# gpurun_out/sweep/sweep_N<N>.jsonl; tools/sweep_table.py turns the files of several N into the table under profiles/.
#   gpurun --gpus 8 -- 'bash tools/sweep.sh 8'        gpurun -- 'bash tools/sweep.sh 1'
set -u
cd "$(dirname "$0")/.."
N=${1:-$(nvidia-smi -L | wc -l)}
TG_LIST=${TG_LIST:-"1000 2000 5000 10000 20000 40000"}
mkdir -p gpurun_out/sweep
out=gpurun_out/sweep/sweep_N${N}.jsonl
: > "$out"
for TG in $TG_LIST; do
  F=$(( TG / N ))
  if [ $(( F * N )) -ne "$TG" ] || [ "$F" -lt 1 ]; then continue; fi
  if [ "$N" -eq 1 ]; then
    timeout 1200 python bench.py --gpus 1 --frames "$F" --steps 10 --warmup 3 --no-legs --no-cpu-baseline --no-e2e --no-mu >> "$out" 2> gpurun_out/sweep/err_N${N}_T${TG}.log
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $(( 29600 + N )) \
      bench.py --gpus "$N" --frames "$F" --steps 10 --warmup 3 --no-legs --no-cpu-baseline --no-e2e --no-mu >> "$out" 2> gpurun_out/sweep/err_N${N}_T${TG}.log
  fi
  echo "N=$N T_global=$TG frames/GPU=$F exit $?"
done
python tools/sweep_table.py gpurun_out/sweep
