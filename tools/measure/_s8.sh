#!/bin/bash
cd /root/repo
run() { # label, env...
  lbl=$1; shift
  env "$@" timeout 600 python tools/measure/fit_only.py cfg2,cfg3,cfg4 0 10 2>&1 | grep -v Warn | sed "s/^/$lbl: /" | cut -c1-150
  env "$@" timeout 600 python tools/measure/fit_only.py cfg2 125 20 2>&1 | grep -v Warn | sed "s/^/$lbl: /" | cut -c1-150
}
run head DNMF_B200_LIB=$PWD/variants/head/libdnmf_b200.so
run tailoff DNMF_FPC_TAIL_OFF=1
run tail DNMF_X=1
