#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each, after the same command ran without ncu):
#   gpurun --timeout 1800 -- 'bash tools/measure/profile_kernels.sh <name>'
# fit_tile_kernel on 1000 cfg2 frames (identity beta / a deformation per frame), on 100 cfg4 frames, and gram_tc_kernel on
# 40 cfg4 frames.  Raw and source pages are exported as CSV next to the reports; the reports themselves are deleted.
out=gpurun_out/${1:-prof}; mkdir -p $out; cd /root/repo
timeout 600 python tools/measure/fit_only.py cfg2,cfg4 0 5 > $out/fit_only.log 2>&1; echo plain $?
for tag in cfg2_identity cfg2_deformed cfg4_identity; do
  cfg=${tag%%_*}
  DNMF_PROFILE_RANGE=$tag timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:fit_tile -c 1 -o $out/fit_$tag python tools/measure/fit_only.py $cfg 0 2 > $out/ncu_$tag.log 2>&1; echo ncu $tag $?
  ncu -i $out/fit_$tag.ncu-rep --page raw --csv > $out/fit_${tag}_raw.csv 2>/dev/null
  ncu -i $out/fit_$tag.ncu-rep --page source --csv > $out/fit_${tag}_src.csv 2>/dev/null
  rm -f $out/fit_$tag.ncu-rep
done
timeout 300 python tools/measure/mu_only.py cfg4 > $out/mu_only.log 2>&1; echo plain mu $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc -c 1 -o $out/gram python tools/measure/mu_only.py cfg4 > $out/ncu_gram.log 2>&1; echo ncu gram $?
ncu -i $out/gram.ncu-rep --page raw --csv > $out/gram_raw.csv 2>/dev/null
ncu -i $out/gram.ncu-rep --page source --csv > $out/gram_src.csv 2>/dev/null
rm -f $out/gram.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg4_mu.csv python tools/measure/mu_only.py cfg4 > /dev/null 2>&1; echo launches mu $?
