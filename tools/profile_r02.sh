#!/bin/bash
# ncu evidence of round 2 (run on a GPU box through gpurun; one GPU).  For every target: one plain run (must exit 0),
# then `ncu --set full` of 3 launches of the target kernel; the raw page goes to gpurun_out/r02_ncu_full_<target>_raw.csv.
# Finally the launch list of a short bench.py run (device time per launch: compare SHARES with the bench, not absolutes).
set -u
mkdir -p gpurun_out
targets=${1:-"cfg2:spn_fused_net cfg5:spn_fused_net cfg5r:spn_fused_net cfg3p:spn_fused_net cfg3c:spn_fused_net pair1:spn_pair_contract pair3:spn_pair_contract zgemm:contract_zgemm dgemm:contract_dgemm mixed:contract_dgemm exchange:mc_sum_exchange"}
for tk in $targets; do
  t=${tk%%:*}; k=${tk##*:}
  python tools/prof_driver.py $t > gpurun_out/plain_$t.log 2>&1 || { echo "plain run of $t failed"; tail -3 gpurun_out/plain_$t.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 3 -f -o gpurun_out/r02_$t python tools/prof_driver.py $t > gpurun_out/ncu_$t.log 2>&1
  if [ -f gpurun_out/r02_$t.ncu-rep ]; then
    ncu -i gpurun_out/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${t}_raw.csv 2>/dev/null
    case $t in cfg2|cfg5|pair1) ;; *) rm -f gpurun_out/r02_$t.ncu-rep;; esac     # keep three reports for the source page
  else
    echo "no report for $t"; tail -5 gpurun_out/ncu_$t.log
  fi
done
python bench.py --suite headline --steps 30 --warmup 3 --e2e-steps 2 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg2.csv \
    python bench.py --suite headline --steps 30 --warmup 3 --e2e-steps 2 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/ncu_bench.log 2>&1
echo profiled
