#!/bin/bash
# Sweep the fused-kernel launch knobs (block size, min blocks/SM, prefetch mode) for one workload on the GPU box.
# usage: tools/knob_sweep.sh cfg5 > gpurun_out/sweep_cfg5.txt
wl=${1:-cfg5}
export SPN_JIT_CACHE=/tmp/spn_sweep_cache
mkdir -p $SPN_JIT_CACHE
run() {
  desc="$1"; shift
  v=$(env "$@" python bench.py --workload $wl --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 |
      python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("%.4g %.4f %s"%(j["value"], j["ms_per_step"], j["clocks"]["sm_mhz"]))')
  echo "$wl $desc : $v"
}
run "default" A=1
for blk in 64 128 256; do
  for mb in 0 2 3 4; do
    run "ring block=$blk minblocks=$mb" SPN_BLOCK=$blk SPN_MIN_BLOCKS=$mb
  done
done
for blk in 64 128; do
  for mb in 0 3 4; do
    run "regs block=$blk minblocks=$mb" SPN_NO_RING=1 SPN_BLOCK=$blk SPN_MIN_BLOCKS=$mb
    run "noprefetch block=$blk minblocks=$mb" SPN_NO_PREFETCH=1 SPN_BLOCK=$blk SPN_MIN_BLOCKS=$mb
  done
done
