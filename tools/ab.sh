#!/bin/bash
# A/B timing of alternative builds (tools/_build/lib_*.so) and runtime switches of the library:
# device-resident legs only.   usage: tools/ab.sh [workloads...]   (default: cfg5 cfg2)
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-passband"
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], '%.4f ms  %.3e evals/s' % (d['ms_per_step'], d['value']))" "$1" "$2"; }
run() { # name env workload
  timeout 120 env $2 $B --workload $3 > gpurun_out/ab_$1_$3.json 2> gpurun_out/ab_$1_$3.err || tail -3 gpurun_out/ab_$1_$3.err
  show gpurun_out/ab_$1_$3.json "$1 $3"; }
for w in ${@:-cfg5 cfg2}; do
  run default "X=1" $w
  if [ $w = cfg5 ]; then
    run notma "MBB_B200_NO_TMA=1" $w
    run nostage "MBB_B200_NO_STAGE_DATA=1" $w
  fi
  for lib in tools/_build/lib_*.so; do
    [ -e "$lib" ] || continue
    n=$(basename $lib .so)
    run $n "MBB_B200_LIB=$PWD/$lib" $w
  done
done
