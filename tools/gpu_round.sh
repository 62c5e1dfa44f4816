#!/bin/bash
# One GPU box visit: parity suite, smoke, the bench lines and the ncu evidence that profiles/ summarises.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- tools/gpu_round.sh r01
# Everything lands in gpurun_out/<tag>_*; tools/ncu_summary.py turns it into profiles/ here.
T=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/${T}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $O/${T}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
python bench.py --workload cfg2 --steps 5 --warmup 3 > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
python tools/bench_chain.py > $O/${T}_chain_cfg4.json 2> $O/${T}_chain_cfg4.err
tools/_build/fp64_probe > $O/${T}_fp64_probe.json 2>&1
tools/_build/rcp_probe > $O/${T}_rcp_probe.json 2>&1
tools/_build/exp_probe > $O/${T}_exp_probe.json 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-passband"
for w in cfg5 cfg2; do
  $B --workload $w > $O/${T}_plain_$w.log 2>&1 || continue      # the program must exit 0 without ncu first
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_$w.csv \
      $B --workload $w > $O/${T}_ncu_l_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:loglike_delta -s 3 -c 1 -o $O/${T}_prof_cfg5 -f \
    $B --workload cfg5 > $O/${T}_ncu_f_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loglike_nodes -s 3 -c 1 -o $O/${T}_prof_cfg2 -f \
    $B --workload cfg2 > $O/${T}_ncu_f_cfg2.log 2>&1
head -c 600 $O/${T}_bench_cfg5.json; echo
echo DONE
