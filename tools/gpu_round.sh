#!/bin/bash
# One GPU box visit (round-1 form; round 2 uses tools/gpu_visit.sh, tools/ncu_capture.py and tools/gpu_scaling.sh):
# parity suite, smoke, the bench lines and the ncu evidence that profiles/ summarises.
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
# full captures, summarised on the box (each .ncu-rep is 25-40 MB and gpurun_out/ may carry 64 MiB
# back: the reports are deleted after the summaries are written; keep one by passing KEEP_REP=name)
full() { # name  kernel-regex  command...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/${T}_prof_$name -f \
      "$@" > $O/${T}_ncu_f_$name.log 2>&1
  python tools/ncu_summary.py full $O/${T}_prof_$name.ncu-rep > $O/${T}_full_$name.md 2>> $O/${T}_ncu_f_$name.log
  python tools/ncu_summary.py facts $O/${T}_prof_$name.ncu-rep > $O/${T}_facts_$name.json 2>> $O/${T}_ncu_f_$name.log
  [ "$KEEP_REP" = "$name" ] || rm -f $O/${T}_prof_$name.ncu-rep
}
full cfg5 loglike_delta $B --workload cfg5
full cfg2 loglike_nodes $B --workload cfg2
full cfg2_gauss loglike_gauss_thread $B --workload cfg2 --math gauss
full cfg5p_gauss loglike_gauss_thread $B --workload cfg5p --math gauss
full ens ens_resident python tools/sampler_probe.py 4 20000
python tools/sampler_probe.py 10 > $O/${T}_sampler.json 2> $O/${T}_sampler.err
head -c 600 $O/${T}_bench_cfg5.json; echo
echo DONE
