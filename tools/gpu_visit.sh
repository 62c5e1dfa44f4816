#!/bin/bash
# Round-2 box visits.  usage: tools/gpu_visit.sh <step>...   (outputs under gpurun_out/r02_*)
O=gpurun_out
mkdir -p $O
for step in "$@"; do
case $step in
ens_tests)
  timeout 900 python -m pytest tests/test_ensemble_gpu.py -x -q 2>&1 | tail -15 | tee $O/r02_ens_tests.log ;;
all_tests)
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r02_all_tests.log ;;
sampler)
  timeout 600 python tools/sampler_probe.py 50 > $O/r02_sampler.json 2> $O/r02_sampler.err; cat $O/r02_sampler.json; tail -3 $O/r02_sampler.err ;;
sampler_ncu)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ens_resident -s 1 -c 1 \
      -o $O/r02_prof_ens_resident -f python tools/sampler_probe.py 4 20000 > $O/r02_ncu_ens.log 2>&1
  python tools/ncu_summary.py full $O/r02_prof_ens_resident.ncu-rep > $O/r02_full_ens_resident.md 2>> $O/r02_ncu_ens.log
  python tools/ncu_summary.py facts $O/r02_prof_ens_resident.ncu-rep > $O/r02_facts_ens_resident.json 2>> $O/r02_ncu_ens.log
  ncu -i $O/r02_prof_ens_resident.ncu-rep --page source --csv > $O/r02_src_ens_resident.csv 2>> $O/r02_ncu_ens.log
  rm -f $O/r02_prof_ens_resident.ncu-rep
  tail -5 $O/r02_ncu_ens.log ;;
variants)
  timeout 600 python tools/bench_variants.py > $O/r02_variants.json 2> $O/r02_variants.err; cat $O/r02_variants.json ;;
variants_early)
  MBB_B200_LIB=$PWD/tools/_build/lib_earlystop.so timeout 600 python tools/bench_variants.py > $O/r02_variants_early.json 2> $O/r02_variants_early.err
  cat $O/r02_variants_early.json
  MBB_B200_LIB=$PWD/tools/_build/lib_earlystop.so timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -5 | tee $O/r02_early_tests.log ;;
bench)
  timeout 1500 python bench.py --steps 10 --warmup 3 > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err; head -c 3000 $O/r02_bench_cfg5.json; tail -3 $O/r02_bench_cfg5.err ;;
bench_ref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; cat $O/r02_bench_ref.json; tail -3 $O/r02_bench_ref.err ;;
ncu_all)
  timeout 1500 python tools/ncu_capture.py r02 2>&1 | tee $O/r02_ncu_capture.log ;;
bench_quick)
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_quick.json 2> $O/r02_bench_quick.err; head -c 1500 $O/r02_bench_quick.json; tail -12 $O/r02_bench_quick.err ;;
probe)
  for v in "1 0" "0 1" "1 1" "0 0"; do timeout 120 tools/_build/delta_probe $v 20000 20 2>&1 | tee -a $O/r02_probe.jsonl; done ;;
probe_ncu)   # PROBE_V="0 1": per-source-line executed instruction counts of one variant
  V=${PROBE_V:-0 1}; T=$(echo $V | tr -d ' ')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:loglike_delta -s 3 -c 1 \
      -o $O/r02_probe_$T -f tools/_build/delta_probe $V 4000 2 > $O/r02_probe_ncu_$T.log 2>&1
  ncu -i $O/r02_probe_$T.ncu-rep --page source --csv --print-source cuda > $O/r02_probe_src_cuda_$T.csv 2>> $O/r02_probe_ncu_$T.log
  ncu -i $O/r02_probe_$T.ncu-rep --page source --csv > $O/r02_probe_src_sass_$T.csv 2>> $O/r02_probe_ncu_$T.log
  python tools/ncu_summary.py full $O/r02_probe_$T.ncu-rep > $O/r02_probe_full_$T.md 2>> $O/r02_probe_ncu_$T.log
  rm -f $O/r02_probe_$T.ncu-rep; tail -3 $O/r02_probe_ncu_$T.log ;;
ens_probe)    # ENS_ARGS="1 0 20000 20 1"
  for a in "1 0 20000 20 1" "1 0 20000 20 0" "0 1 20000 20 1"; do timeout 120 tools/_build/ens_probe $a 2>&1 | tee -a $O/r02_ens_probe.jsonl; done ;;
ens_probe_ncu)
  V=${PROBE_V:-1 0}; T=$(echo $V | tr -d ' ')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ens_resident -s 1 -c 1 \
      -o $O/r02_ensp_$T -f tools/_build/ens_probe $V 4000 10 1 > $O/r02_ensp_ncu_$T.log 2>&1
  ncu -i $O/r02_ensp_$T.ncu-rep --page source --csv > $O/r02_ensp_src_sass_$T.csv 2>> $O/r02_ensp_ncu_$T.log
  python tools/ncu_summary.py full $O/r02_ensp_$T.ncu-rep > $O/r02_ensp_full_$T.md 2>> $O/r02_ensp_ncu_$T.log
  rm -f $O/r02_ensp_$T.ncu-rep; tail -3 $O/r02_ensp_ncu_$T.log ;;
*) echo "unknown step $step" ;;
esac
done
echo DONE
