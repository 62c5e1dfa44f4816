#!/bin/bash
# Multi-GPU visit: /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- tools/gpu_scaling.sh N [full]
#   weak-scaling bench line (the driver's command), the same workload strongly scaled (1e5 sources
#   in total), and the copy-only host<->device probe; outputs gpurun_out/r02_scale_*_N.json
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = 1 ] && TR="python"
EXTRA="--no-cpu-baseline"
[ "$2" = full ] && EXTRA=""
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > $O/r02_scale_weak_$N.json 2> $O/r02_scale_weak_$N.err
tail -2 $O/r02_scale_weak_$N.err; head -c 400 $O/r02_scale_weak_$N.json; echo
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-sublegs > $O/r02_scale_strong_$N.json 2> $O/r02_scale_strong_$N.err
tail -2 $O/r02_scale_strong_$N.err; head -c 400 $O/r02_scale_strong_$N.json; echo
timeout 300 $TR tools/h2d_probe.py > $O/r02_h2d_probe_$N.json 2> $O/r02_h2d_probe_$N.err
cat $O/r02_h2d_probe_$N.json; tail -2 $O/r02_h2d_probe_$N.err
echo DONE
