#!/bin/bash
# Local wrapper: make sure every built artefact is current, then hand the command to gpurun.
#   tools/gpu.sh [--gpus N] <timeout_s> '<command>'
set -e
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
make -s -C mbb_emcee_b200/csrc
make -s -C oracle
make -s -C tools
/usr/local/graft/bin/gpurun $G --timeout $T -- "$@"
