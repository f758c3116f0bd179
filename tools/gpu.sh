#!/bin/bash
# Build the in-tree .so, then run a script on the GPU box:  tools/gpu.sh [--gpus N] <script.sh> <logname> [timeout]
set -e
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
python -m infinicube_b200.build > /dev/null
python -c "import __graft_entry__ as g; g.build()" > /dev/null
/usr/local/graft/bin/gpurun $GP --timeout ${3:-1500} -- "bash $1" > gpurun_out/$2.out 2>&1 || true
tail -60 gpurun_out/$2.out
