#!/bin/bash
# bring-up of the tensor-core edge kernels: per-stage diagnostics (tests/tc_debug.py), each stage in its own process
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LOG=gpurun_out/tc_debug.log
: > $LOG
for st in ${STAGES:-"plan phi fwd|fwd0|dst|dst0|src"}; do :; done
IFS='|' read -ra LIST <<< "${STAGES:-plan phi fwd|fwd0|dst|dst0|src}"
for st in "${LIST[@]}"; do
  echo "=== stage $st" >> $LOG
  timeout 240 python tests/tc_debug.py ${NSIDE:-6} $st >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
if [ -n "$BIG" ]; then
  echo "=== timing $BIG^3" >> $LOG
  timeout 400 python tests/tc_debug.py $BIG plan fwd dst src >> $LOG 2>&1
  echo "exit $?" >> $LOG
fi
tail -80 $LOG
