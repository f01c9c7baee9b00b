#!/bin/bash
# A/B timing of alternative builds of the library (HERMNET_B200_LIB) on the $BIG^3 lattice
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LOG=gpurun_out/tc_ab.log
: > $LOG
for lib in $LIBS; do
  echo "=== $lib" >> $LOG
  HERMNET_B200_LIB=$PWD/hermnet_b200/lib/$lib timeout 300 python tests/tc_debug.py ${BIG:-48} ${STAGES:-fwd} 2>&1 | grep -E "ms per launch|vs quad|Error|error" >> $LOG
done
cat $LOG
