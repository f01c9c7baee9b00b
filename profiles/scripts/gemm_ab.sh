#!/bin/bash
# A/B timing of alternative builds of the GEMM (HERMNET_B200_LIB) + one ncu --set full capture of the product build
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LOG=gpurun_out/gemm_ab.log
: > $LOG
for lib in $LIBS; do
  echo "=== $lib" >> $LOG
  HERMNET_B200_LIB=$PWD/hermnet_b200/lib/$lib timeout 300 python profiles/gemm_time.py 10 2>&1 | tail -12 >> $LOG
done
cat $LOG
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 30 -c 40 -f -o gpurun_out/prof_gemm \
      python profiles/gemm_time.py 1 > gpurun_out/prof_gemm.log 2>&1
  tail -2 gpurun_out/prof_gemm.log
fi
