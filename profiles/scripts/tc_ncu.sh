#!/bin/bash
# ncu --set full capture of the tensor-core edge kernels on a 48^3 (or $BIG^3) lattice; LIB selects an A/B build
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
[ -n "$LIB" ] && export HERMNET_B200_LIB=$PWD/hermnet_b200/lib/$LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_tc -s ${SKIP:-3} -c ${COUNT:-3} -f -o gpurun_out/${OUT:-prof_tc} \
    python tests/tc_debug.py ${BIG:-48} ${STAGES:-fwd dst src} > gpurun_out/${OUT:-prof_tc}.log 2>&1
echo "exit $?" >> gpurun_out/${OUT:-prof_tc}.log
tail -3 gpurun_out/${OUT:-prof_tc}.log
