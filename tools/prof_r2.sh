# one ncu --set full capture of the forward solve kernel (bench workload, eager launches), after a plain run of the same command
set -e
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline ${LAYOUT:+--layout $LAYOUT}"
export TACD_BENCH_NO_GRAPH=1
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-ilqr_forward_group} -s ${SKIP:-3} -c 1 -f -o gpurun_out/${OUT:-prof_fwd} $CMD > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_ncu.log
