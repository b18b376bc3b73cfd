# ncu captures of the bench workload (eager launches): launch lists of the three layouts + one --set full capture of the forward
# solve kernel, each after a plain run of the same command
set -e
export TACD_BENCH_NO_GRAPH=1
for lay in shared per_element diag; do
  CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --layout $lay"
  $CMD > gpurun_out/prof_plain_$lay.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$lay.csv $CMD > gpurun_out/prof_ncu_$lay.log 2>&1
done
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --layout ${LAYOUT:-shared}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-ilqr_forward_group} -s ${SKIP:-3} -c 1 -f -o gpurun_out/${OUT:-prof_fwd} $CMD > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log
