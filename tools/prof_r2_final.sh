# round-2 final profile pass (developer tool): launch lists of the three cost layouts + ncu --set full captures of the kernels
# DESIGN.md cites, each after a plain run of the same command.  Outputs under gpurun_out/ (summaries go to profiles/ with
# tools/ncu_summary.py).
export TACD_BENCH_NO_GRAPH=1
for lay in shared per_element diag; do
  CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --layout $lay"
  $CMD > gpurun_out/prof_plain_$lay.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_$lay.csv $CMD > gpurun_out/prof_ncu_$lay.log 2>&1
done
cap() {  # name kernel-regex skip command...
  local out=$1 k=$2 skip=$3; shift 3
  "$@" > gpurun_out/prof_plain_$out.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/$out "$@" > gpurun_out/prof_ncu_$out.log 2>&1
  tail -2 gpurun_out/prof_ncu_$out.log | head -1 | cut -c1-200
}
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
cap r2f_fwd_group_shared ilqr_forward_group 3 $B --layout shared
cap r2f_bwd_small_shared ilqr_backward_small 3 $B --layout shared
cap r2f_fwd_group_diag ilqr_forward_group 3 $B --layout diag
cap r2f_fwd_lti32 ilqr_forward_lti 3 python tools/bench_configs.py lti32
cap r2f_bwd_lti32 ilqr_backward_lti 3 python tools/bench_configs.py lti32
cap r2f_fwd_lti8 ilqr_forward_lti 3 python tools/bench_configs.py lti8
