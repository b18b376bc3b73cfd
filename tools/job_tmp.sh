export TACD_BENCH_NO_GRAPH=1
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
$B > gpurun_out/prof_plain_cm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ilqr_forward_group -s 9 -c 3 -f -o gpurun_out/r2h_fwd_group_cm $B > gpurun_out/prof_ncu_cm.log 2>&1; tail -3 gpurun_out/prof_ncu_cm.log | cut -c1-150
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_shared.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ilqr_backward_small -s 3 -c 1 -f -o gpurun_out/r2h_bwd_small $B > gpurun_out/prof_ncu_bwd.log 2>&1; tail -2 gpurun_out/prof_ncu_bwd.log | cut -c1-150
