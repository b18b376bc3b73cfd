TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29513 tools/bench_config5_sharded.py --total 32768 2>/dev/null | grep '^{' > gpurun_out/r2_cfg5_n8.jsonl; cat gpurun_out/r2_cfg5_n8.jsonl
timeout 400 $TR --master-port 29511 tools/bench_ppo_update.py 2>/dev/null | grep '^{' > gpurun_out/r2_ppo_update_n8.jsonl; cat gpurun_out/r2_ppo_update_n8.jsonl
