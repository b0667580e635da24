TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_peer_comm.py -x -q 2>&1 | tail -3
timeout 200 $TR --master-port 29521 tools/mc_sum_bench.py 2>gpurun_out/mc8.err | tail -1 > gpurun_out/mc_sum_n8.json; cat gpurun_out/mc_sum_n8.json
timeout 200 $TR --master-port 29522 bench.py --gpus 8 --workload cfg5 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 --mc-sum step 2>gpurun_out/b8.err | tail -1 > gpurun_out/r01b_bench_cfg5_n8_step.json; tail -c 900 gpurun_out/r01b_bench_cfg5_n8_step.json
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --workload cfg4 --steps 2 2>gpurun_out/b84.err | tail -1 > gpurun_out/r01b_bench_cfg4_n8.json; tail -c 900 gpurun_out/r01b_bench_cfg4_n8.json
