# 8-GPU weak scaling of the final round-2 kernels: bench line + per-phase breakdown per rank (tight timeouts)
P=29517
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --skip-aux > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err
tail -c 600 gpurun_out/r2e_bench_n8.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+1)) profiles/debug/n_gpu_breakdown.py > gpurun_out/r2e_n8_breakdown.txt 2> gpurun_out/r2e_n8_breakdown.err
cat gpurun_out/r2e_n8_breakdown.txt
