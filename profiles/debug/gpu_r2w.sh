# single-GPU bench line of the final round-2 kernels (all legs) + the per-phase breakdown at one rank
timeout 600 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 300 gpurun_out/r2b_bench_n1.json; echo
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 profiles/debug/n_gpu_breakdown.py > gpurun_out/r2b_n1_breakdown.txt 2> gpurun_out/r2b_n1_breakdown.err
cat gpurun_out/r2b_n1_breakdown.txt
