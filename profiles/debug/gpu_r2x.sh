# 2 GPUs: the multi-GPU equivalence test, the bench line and the per-phase breakdown
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=150 2>&1 | tail -3
cp gpurun_out/r2_multi_gpu_equivalence_n2.log gpurun_out/r2e_multi_gpu_equivalence_n2.log 2>/dev/null
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --skip-aux > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -c 400 gpurun_out/r2e_bench_n2.json; echo
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 profiles/debug/n_gpu_breakdown.py 2>/dev/null | tee gpurun_out/r2e_n2_breakdown.txt
