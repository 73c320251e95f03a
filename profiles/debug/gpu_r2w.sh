# single-GPU: whole GPU test suite, bench line of the final round-2 kernels (all legs), per-phase breakdown, ncu capture of the update kernel
timeout 600 python -m pytest tests -m gpu -q --timeout=200 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
tail -c 300 gpurun_out/r2e_bench_n1.json; echo
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 profiles/debug/n_gpu_breakdown.py 2>/dev/null | grep "ms per rank" > gpurun_out/r2e_n1_breakdown.txt
cat gpurun_out/r2e_n1_breakdown.txt
timeout 120 python profiles/profile_r2.py iteration > gpurun_out/plain_r2e_iteration.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ppo_tc_kernel -s 45 -c 1 -f -o gpurun_out/prof_r2e_ppo_tc python profiles/profile_r2.py iteration > gpurun_out/ncu_r2e_ppo_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python profiles/profile_r2.py iteration > gpurun_out/ncu_r2e_launches.log 2>&1
tail -2 gpurun_out/ncu_r2e_ppo_tc.log
