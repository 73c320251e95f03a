timeout 120 python profiles/profile_r2.py iteration > gpurun_out/plain_r2b_iteration.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ppo_tc_kernel -s 45 -c 1 -f -o gpurun_out/prof_r2b_ppo_tc python profiles/profile_r2.py iteration > gpurun_out/ncu_r2b_ppo_tc.log 2>&1
tail -3 gpurun_out/ncu_r2b_ppo_tc.log
