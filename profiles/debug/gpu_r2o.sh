set -x
for w in k1 iteration head64 head256; do python profiles/profile_r2.py $w > gpurun_out/plain_r2_$w.log 2>&1 || exit 1; done
ncu --set full --clock-control none --import-source on -k regex:generate_fields_f32 -s 2 -c 1 -f -o gpurun_out/prof_r2_k1 python profiles/profile_r2.py k1 > gpurun_out/ncu_r2_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_pipe_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_rollout python profiles/profile_r2.py iteration > gpurun_out/ncu_r2_rollout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ppo_tc_kernel -s 45 -c 1 -f -o gpurun_out/prof_r2_ppo_tc python profiles/profile_r2.py iteration > gpurun_out/ncu_r2_ppo_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stop_head_segment_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_head64 python profiles/profile_r2.py head64 > gpurun_out/ncu_r2_head64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stop_head_stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_head256 python profiles/profile_r2.py head256 > gpurun_out/ncu_r2_head256.log 2>&1
tail -2 gpurun_out/ncu_r2_*.log
ls -la gpurun_out/*.ncu-rep | tail -6
