set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r1u.log 2>&1; tail -2 gpurun_out/pytest_gpu_r1u.log
python bench.py > gpurun_out/bench_r1u.log 2> gpurun_out/bench_r1u.err; tail -c 300 gpurun_out/bench_r1u.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1u.log 2> gpurun_out/bench_ref_r1u.err; tail -c 400 gpurun_out/bench_ref_r1u.log
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-aux > gpurun_out/plain_r1u.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1u.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-aux > gpurun_out/ncu_launch_r1u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ppo_tc_kernel -s 60 -c 1 -o gpurun_out/prof_r1u_ppo_tc -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-aux > gpurun_out/ncu_full_r1u.log 2>&1
tail -2 gpurun_out/ncu_full_r1u.log
