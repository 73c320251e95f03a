set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python profiles/profile_k2.py 1048576 && python profiles/profile_k2.py 1048576 fast && python profiles/profile_k2.py 4096 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -o gpurun_out/prof_r1g_k2 -f python profiles/profile_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k2.log
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_r1g.log 2> gpurun_out/bench_r1g.err; tail -c 600 gpurun_out/bench_r1g.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1g.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['rollout_env_steps_per_sec'], d['kernels']['rollout']); print(d['plume_kernels'])"
