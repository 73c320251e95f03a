python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_r1j.log 2> gpurun_out/bench_r1j.err; tail -c 400 gpurun_out/bench_r1j.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1j.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['rollout_env_steps_per_sec'], d['kernels']['rollout'])"
