python -m pytest tests/test_gpu_learner.py tests/test_gpu_multi.py -x -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --skip-cpu --skip-aux > gpurun_out/bench_r1k.log 2> gpurun_out/bench_r1k.err; tail -c 400 gpurun_out/bench_r1k.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1k.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['rollout_env_steps_per_sec'], d['kernels']['ppo_grad(tcgen05 f16 split)'])"
