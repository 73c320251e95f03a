python -m pytest tests/test_gpu_learner.py -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --skip-cpu --skip-aux > gpurun_out/bench_r1o.log 2> gpurun_out/bench_r1o.err; tail -c 300 gpurun_out/bench_r1o.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1o.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernels']['ppo_grad(tcgen05 f16 split)'])"
bash profiles/debug/gpu_tl.sh 2>&1 | grep timeline | head -2
