python -m pytest tests/test_gpu_learner.py tests/test_gpu_tc_gemm.py -x -q 2>&1 | tail -25
python bench.py --steps 5 --warmup 3 --skip-cpu --skip-aux > gpurun_out/bench_r1n.log 2> gpurun_out/bench_r1n.err; tail -c 300 gpurun_out/bench_r1n.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1n.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernels']['ppo_grad(tcgen05 3xTF32)'])"
