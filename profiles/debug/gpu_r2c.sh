set -x
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu_r2c.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2c.log
python profiles/sweep.py --hidden 32 64 --envs 4096 > gpurun_out/sweep_r2c.jsonl 2> gpurun_out/sweep_r2c.err; cat gpurun_out/sweep_r2c.jsonl; tail -3 gpurun_out/sweep_r2c.err
python profiles/debug/variant_bench.py 2>&1 | tail -1
python profiles/debug/variant_bench.py profiles/debug/libplume_b200_lt32x2.so 2>&1 | tail -1
python bench.py --skip-cpu --skip-aux > gpurun_out/bench_r2c.log 2> gpurun_out/bench_r2c.err; tail -c 300 gpurun_out/bench_r2c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['rollout_env_steps_per_sec'])
print(json.dumps(d['plume_kernels'], indent=0)[:1500])
print(d['kernels'])
PY
