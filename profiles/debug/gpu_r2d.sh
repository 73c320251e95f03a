set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/pytest_gpu_r2d.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2d.log
timeout 300 python profiles/debug/variant_bench.py 2>&1 | tail -1
timeout 300 python profiles/debug/variant_bench.py profiles/debug/libplume_b200_lockstep.so 2>&1 | tail -1
timeout 600 python bench.py --skip-cpu > gpurun_out/bench_r2d.log 2> gpurun_out/bench_r2d.err; tail -c 300 gpurun_out/bench_r2d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['rollout_env_steps_per_sec'])
for k,v in d['plume_kernels'].items(): print(k, v['ms'], v['frac'])
print(d['kernels'])
PY
