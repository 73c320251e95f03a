timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/pytest_gpu_r2g.log 2>&1; tail -12 gpurun_out/pytest_gpu_r2g.log
python profiles/debug/variant_bench.py profiles/debug/libplume_b200_rtl.so > gpurun_out/rtl.log 2>&1; grep -E "mlp half" gpurun_out/rtl.log | head -4; grep -E "rollout timeline t=101" gpurun_out/rtl.log | head -4
python profiles/debug/variant_bench.py 2>&1 | tail -1
python bench.py --skip-cpu --skip-aux > gpurun_out/bench_r2g.log 2> gpurun_out/bench_r2g.err; tail -c 300 gpurun_out/bench_r2g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2g.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['rollout_env_steps_per_sec'])
print(d['kernels']['rollout'])
PY
