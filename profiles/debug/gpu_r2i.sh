set -x
timeout 900 python profiles/learning_curve.py --envs 1 --horizon 256 --episodes 2000 --seeds 0 1 2 > gpurun_out/learning_n1.jsonl 2> gpurun_out/learning_n1.err; cat gpurun_out/learning_n1.jsonl; tail -3 gpurun_out/learning_n1.err
timeout 600 python profiles/learning_curve.py --envs 8 --horizon 32 --episodes 2000 --seeds 0 1 > gpurun_out/learning_n8.jsonl 2> gpurun_out/learning_n8.err; cat gpurun_out/learning_n8.jsonl; tail -3 gpurun_out/learning_n8.err
timeout 900 python bench.py > gpurun_out/bench_r2i.log 2> gpurun_out/bench_r2i.err; tail -c 500 gpurun_out/bench_r2i.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['rollout_env_steps_per_sec'], d['e2e'])
print(json.dumps(d['other_configs'], indent=1)[:3000])
print(d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2i.log 2>&1; tail -c 600 gpurun_out/bench_ref_r2i.log
