python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/profile_k2.py 1048576 && python profiles/profile_k2.py 1048576 fast
python profiles/profile_lstm_train.py 64 && python profiles/profile_lstm_train.py 512 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_train_kernel -s 40 -c 1 -o gpurun_out/prof_r1h_lstm_train -f python profiles/profile_lstm_train.py 64 > gpurun_out/ncu_lt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -o gpurun_out/prof_r1h_k2 -f python profiles/profile_k2.py 1048576 > gpurun_out/ncu_k2.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1h.log 2> gpurun_out/bench_r1h.err; tail -c 400 gpurun_out/bench_r1h.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1h.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['rollout_env_steps_per_sec']); print(d['lstm_train']); print({k:(v['ms'],round(v['frac'],3)) for k,v in d['plume_kernels'].items()})"
