timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu_r2j.log 2>&1; tail -12 gpurun_out/pytest_gpu_r2j.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import uav_wrf_les_ppo_lstm_b200 as pb
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=256)
tr.train_iteration(); torch.cuda.synchronize()
t=time.time(); tr.train_iteration(); torch.cuda.synchronize(); print("minibatch-256 iteration:", (time.time()-t)*1e3, "ms")
PY
python profiles/debug/variant_bench.py 2>&1 | tail -1
