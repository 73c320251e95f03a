cd uav-wrf-les-ppo-lstm_b200 && PLUME_NVCC_EXTRA="-DPLUME_TC_TIMELINE" python build.py --force > /dev/null 2>&1; cd ..
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as pb
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096*256//4)
tr.train_iteration(); torch.cuda.synchronize()
PY
