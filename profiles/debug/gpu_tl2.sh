# per-tile and per-kernel cycle timeline of ppo_tc_kernel from a -DPLUME_TC_TIMELINE build linked HERE into
# profiles/debug/libplume_b200_tl.so (git-ignored; the product library is untouched)
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as pb
pb._lib.LIB_PATH = 'profiles/debug/libplume_b200_tl.so'
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096*256//4)
tr.train_iteration(); torch.cuda.synchronize()
PY
