cd uav-wrf-les-ppo-lstm_b200
for mb in 8 10 12; do
  PLUME_NVCC_EXTRA="-DPLUME_STEP_MIN_BLOCKS=$mb" python build.py --force > /dev/null 2>&1
  echo "== min blocks $mb"; (cd .. && python profiles/profile_k2.py 1048576 && python profiles/profile_k2.py 1048576 fast)
done
PLUME_NVCC_EXTRA="-DPLUME_STEP_MIN_BLOCKS=8" python build.py --force > /dev/null 2>&1
cd ..
echo "== L2 fetch 32"; PLUME_L2_FETCH=32 python profiles/profile_k2.py 1048576
echo "== L2 fetch 64"; PLUME_L2_FETCH=64 python profiles/profile_k2.py 1048576
echo "== walk 300"; PLUME_K2_WALK=300 python profiles/profile_k2.py 1048576
