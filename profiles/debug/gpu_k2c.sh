python profiles/profile_k2.py 1048576
PLUME_L2_FETCH=32 python profiles/profile_k2.py 1048576
PLUME_L2_FETCH=64 python profiles/profile_k2.py 1048576
PLUME_L2_FETCH=128 python profiles/profile_k2.py 1048576
python profiles/profile_k2.py 262144
python profiles/profile_k2.py 65536
python profiles/profile_k2.py 4096
