"""Runs pytest in-process against a variant build of the library: python profiles/debug/pytest_variant.py <lib.so> <pytest args...>"""
import sys
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as pb
pb._lib.LIB_PATH = sys.argv[1]
import pytest
sys.exit(pytest.main(sys.argv[2:]))
