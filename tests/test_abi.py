"""The C-ABI library loads without a GPU and exports every symbol include/plume_b200.h
declares (no compute calls here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "plume_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plume_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_declared_symbols():
    import __graft_entry__ as ge
    ge.build()
    import uav_wrf_les_ppo_lstm_b200 as pb
    lib = pb._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(pb._lib.EXPORTED_SYMBOLS) == declared
    assert lib.plume_abi_version() == 4


def test_struct_layouts_match_header():
    import ctypes as C
    import uav_wrf_les_ppo_lstm_b200 as pb
    L = pb._lib
    assert C.sizeof(L.EnvConfig) == 4 * 4 + 9 * 8 + 8 + 8
    assert C.sizeof(L.EnvState) == 8 + 20 * 8
    assert C.sizeof(L.LstmParams) == 16 + 8 * 8
    assert C.sizeof(L.RolloutBuffers) == 28 * 8
    assert C.sizeof(L.PpoBatch) == 8 * 8
    header = open(os.path.join(ROOT, "include", "plume_b200.h")).read()
    for name, (off, _) in L.MLP_OFFSETS.items():
        pass
    offs = dict(re.findall(r"#define PLUME_OFF_(\w+) (\d+)", header))
    want = {"W1": "feature.0.weight", "B1": "feature.0.bias", "G1": "feature.1.weight", "BE1": "feature.1.bias",
            "W2": "feature.3.weight", "B2": "feature.3.bias", "G2": "feature.4.weight", "BE2": "feature.4.bias",
            "WA": "actor.weight", "BA": "actor.bias", "WC": "critic.weight", "BC": "critic.bias"}
    for macro, key in want.items():
        assert int(offs[macro]) == L.MLP_OFFSETS[key][0]
    assert int(re.search(r"#define PLUME_MLP_PARAMS (\d+)", header).group(1)) == L.MLP_PARAMS


def test_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import uav_wrf_les_ppo_lstm_b200 as pb
    with pytest.raises(RuntimeError):
        pb.VecMethaneEnv(4, device="cuda")
    m = pb.PPOActorCritic(device="cpu")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 6))
