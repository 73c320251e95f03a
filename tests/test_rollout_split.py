"""Host logic of the overlapped rollout collection (rollout.overlap_split): no GPU needed."""
from uav_wrf_les_ppo_lstm_b200.rollout import overlap_split


def test_split_of_the_bench_shape():
    # 4096 envs = 128 lockstep CTAs on 148 SMs: the head of the first 32 rows fits under the other 224
    assert overlap_split(256, 4096, 148) == (32, 224)


def test_no_split_without_idle_sms_or_without_gain():
    assert overlap_split(256, 4736, 148) is None          # 148 tiles: no SM left
    assert overlap_split(256, 8192, 148) is None
    assert overlap_split(256, 64, 148) is None            # the head is cheaper than a second launch
    assert overlap_split(16, 4096, 148) is None


def test_split_properties():
    for T in (64, 128, 256, 512):
        for N in (256, 1024, 2048, 3072, 4096, 4500):
            sp = overlap_split(T, N, 148)
            if sp is None:
                continue
            assert len(sp) == 2 and sum(sp) == T and sp[0] % 8 == 0 and 8 <= sp[0] <= T // 2
