"""Known-answer vectors of Random123's philox4x32-10 (kat_vectors) against the oracle's numpy
restatement and against the product's device function compiled for the host."""
import numpy as np

from oracle import philox as ph

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_kat_oracle():
    for ctr, key, want in KAT:
        got = ph.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_uniform_ranges():
    r = np.array([0, 1, 255, 256, 0xFFFFFFFF], dtype=np.uint32)
    u = ph.uniform24(r)
    assert u.min() >= 0.0 and u.max() < 1.0
    u0 = ph.uniform24_open0(r)
    assert u0.min() > 0.0 and u0.max() <= 1.0
    d = ph.uniform53(r, r)
    assert d.min() >= 0.0 and d.max() < 1.0


def test_box_muller_moments():
    idx = np.arange(200000)
    z0, z1 = ph.step_noise64(1234, 7, 3, idx)
    z = np.concatenate([z0, z1])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(np.mean(z0 * z1)) < 0.01


def test_field_stream_moments_and_layout():
    """The field stream (one Philox4x32-7 call per four cells): the draws of a whole plume are standard normal /
    uniform, pairs are uncorrelated, and the four cells of a quad take their bits from the documented places."""
    cells = np.arange(500 * 500)
    z, u = ph.field_noise64(99, 5, 2, cells)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and np.abs(z).max() < 5.3     # 20-bit radius
    assert abs(u.mean() - 0.5) < 0.005 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(np.mean(z[0::2] * z[1::2])) < 0.01 and abs(np.corrcoef(z, u)[0, 1]) < 0.01
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.01
    w = [int(x) for x in ph.field_words(99, 5, 2, 1000)]              # quad 250
    k0, k1 = ph.seed_key(99)
    assert tuple(w) == tuple(int(x) for x in ph.philox4x32(250, 2, 5, ph.TAG_FIELD, k0, k1, 7))
    _, uq = ph.field_noise64(99, 5, 2, np.arange(1000, 1004))
    want = [(w[0] & 0xFFF), (w[1] & 0xFFF), (w[3] & 0xFFF), ((w[3] >> 12) & 0xFFF)]
    assert [int(round(float(x) * 4096)) for x in uq] == want
