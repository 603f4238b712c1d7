"""CPU: the numpy restatement of Philox4x32-10 (oracle/philox.py) against the known-answer
vectors published with Random123 (kat_vectors: philox4x32 10 rounds)."""
import numpy as np

from oracle import philox


def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kat:
        r = philox.philox4x32(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert [int(x) for x in r] == list(out)


def test_first_uniform_range():
    u = philox.first_uniform(0x123456789abcdef0, 7, 1, np.arange(100000))
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1.0 / 12.0) < 2e-3
