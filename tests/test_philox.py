"""Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
import numpy as np
import pytest

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_numpy_kat():
    import philox
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_philox_c_oracle_kat(oracle):
    for ctr, key, want in KAT:
        got = oracle.philox(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_draw_word_addressing(oracle):
    """C oracle and NumPy restatement agree on the (seed, stream, sweep, doc, pos) -> word addressing."""
    import philox
    lib = oracle.load()
    seed = 0x1234567890abcdef
    doc = np.array([0, 0, 0, 0, 0, 7, 7, 1023, 2**31 + 5, 2**32 - 1], dtype=np.uint64)
    pos = np.array([0, 1, 2, 3, 4, 5, 1023, 77, 2**20 + 3, 9], dtype=np.uint64)
    for stream in (0, 1, 2):
        for sweep in (0, 1, 77):
            want = philox.draw_words(seed, stream, sweep, doc, pos)
            got = [lib.oracle_draw_word(seed, stream, sweep, int(d), int(q)) for d, q in zip(doc, pos)]
            assert [int(w) for w in want] == got
    # four consecutive positions of one document share one Philox block
    blk = philox.philox4x32_10(np.array([5, 9, 3, 0], dtype=np.uint32),
                               np.array([seed & 0xffffffff, seed >> 32], dtype=np.uint32))
    assert [int(x) for x in blk] == [lib.oracle_draw_word(seed, 0, 3, 9, 20 + i) for i in range(4)]


@pytest.mark.gpu
def test_philox_device_kat(gibbs):
    ctr = np.array([k[0] for k in KAT], dtype=np.uint32)
    key = np.array([k[1] for k in KAT], dtype=np.uint32)
    got = gibbs.philox_kat(ctr, key)
    want = np.array([k[2] for k in KAT], dtype=np.uint32)
    assert np.array_equal(got, want)
