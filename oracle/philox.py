"""Philox4x32-10 counter-based RNG, NumPy restatement (TEST INFRASTRUCTURE ONLY).

This file belongs to ``oracle/``: it may be imported by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s cpu_baseline leg, never by the
product package ``lda_thesis_b200``.

Algorithm: Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11),
Philox-4x32 with 10 rounds, as published in Random123 (philox.h).  Pinned by the
Random123 known-answer vectors quoted in SURVEY.md §8(c)3 (see tests/test_philox.py).

Draw addressing used by every sampler in this repo (GPU, C oracle, patched
reference) -- one 32-bit word per (stream, sweep, global document id d, position p of the draw in d):

    ctr = (p >> 2, d, sweep, stream)    key = (lo32(seed), hi32(seed))
    word = philox4x32_10(ctr, key)[p & 3]

(addressing by document keeps the stream independent of how documents are sharded over GPUs, and
lets a thread that walks one document reuse each Philox block for four consecutive draws)

The reference itself never seeds its RNG, so this addressing is the
repo's own convention; what it replaces is the single ``multinom_draw(1, prob)`` call
per pair at LabeledLDA.py:119 / CascadeLDA.py:415 / HSLDA.py:261.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_SWEEP = 0   # one word per training draw
STREAM_INIT = 1    # one word per draw for device-side z initialisation
STREAM_TEST = 2    # test-time (frozen phi) chains
STREAM_HOST = 3    # host-side helpers (synthetic HSLDA state, ...)
STREAM_TEST_INIT = 4   # start state of the test-time chains (prep4test)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32  ->  (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    key = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = key[..., 0].copy(), key[..., 1].copy()
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & MASK32
        k1 = (k1 + np.uint64(PHILOX_W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def draw_words(seed, stream, sweep, doc, pos):
    """32-bit word for each (global document id, position inside the document) pair (array-likes)."""
    doc = np.atleast_1d(np.asarray(doc, dtype=np.uint64))
    pos = np.atleast_1d(np.asarray(pos, dtype=np.uint64))
    doc, pos = np.broadcast_arrays(doc, pos)
    ctr = np.empty(pos.shape + (4,), dtype=np.uint64)
    ctr[..., 0] = (pos >> np.uint64(2)) & MASK32
    ctr[..., 1] = doc & MASK32
    ctr[..., 2] = np.uint64(sweep & 0xFFFFFFFF)
    ctr[..., 3] = np.uint64(stream & 0xFFFFFFFF)
    key = np.empty(pos.shape + (2,), dtype=np.uint64)
    key[..., 0] = np.uint64(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint64((seed >> 32) & 0xFFFFFFFF)
    out = philox4x32_10(ctr, key)
    return np.take_along_axis(out, (pos & np.uint64(3)).astype(np.int64)[..., None], axis=-1)[..., 0]


def draw_word(seed, stream, sweep, doc, pos):
    return int(draw_words(seed, stream, sweep, [doc], [pos])[0])


def u01_f64(word):
    """fp64 uniform in (0,1) with 32-bit resolution: (word + 0.5) * 2^-32 (exact)."""
    return (np.asarray(word, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)


def u01_f32(word):
    """fp32 uniform in [0,1) with 24-bit resolution: (word >> 8) * 2^-24 (exact)."""
    w = np.asarray(word, dtype=np.uint32) >> np.uint32(8)
    return w.astype(np.float32) * np.float32(1.0 / 16777216.0)
