"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY -- see gibbs_oracle.h).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Never by the product package lda_thesis_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
f64p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    i32, i64, u32, u64, dbl = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    lib.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_draw_word.argtypes = [u64, u32, u32, u64, u64]
    lib.oracle_draw_word.restype = u32
    lib.oracle_openmp_threads.restype = C.c_int
    lib.oracle_init_z.argtypes = [i64, i64p, i64p, i32p, i32p, u64, u64]
    lib.oracle_counts_build.argtypes = [i64, i64p, i32p, i32p, i32p, i64p, i32p, i32, i32, i32, i32p, i32p, i32p]
    lib.oracle_llda_exact_sweep.argtypes = [i64, i64, i64p, i32p, i32p, i32p, i64p, i32p, i32, i32, i32, dbl, dbl,
                                            i32p, i32p, i32p, u64, u32, u64]
    lib.oracle_llda_snapshot_sweep.argtypes = [i64, i64p, i32, i64p, i32p, i32p, i32p, i64p, i32p, i32, i32, i32,
                                               dbl, dbl, i32p, i32p, i32p, u64, u32, u64, i32]
    lib.oracle_test_chains.argtypes = [i32, i32, f64p, dbl, dbl, i64, i64p, i32p, i32p, i64p, i32p, i32p, i32, i32, i32,
                                       u64, i64, f64p]
    lib.oracle_test_chains.restype = C.c_int
    for n in ("oracle_init_z", "oracle_counts_build", "oracle_llda_exact_sweep", "oracle_llda_snapshot_sweep"):
        getattr(lib, n).restype = C.c_int
    _lib = lib
    return lib


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def _p(x, tp):
    return None if x is None else x.ctypes.data_as(tp)


def ldk_of(K):
    return (int(K) + 31) // 32 * 32


def philox(ctr4, key2):
    lib = load()
    ctr4 = _a(ctr4, np.uint32).reshape(-1, 4)
    key2 = _a(key2, np.uint32).reshape(-1, 2)
    out = np.empty_like(ctr4)
    for i in range(ctr4.shape[0]):
        lib.oracle_philox4x32_10(_p(ctr4[i], u32p), _p(key2[i], u32p), _p(out[i], u32p))
    return out


def openmp_threads():
    return int(load().oracle_openmp_threads())


class LldaOracle(object):
    """Host mirror of one device shard: owns z and the three count arrays in the device layout."""

    def __init__(self, doc_ptr, word, freq, lab_ptr, lab_idx, K, V, alpha, beta, seed=0, z=None, doc_base=0):
        self.lib = load()
        self.doc_ptr = _a(doc_ptr, np.int64)
        self.word = _a(word, np.int32)
        self.freq = None if freq is None else _a(freq, np.int32)
        self.lab_ptr = _a(lab_ptr, np.int64)
        self.lab_idx = _a(lab_idx, np.int32)
        self.D = self.doc_ptr.shape[0] - 1
        self.N = int(self.doc_ptr[-1])
        self.K, self.V, self.ldk = int(K), int(V), ldk_of(K)
        self.alpha, self.beta, self.seed, self.doc_base = float(alpha), float(beta), int(seed), int(doc_base)
        self.sweep = 0
        if z is None:
            self.z = np.zeros(self.N, dtype=np.int32)
            rc = self.lib.oracle_init_z(self.D, _p(self.doc_ptr, i64p), _p(self.lab_ptr, i64p), _p(self.lab_idx, i32p),
                                        _p(self.z, i32p), self.seed, self.doc_base)
            if rc:
                raise RuntimeError("oracle_init_z failed (%d)" % rc)
        else:
            self.z = _a(z, np.int32).copy()
        self.n_wk_pad = np.zeros((self.V, self.ldk), dtype=np.int32)
        self.n_dk_act = np.zeros(int(self.lab_ptr[-1]), dtype=np.int32)
        self.n_k = np.zeros(self.K, dtype=np.int32)
        self.rebuild()

    def rebuild(self):
        rc = self.lib.oracle_counts_build(self.D, _p(self.doc_ptr, i64p), _p(self.word, i32p), _p(self.freq, i32p),
                                          _p(self.z, i32p), _p(self.lab_ptr, i64p), _p(self.lab_idx, i32p),
                                          self.K, self.V, self.ldk, _p(self.n_wk_pad, i32p), _p(self.n_dk_act, i32p),
                                          _p(self.n_k, i32p))
        if rc:
            raise RuntimeError("oracle_counts_build failed (%d): z outside the label list?" % rc)

    def add_word_topic_counts(self, word, topic, count):
        """n_wk[word, topic] += count (COO, duplicates accumulate); n_k and n_dk are left alone.  Used to
        reproduce the spurious initial counts of SubLDA.__init__ (CascadeLDA.py:382-385)."""
        np.add.at(self.n_wk_pad, (np.asarray(word, dtype=np.int64), np.asarray(topic, dtype=np.int64)),
                  np.asarray(count, dtype=np.int32))

    @property
    def n_wk(self):
        return self.n_wk_pad[:, :self.K]

    def exact_sweep(self, n=1):
        for _ in range(n):
            rc = self.lib.oracle_llda_exact_sweep(0, self.D, _p(self.doc_ptr, i64p), _p(self.word, i32p),
                                                  _p(self.freq, i32p), _p(self.z, i32p), _p(self.lab_ptr, i64p),
                                                  _p(self.lab_idx, i32p), self.K, self.V, self.ldk, self.alpha,
                                                  self.beta, _p(self.n_wk_pad, i32p), _p(self.n_dk_act, i32p),
                                                  _p(self.n_k, i32p), self.seed, self.sweep, self.doc_base)
            if rc:
                raise RuntimeError("oracle_llda_exact_sweep failed (%d)" % rc)
            self.sweep += 1

    def snapshot_sweep(self, n=1, n_refresh=1, tile_docs=256, n_threads=1):
        """Tiles are `tile_docs` consecutive GLOBAL document ids; tile i belongs to block i % n_refresh."""
        first_tile = self.doc_base // tile_docs
        n_tiles = (self.doc_base + self.D + tile_docs - 1) // tile_docs - first_tile
        # order tiles so that `i % n_blocks == b` in the C code selects block (tile_base + i) % n_refresh:
        # the C routine takes an explicit tile list, so pass tiles grouped by block, round-robin interleaved.
        tiles_by_block = [[] for _ in range(n_refresh)]
        for i in range(n_tiles):
            tiles_by_block[(first_tile + i) % n_refresh].append(i)
        width = max([len(t) for t in tiles_by_block] + [1])
        order = []
        for r in range(width):
            for b in range(n_refresh):
                order.append(tiles_by_block[b][r] if r < len(tiles_by_block[b]) else -1)
        tile_rng = np.zeros((len(order), 2), dtype=np.int64)
        for q, i in enumerate(order):
            if i >= 0:
                tile_rng[q, 0] = max(0, (first_tile + i) * tile_docs - self.doc_base)
                tile_rng[q, 1] = min(self.D, (first_tile + i + 1) * tile_docs - self.doc_base)
        for _ in range(n):
            rc = self.lib.oracle_llda_snapshot_sweep(len(order), _p(tile_rng, i64p), n_refresh, _p(self.doc_ptr, i64p),
                                                     _p(self.word, i32p), _p(self.freq, i32p), _p(self.z, i32p),
                                                     _p(self.lab_ptr, i64p), _p(self.lab_idx, i32p), self.K, self.V,
                                                     self.ldk, self.alpha, self.beta, _p(self.n_wk_pad, i32p),
                                                     _p(self.n_dk_act, i32p), _p(self.n_k, i32p), self.seed, self.sweep,
                                                     self.doc_base, n_threads)
            if rc:
                raise RuntimeError("oracle_llda_snapshot_sweep failed (%d)" % rc)
            self.sweep += 1

    # -- outputs, restating LabeledLDA.py:231-239 / CascadeLDA.py:394-395 on the transposed layout
    def phi(self, smoothed=True):
        n_kv = self.n_wk.T.astype(np.float64)
        if smoothed:
            return (n_kv + self.beta) / (self.n_k[:, None].astype(np.float64) + self.V * self.beta)
        with np.errstate(invalid="ignore", divide="ignore"):
            return n_kv / n_kv.sum(axis=1, keepdims=True)

    def n_dk_dense(self):
        out = np.zeros((self.D, self.K), dtype=np.int64)
        rows = np.repeat(np.arange(self.D), np.diff(self.lab_ptr))
        out[rows, self.lab_idx] = self.n_dk_act
        return out

    def theta(self):
        labs = np.zeros((self.D, self.K))
        rows = np.repeat(np.arange(self.D), np.diff(self.lab_ptr))
        labs[rows, self.lab_idx] = 1.0
        num = self.n_dk_dense() + labs * self.alpha
        return num / num.sum(axis=1)[:, None]


def test_chains(phi_KV, doc_ptr, word, freq, it, thinning, alpha, lab_ptr=None, lab_idx=None, z_init=None, init="llda",
                beta_fb=0.0, seed=0, chain_base=0):
    """CPU restatement of gibbs_test_run -> (th_hat, z)."""
    lib = load()
    phi_KV = _a(phi_KV, np.float64)
    K, V = phi_KV.shape
    doc_ptr, word = _a(doc_ptr, np.int64), _a(word, np.int32)
    freq = None if freq is None else _a(freq, np.int32)
    n = doc_ptr.shape[0] - 1
    if lab_ptr is not None:
        lab_ptr, lab_idx = _a(lab_ptr, np.int64), _a(lab_idx, np.int32)
        th = np.zeros(int(lab_ptr[-1]), dtype=np.float64)
    else:
        th = np.zeros((n, K), dtype=np.float64)
    mode = {"given": 0, "llda": 1, "cascade": 2}[init]
    z = _a(z_init, np.int32).copy() if mode == 0 else np.zeros(int(doc_ptr[-1]), dtype=np.int32)
    rc = lib.oracle_test_chains(K, V, _p(phi_KV, f64p), float(alpha), float(beta_fb), n, _p(doc_ptr, i64p), _p(word, i32p),
                                _p(freq, i32p), _p(lab_ptr, i64p), _p(lab_idx, i32p), _p(z, i32p), mode, int(it),
                                int(thinning), int(seed), int(chain_base), _p(th, f64p))
    if rc:
        raise RuntimeError("oracle_test_chains failed (%d)" % rc)
    return th, z
test_chains.__test__ = False


def perplexity(o):
    """LabeledLDA.py:256-265 from an LldaOracle's counts (not weighted by f, as in the reference)."""
    phi, theta = o.phi(), o.theta()
    doc_of = np.repeat(np.arange(o.D), np.diff(o.doc_ptr))
    dots = np.einsum("kn,nk->n", phi[:, o.word], theta[doc_of])
    return float(np.exp(-np.log(dots).sum() / o.N))
