"""ctypes binding of libgibbs_b200.so (include/gibbs_b200.h).

The library is the product: there is no Python/NumPy implementation of the sweep in this package.
If the shared object is missing or no CUDA device is present, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgibbs_b200.so")

KIND_LLDA, KIND_HSLDA = 0, 1
MODE_EXACT, MODE_SNAPSHOT = 0, 1
MODES = {"exact": MODE_EXACT, "snapshot": MODE_SNAPSHOT}
FETCH = {"auto": 0, "dense": 1, "gather": 2}
FETCH_NAME = {1: "dense", 2: "gather"}
COMM_ID_BYTES = 128


class GibbsDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mode", C.c_int32), ("D", C.c_int64), ("V", C.c_int32), ("K", C.c_int32),
                ("alpha", C.c_double), ("beta", C.c_double), ("seed", C.c_uint64), ("device", C.c_int32),
                ("n_refresh", C.c_int32), ("doc_base", C.c_int64), ("reserved64", C.c_int64),
                ("tile_docs", C.c_int32), ("row_fetch", C.c_int32)]


class GibbsStats(C.Structure):
    _fields_ = [("draws", C.c_int64), ("sweeps", C.c_int64), ("last_sweep_ms", C.c_double),
                ("last_merge_ms", C.c_double), ("last_call_ms", C.c_double), ("last_launches", C.c_int64),
                ("bytes_per_draw", C.c_double), ("bytes_per_draw_dense", C.c_double),
                ("bytes_per_draw_gather", C.c_double), ("row_fetch", C.c_int32), ("reserved", C.c_int32),
                ("ldk", C.c_int32), ("max_active", C.c_int32), ("changed", C.c_int64), ("device_bytes", C.c_int64)]


_lib = None


def _p(tp):
    return C.POINTER(tp)


def load_library():
    """Load libgibbs_b200.so and declare every prototype of include/gibbs_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libgibbs_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C lda_thesis_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    lib.gibbs_last_error.restype = C.c_char_p
    lib.gibbs_version.restype = C.c_char_p
    lib.gibbs_device_count.restype = C.c_int
    lib.gibbs_create.argtypes = [_p(vp), _p(GibbsDesc)]
    lib.gibbs_destroy.argtypes = [vp]
    lib.gibbs_destroy.restype = None
    lib.gibbs_load.argtypes = [vp, _p(i64), _p(i32), _p(i32), _p(i32), _p(i64), _p(i32), _p(i32)]
    lib.gibbs_sweep.argtypes = [vp, i32]
    lib.gibbs_comm_unique_id.argtypes = [C.c_char_p]
    lib.gibbs_comm_init.argtypes = [vp, i32, i32, C.c_char_p]
    lib.gibbs_emit_theta_csr.argtypes = [vp, _p(dbl), i32]
    lib.gibbs_trim.argtypes = [vp]
    lib.gibbs_stream.argtypes = [vp, _p(vp)]
    lib.gibbs_get_state.argtypes = [vp, _p(i32), _p(i32), _p(i32), _p(i32)]
    lib.gibbs_set_z.argtypes = [vp, _p(i32)]
    lib.gibbs_add_counts.argtypes = [vp, i64, _p(i32), _p(i32), _p(i32)]
    lib.gibbs_emit_phi.argtypes = [vp, _p(dbl), i32]
    lib.gibbs_emit_theta.argtypes = [vp, _p(dbl), i32]
    lib.gibbs_stats.argtypes = [vp, _p(GibbsStats)]
    lib.gibbs_set_sweep_counter.argtypes = [vp, C.c_uint32]
    lib.gibbs_hslda_set.argtypes = [vp, i32, _p(dbl), _p(dbl), _p(dbl), _p(dbl)]
    lib.gibbs_thin_accumulate.argtypes = [vp, dbl, dbl, i32, i32]
    lib.gibbs_thin_get.argtypes = [vp, _p(dbl), _p(dbl)]
    lib.gibbs_perplexity.argtypes = [vp, _p(dbl), _p(i64)]
    lib.gibbs_test_create.argtypes = [_p(vp), i32, i32, i32, _p(dbl)]
    lib.gibbs_test_destroy.argtypes = [vp]
    lib.gibbs_test_destroy.restype = None
    lib.gibbs_test_run.argtypes = [vp, dbl, dbl, i64, _p(i64), _p(i32), _p(i32), _p(i64), _p(i32), _p(i32), i32, i32, i32,
                                   u64, i64, _p(dbl)]
    lib.gibbs_philox_kat.argtypes = [i32, i32, _p(C.c_uint32), _p(C.c_uint32), _p(C.c_uint32)]
    for name in ("gibbs_create", "gibbs_load", "gibbs_sweep", "gibbs_comm_unique_id", "gibbs_comm_init",
                 "gibbs_emit_theta_csr", "gibbs_trim", "gibbs_stream", "gibbs_get_state", "gibbs_set_z", "gibbs_add_counts", "gibbs_emit_phi",
                 "gibbs_emit_theta", "gibbs_stats", "gibbs_set_sweep_counter", "gibbs_hslda_set",
                 "gibbs_thin_accumulate", "gibbs_thin_get", "gibbs_perplexity", "gibbs_test_create", "gibbs_test_run",
                 "gibbs_philox_kat"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = _lib.gibbs_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def flatten_docs(docs, freqs=None):
    """list of per-document id lists -> (doc_ptr int64[D+1], word int32[N], freq int32[N])."""
    lens = np.fromiter((len(d) for d in docs), dtype=np.int64, count=len(docs))
    doc_ptr = np.zeros(len(docs) + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    n = int(doc_ptr[-1])
    word = np.fromiter((w for d in docs for w in d), dtype=np.int32, count=n)
    if freqs is None:
        freq = np.ones(n, dtype=np.int32)
    else:
        freq = np.fromiter((f for d in freqs for f in d), dtype=np.int32, count=n)
    return doc_ptr, word, freq


def labels_to_csr(labs):
    """Dense 0/1 matrix [D, K] (LabeledLDA.py:63) -> (lab_ptr int64[D+1], lab_idx int32[nnz]) ascending ids."""
    labs = np.asarray(labs)
    rows, cols = np.nonzero(labs)
    lab_ptr = np.zeros(labs.shape[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=labs.shape[0]), out=lab_ptr[1:])
    return lab_ptr, cols.astype(np.int32)


class GibbsSampler(object):
    """One device-resident corpus shard + its count tables.  Thin object wrapper over the C-ABI."""

    def __init__(self, D, V, K, alpha, beta, seed=0, mode="snapshot", kind=KIND_LLDA, device=0, n_refresh=1,
                 doc_base=0, tile_docs=0, row_fetch="auto"):
        lib = load_library()
        self._lib = lib
        self._h = C.c_void_p()
        self.D, self.V, self.K = int(D), int(V), int(K)
        self.mode = mode
        self.n_refresh = 1 if mode == "exact" else max(1, int(n_refresh))
        desc = GibbsDesc(kind=kind, mode=MODES[mode], D=self.D, V=self.V, K=self.K, alpha=float(alpha),
                         beta=float(beta), seed=int(seed) & 0xFFFFFFFFFFFFFFFF, device=int(device),
                         n_refresh=self.n_refresh, doc_base=int(doc_base), reserved64=0,
                         tile_docs=int(tile_docs), row_fetch=FETCH[row_fetch])
        _check(lib.gibbs_create(C.byref(self._h), C.byref(desc)), "gibbs_create")
        self.N = 0
        self.n_lab = 0
        self._keep = []

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gibbs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- corpus
    def load(self, doc_ptr, word, freq, z_init, lab_ptr, lab_idx, seg=None):
        doc_ptr = _arr(doc_ptr, np.int64)
        word = _arr(word, np.int32)
        freq = None if freq is None else _arr(freq, np.int32)
        z_init = None if z_init is None else _arr(z_init, np.int32)
        lab_ptr = _arr(lab_ptr, np.int64)
        lab_idx = _arr(lab_idx, np.int32)
        seg = None if seg is None else _arr(seg, np.int32)
        if doc_ptr.shape[0] != self.D + 1 or lab_ptr.shape[0] != self.D + 1:
            raise ValueError("doc_ptr / lab_ptr must have D+1 entries")
        self.N = int(doc_ptr[-1])
        self.n_lab = int(lab_ptr[-1])
        if word.shape[0] != self.N or (freq is not None and freq.shape[0] != self.N) or \
                (z_init is not None and z_init.shape[0] != self.N) or lab_idx.shape[0] != self.n_lab:
            raise ValueError("array lengths do not match the CSR offsets")
        _check(self._lib.gibbs_load(self._h, _ptr(doc_ptr, C.c_int64), _ptr(word, C.c_int32), _ptr(freq, C.c_int32),
                                    _ptr(z_init, C.c_int32), _ptr(lab_ptr, C.c_int64), _ptr(lab_idx, C.c_int32),
                                    _ptr(seg, C.c_int32)), "gibbs_load")

    # -- sweeps
    def sweep(self, n=1):
        _check(self._lib.gibbs_sweep(self._h, int(n)), "gibbs_sweep")

    def comm_init(self, nranks, rank, uid):
        """Join the NCCL communicator (before load).  uid: the 128 bytes of comm_unique_id() from rank 0."""
        uid = bytes(uid)
        if len(uid) != COMM_ID_BYTES:
            raise ValueError("uid must be %d bytes" % COMM_ID_BYTES)
        _check(self._lib.gibbs_comm_init(self._h, int(nranks), int(rank), uid), "gibbs_comm_init")

    def row_fetch(self):
        return FETCH_NAME[self.stats()["row_fetch"]]

    def trim(self):
        _check(self._lib.gibbs_trim(self._h), "gibbs_trim")

    def stream(self):
        s = C.c_void_p()
        _check(self._lib.gibbs_stream(self._h, C.byref(s)), "gibbs_stream")
        return s.value or 0

    def set_sweep_counter(self, sweep):
        _check(self._lib.gibbs_set_sweep_counter(self._h, int(sweep)), "gibbs_set_sweep_counter")

    # -- state
    def get_state(self, z=True, n_wk=True, n_dk_act=True, n_k=True, n_dk_len=None):
        out = {}
        zb = np.empty(self.N, dtype=np.int32) if z else None
        wb = np.empty((self.V, self.K), dtype=np.int32) if n_wk else None
        db = np.empty(self.n_lab if n_dk_len is None else n_dk_len, dtype=np.int32) if n_dk_act else None
        kb = np.empty(self.K, dtype=np.int32) if n_k else None
        _check(self._lib.gibbs_get_state(self._h, _ptr(zb, C.c_int32), _ptr(wb, C.c_int32), _ptr(db, C.c_int32),
                                         _ptr(kb, C.c_int32)), "gibbs_get_state")
        out["z"], out["n_wk"], out["n_dk_act"], out["n_k"] = zb, wb, db, kb
        return out

    def alloc_state_buffers(self, pinned=False, n_wk=True):
        """Host buffers for get_state_into; pinned=True page-locks them (torch is only the allocator)."""
        shapes = {"z": (self.N,), "n_dk_act": (self.n_lab,), "n_k": (self.K,)}
        if n_wk:
            shapes["n_wk"] = (self.V, self.K)
        out = {}
        if pinned:
            import torch
            self._pinned = getattr(self, "_pinned", [])
            for k, shp in shapes.items():
                t = torch.empty(shp, dtype=torch.int32).pin_memory()
                self._pinned.append(t)
                out[k] = t.numpy()
        else:
            for k, shp in shapes.items():
                out[k] = np.empty(shp, dtype=np.int32)
        return out

    def get_state_into(self, bufs):
        g = lambda k: _ptr(bufs[k], C.c_int32) if k in bufs else None
        _check(self._lib.gibbs_get_state(self._h, g("z"), g("n_wk"), g("n_dk_act"), g("n_k")), "gibbs_get_state")
        return bufs

    def set_z(self, z):
        z = _arr(z, np.int32)
        if z.shape[0] != self.N:
            raise ValueError("z must have one entry per draw")
        _check(self._lib.gibbs_set_z(self._h, _ptr(z, C.c_int32)), "gibbs_set_z")

    def add_counts(self, word, topic, count):
        word, topic, count = _arr(word, np.int32), _arr(topic, np.int32), _arr(count, np.int32)
        if not (word.shape == topic.shape == count.shape):
            raise ValueError("word, topic, count must have the same length")
        _check(self._lib.gibbs_add_counts(self._h, word.shape[0], _ptr(word, C.c_int32), _ptr(topic, C.c_int32),
                                          _ptr(count, C.c_int32)), "gibbs_add_counts")

    def emit_phi(self, smoothed=True):
        out = np.empty((self.K, self.V), dtype=np.float64)
        _check(self._lib.gibbs_emit_phi(self._h, _ptr(out, C.c_double), 1 if smoothed else 0), "gibbs_emit_phi")
        return out

    def emit_theta_csr(self, smoothed=True):
        out = np.empty(self.n_lab, dtype=np.float64)
        _check(self._lib.gibbs_emit_theta_csr(self._h, _ptr(out, C.c_double), 1 if smoothed else 0),
               "gibbs_emit_theta_csr")
        return out

    def emit_theta(self, smoothed=True):
        out = np.empty((self.D, self.K), dtype=np.float64)
        _check(self._lib.gibbs_emit_theta(self._h, _ptr(out, C.c_double), 1 if smoothed else 0), "gibbs_emit_theta")
        return out

    def thin_accumulate(self, c_old, c_new, smoothed=True, phi=True, theta=True):
        """hat = c_old * hat + c_new * current on the device (first call after load: hat = current)."""
        _check(self._lib.gibbs_thin_accumulate(self._h, float(c_old), float(c_new), 1 if smoothed else 0,
                                               (1 if phi else 0) | (2 if theta else 0)), "gibbs_thin_accumulate")

    def thin_get(self, phi=True, theta=True):
        ph = np.empty((self.K, self.V), dtype=np.float64) if phi else None
        th = np.empty(self.n_lab, dtype=np.float64) if theta else None
        _check(self._lib.gibbs_thin_get(self._h, _ptr(ph, C.c_double), _ptr(th, C.c_double)), "gibbs_thin_get")
        return ph, th

    def perplexity(self):
        """LabeledLDA.py:256-265 on the live counts."""
        s, n = C.c_double(), C.c_int64()
        _check(self._lib.gibbs_perplexity(self._h, C.byref(s), C.byref(n)), "gibbs_perplexity")
        return float(np.exp(s.value / n.value)) if n.value else float("nan")

    def stats(self):
        st = GibbsStats()
        _check(self._lib.gibbs_stats(self._h, C.byref(st)), "gibbs_stats")
        return {k: getattr(st, k) for k, _ in GibbsStats._fields_}

    def hslda_set(self, eta, a_act, mean_a_act, alpha_beta):
        eta = _arr(eta, np.float64)
        a_act = _arr(a_act, np.float64)
        mean_a_act = _arr(mean_a_act, np.float64)
        alpha_beta = _arr(alpha_beta, np.float64)
        if eta.ndim != 2 or eta.shape[1] != self.K or alpha_beta.shape[0] != self.K:
            raise ValueError("eta must be [L, K], alpha_beta [K]")
        if a_act.shape[0] != self.n_lab or mean_a_act.shape[0] != self.n_lab:
            raise ValueError("a_act / mean_a_act must align with lab_idx")
        _check(self._lib.gibbs_hslda_set(self._h, eta.shape[0], _ptr(eta, C.c_double), _ptr(a_act, C.c_double),
                                         _ptr(mean_a_act, C.c_double), _ptr(alpha_beta, C.c_double)),
               "gibbs_hslda_set")


TEST_INIT = {"given": 0, "llda": 1, "cascade": 2}


class TestChains(object):
    """Frozen-phi test chains (LabeledLDA.py:155-212, CascadeLDA.py:186-247): a device copy of phi + gibbs_test_run."""
    __test__ = False          # not a pytest class

    def __init__(self, phi_KV, device=0):
        lib = load_library()
        self._lib = lib
        phi_KV = _arr(phi_KV, np.float64)
        if phi_KV.ndim != 2:
            raise ValueError("phi must be [K, V]")
        self.K, self.V = int(phi_KV.shape[0]), int(phi_KV.shape[1])
        self._t = C.c_void_p()
        _check(lib.gibbs_test_create(C.byref(self._t), int(device), self.K, self.V, _ptr(phi_KV, C.c_double)),
               "gibbs_test_create")

    def close(self):
        if getattr(self, "_t", None) is not None and self._t:
            self._lib.gibbs_test_destroy(self._t)
            self._t = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, doc_ptr, word, freq, it, thinning, alpha, lab_ptr=None, lab_idx=None, z_init=None, init="llda",
            beta_fb=0.0, seed=0, chain_base=0):
        """-> (th_hat, z).  th_hat is [n_chains, K] without topic lists, else aligned with lab_idx."""
        doc_ptr = _arr(doc_ptr, np.int64)
        word = _arr(word, np.int32)
        freq = None if freq is None else _arr(freq, np.int32)
        n = doc_ptr.shape[0] - 1
        N = int(doc_ptr[-1]) if n >= 0 else 0
        if word.shape[0] != N or (freq is not None and freq.shape[0] != N):
            raise ValueError("array lengths do not match the CSR offsets")
        if lab_ptr is not None:
            lab_ptr, lab_idx = _arr(lab_ptr, np.int64), _arr(lab_idx, np.int32)
            if lab_ptr.shape[0] != n + 1 or lab_idx.shape[0] != int(lab_ptr[-1]):
                raise ValueError("lab_ptr / lab_idx do not match")
            th = np.zeros(int(lab_ptr[-1]), dtype=np.float64)
        else:
            th = np.zeros((n, self.K), dtype=np.float64)
        if init == "given":
            z = _arr(z_init, np.int32).copy()
            if z.shape[0] != N:
                raise ValueError("z_init must have one entry per draw")
        else:
            z = np.zeros(N, dtype=np.int32)
        _check(self._lib.gibbs_test_run(self._t, float(alpha), float(beta_fb), n, _ptr(doc_ptr, C.c_int64),
                                        _ptr(word, C.c_int32), _ptr(freq, C.c_int32), _ptr(lab_ptr, C.c_int64),
                                        _ptr(lab_idx, C.c_int32), _ptr(z, C.c_int32), TEST_INIT[init], int(it),
                                        int(thinning), int(seed) & 0xFFFFFFFFFFFFFFFF, int(chain_base),
                                        _ptr(th, C.c_double)), "gibbs_test_run")
        return th, z


def philox_kat(ctr4, key2, device=0):
    lib = load_library()
    ctr4 = _arr(ctr4, np.uint32).reshape(-1, 4)
    key2 = _arr(key2, np.uint32).reshape(-1, 2)
    out = np.empty_like(ctr4)
    _check(lib.gibbs_philox_kat(int(device), ctr4.shape[0], _ptr(ctr4, C.c_uint32), _ptr(key2, C.c_uint32),
                                _ptr(out, C.c_uint32)), "gibbs_philox_kat")
    return out


def comm_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib.gibbs_comm_unique_id(buf), "gibbs_comm_unique_id")
    return buf.raw


def device_count():
    return int(load_library().gibbs_device_count())
