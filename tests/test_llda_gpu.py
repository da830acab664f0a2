"""Device path vs CPU oracle, through the C-ABI (lda_thesis_b200/_lib.py -> libgibbs_b200.so).

Everything here needs a B200: `pytest -m gpu`.  The comparisons are device vs oracle/liboracle.so and
device vs the golden fixtures the UNMODIFIED reference produced (tests/golden, oracle/make_golden.py);
never device vs device.  Integer state is compared bit-for-bit; phi/theta with rtol 1e-12 (they are
pure fp64 functions of the integer counts; the task's tolerance is 1e-5 rel-inf).

Parity statement (SURVEY.md §7 hard part 1): `exact` mode reproduces the reference's corpus-sequential
chain (LabeledLDA.py:101-125) given identical draws; `snapshot` mode is a different (document-parallel)
schedule and is compared with the oracle's restatement of the SAME schedule.
"""
import numpy as np
import pytest

from conftest import load_golden
from helpers import assert_invariants, assert_state_equal, make_corpus

pytestmark = pytest.mark.gpu

ALPHA, BETA = 0.1, 0.01


def _pair(gibbs, oracle, c, mode, seed=3, z=None, n_refresh=1, tile_docs=0, seg=None, alpha=ALPHA, beta=BETA,
          row_fetch="auto"):
    o = oracle.LldaOracle(c["doc_ptr"], c["word"], c["freq"], c["lab_ptr"], c["lab_idx"], c["K"], c["V"],
                          alpha, beta, seed=seed, z=z)
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], alpha, beta, seed=seed, mode=mode, n_refresh=n_refresh,
                           tile_docs=tile_docs, row_fetch=row_fetch)
    g.load(c["doc_ptr"], c["word"], c["freq"], z, c["lab_ptr"], c["lab_idx"], seg=seg)
    return g, o


def test_device_init_matches_oracle(gibbs, oracle):
    c = make_corpus(1, D=300, K=37, V=500, label_lens=[1, 2, 5, 8, 9, 17, 33])
    g, o = _pair(gibbs, oracle, c, "snapshot")
    st = g.get_state()
    assert_state_equal(st, o, "init")
    assert_invariants(st, c)
    g.close()


@pytest.mark.parametrize("name", ["llda_abstracts300.npz", "sublda_abstracts300.npz"])
def test_exact_matches_unmodified_reference(gibbs, oracle, name):
    """Golden: reference training_iteration() with only multinom_draw patched (oracle/make_golden.py)."""
    gold = load_golden(name)
    c = dict(D=gold["doc_ptr"].shape[0] - 1, K=int(gold["K"]), V=int(gold["V"]), doc_ptr=gold["doc_ptr"],
             word=gold["word"], freq=gold["freq"], lab_ptr=gold["lab_ptr"], lab_idx=gold["lab_idx"])
    g, o = _pair(gibbs, oracle, c, "exact", seed=int(gold["seed"]), z=gold["z_init"],
                 alpha=float(gold["alpha"]), beta=float(gold["beta"]))
    extra = 0
    if "n_k_v_init" in gold:      # SubLDA's spurious initial counts (CascadeLDA.py:382-385)
        diff = gold["n_k_v_init"].T - o.n_wk
        w, k = np.nonzero(diff)
        o.add_word_topic_counts(w, k, diff[w, k])
        g.add_counts(w, k, diff[w, k])
        extra = int(diff.sum())
    rows = np.repeat(np.arange(c["D"]), np.diff(c["lab_ptr"]))
    for s in range(3):
        g.sweep(1)
        o.exact_sweep(1)
        st = g.get_state()
        assert_state_equal(st, o, "sweep %d vs oracle" % s)
        assert np.array_equal(st["z"], gold["s%d_z" % s]), "z differs from the reference after sweep %d" % s
        assert np.array_equal(st["n_wk"].T, gold["s%d_n_k_v" % s])
        assert np.array_equal(st["n_k"], gold["s%d_n_zk" % s])
        dense = np.zeros((c["D"], c["K"]), dtype=np.int64)
        dense[rows, c["lab_idx"]] = st["n_dk_act"]
        assert np.array_equal(dense, gold["s%d_n_d_k" % s])
        assert_invariants(st, c, extra)
    if "phi_sub" in gold:
        phi, theta = g.emit_phi(True), g.emit_theta(True)
        assert np.allclose(phi[:, gold["phi_cols"]], gold["phi_sub"], rtol=1e-12, atol=0)
        assert np.allclose(phi.sum(axis=1), gold["phi_rowsum"], rtol=1e-12, atol=0)
        assert np.allclose(theta, gold["theta"], rtol=1e-12, atol=0)
    else:
        ph = g.emit_phi(False)    # CascadeLDA.py:394-395: n_k_v / rowsum, unsmoothed
        assert np.allclose(ph, gold["ph"], rtol=1e-12, atol=0, equal_nan=True)
    g.close()


def test_exact_synthetic(gibbs, oracle):
    c = make_corpus(2, D=120, K=70, V=300, label_lens=[1, 3, 8, 31, 32, 33, 64, 65, 70], mean_pairs=9)
    g, o = _pair(gibbs, oracle, c, "exact")
    for s in range(3):
        g.sweep(1)
        o.exact_sweep(1)
        st = g.get_state()
        assert_state_equal(st, o, "exact sweep %d" % s)
        assert_invariants(st, c)
    assert np.allclose(g.emit_phi(), o.phi(), rtol=1e-12, atol=0)
    assert np.allclose(g.emit_theta(), o.theta(), rtol=1e-12, atol=0)
    g.close()


BOUNDARY_LENS = [1, 2, 7, 8, 9, 16, 17, 32, 33, 64, 65, 128, 129, 256, 257, 512]


@pytest.mark.parametrize("row_fetch", ["dense", "gather", "auto"])
@pytest.mark.parametrize("n_refresh,tile_docs,with_freq,K", [
    (1, 0, True, 600), (4, 8, True, 600), (3, 5, False, 517), (1, 0, True, 33), (2, 16, True, 100)])
def test_snapshot_matches_oracle(gibbs, oracle, n_refresh, tile_docs, with_freq, K, row_fetch):
    """Every label-list bin boundary, freq=NULL, K not a multiple of 32, 1-draw and empty documents."""
    lens = [a for a in BOUNDARY_LENS if a <= K]
    c = make_corpus(10 + n_refresh, D=7 * len(lens) + 5, K=K, V=400, label_lens=lens, mean_pairs=10,
                    with_freq=with_freq, empty_docs=(3, 40), one_draw_docs=(0, 17))
    g, o = _pair(gibbs, oracle, c, "snapshot", n_refresh=n_refresh, tile_docs=tile_docs, row_fetch=row_fetch)
    td = tile_docs or 256
    for s in range(4):
        g.sweep(1)
        o.snapshot_sweep(1, n_refresh=n_refresh, tile_docs=td)
        st = g.get_state()
        assert_state_equal(st, o, "snapshot sweep %d" % s)
        assert_invariants(st, c)
    assert g.stats()["sweeps"] == 4 and g.stats()["draws"] == 4 * g.N
    assert np.allclose(g.emit_phi(), o.phi(), rtol=1e-12, atol=0)
    theta = g.emit_theta()
    assert np.allclose(theta, o.theta(), rtol=1e-12, atol=0)
    rows = np.repeat(np.arange(c["D"]), np.diff(c["lab_ptr"]))
    assert np.array_equal(g.emit_theta_csr(), theta[rows, c["lab_idx"]])
    g.close()


def test_snapshot_long_documents(gibbs, oracle):
    """Documents much longer than a chunk / the prefetch rings, in every kernel variant."""
    for row_fetch in ("dense", "gather"):
        c = make_corpus(21, D=60, K=40, V=3000, label_lens=[1, 3, 4, 5, 8, 9, 16, 17, 32, 33], mean_pairs=300)
        g, o = _pair(gibbs, oracle, c, "snapshot", row_fetch=row_fetch)
        for s in range(2):
            g.sweep(1)
            o.snapshot_sweep(1)
            assert_state_equal(g.get_state(), o, "%s long docs sweep %d" % (row_fetch, s))
        g.close()


def test_reload_reuses_handle(gibbs, oracle):
    """gibbs_load on a live handle (the bench's e2e leg): state restarts from the host-side z."""
    c = make_corpus(22, D=150, K=30, V=200, label_lens=[2, 5, 7])
    g, o = _pair(gibbs, oracle, c, "snapshot")
    g.sweep(2)
    o.snapshot_sweep(2)
    z = g.get_state()["z"]
    g.load(c["doc_ptr"], c["word"], c["freq"], z, c["lab_ptr"], c["lab_idx"])
    assert_state_equal(g.get_state(), o, "after reload")
    g.set_sweep_counter(2)
    g.sweep(1)
    o.snapshot_sweep(1)
    assert_state_equal(g.get_state(), o, "sweep after reload")
    g.close()


def test_snapshot_multi_sweep_call_equals_single_calls(gibbs, oracle):
    c = make_corpus(5, D=200, K=64, V=300, label_lens=[2, 4, 6, 9])
    g, o = _pair(gibbs, oracle, c, "snapshot", n_refresh=2, tile_docs=16)
    g.sweep(5)
    o.snapshot_sweep(5, n_refresh=2, tile_docs=16)
    assert_state_equal(g.get_state(), o, "5 sweeps in one call")
    g.close()


def test_snapshot_topic_segments(gibbs, oracle):
    """seg != NULL: each document only stages its topic block (a CascadeLDA node's block)."""
    rng = np.random.default_rng(4)
    K, V, D = 96, 200, 150
    blocks = [(0, 12), (12, 20), (20, 52), (52, 96)]
    doc_ptr, lab_ptr, words, freqs, labs, seg = [0], [0], [], [], [], []
    for d in range(D):
        lo, hi = blocks[d % len(blocks)]
        n = int(max(1, rng.poisson(8)))
        words.append(np.sort(rng.choice(V, size=n, replace=False)))
        freqs.append(rng.integers(1, 4, size=n))
        a = int(rng.integers(1, hi - lo + 1))
        labs.append(np.concatenate(([lo], lo + 1 + np.sort(rng.choice(hi - lo - 1, size=a - 1, replace=False)))))
        doc_ptr.append(doc_ptr[-1] + n)
        lab_ptr.append(lab_ptr[-1] + a)
        seg += [lo, hi]
    c = dict(D=D, K=K, V=V, doc_ptr=np.array(doc_ptr), word=np.concatenate(words).astype(np.int32),
             freq=np.concatenate(freqs).astype(np.int32), lab_ptr=np.array(lab_ptr),
             lab_idx=np.concatenate(labs).astype(np.int32))
    g, o = _pair(gibbs, oracle, c, "snapshot", seg=np.array(seg, dtype=np.int32))
    for s in range(3):
        g.sweep(1)
        o.snapshot_sweep(1)
        assert_state_equal(g.get_state(), o, "seg sweep %d" % s)
    g.close()


def test_set_z_rebuilds_counts(gibbs, oracle):
    c = make_corpus(6, D=100, K=20, V=100, label_lens=[3, 5])
    g, o = _pair(gibbs, oracle, c, "snapshot")
    g.sweep(2)
    o.snapshot_sweep(2)
    z0 = oracle.LldaOracle(c["doc_ptr"], c["word"], c["freq"], c["lab_ptr"], c["lab_idx"], c["K"], c["V"],
                           ALPHA, BETA, seed=99)
    g.set_z(z0.z)
    assert_state_equal(g.get_state(), z0, "after set_z")
    g.close()


def test_error_paths(gibbs):
    c = make_corpus(7, D=20, K=10, V=50, label_lens=[2, 3])
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA)
    with pytest.raises(RuntimeError, match="not loaded"):
        g.sweep(1)
    bad_z = np.full(c["word"].shape[0], 9, dtype=np.int32)          # topic 9 is in (almost) no label list
    with pytest.raises(RuntimeError, match="label list"):
        g.load(c["doc_ptr"], c["word"], c["freq"], bad_z, c["lab_ptr"], c["lab_idx"])
    g.close()
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA)
    big_f = c["freq"].copy()
    big_f[0] = 70000
    with pytest.raises(RuntimeError, match="65535"):
        g.load(c["doc_ptr"], c["word"], big_f, None, c["lab_ptr"], c["lab_idx"])
    g.close()
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA)
    bad_w = c["word"].copy()
    bad_w[1] = c["V"]
    with pytest.raises(RuntimeError, match="out of range"):
        g.load(c["doc_ptr"], bad_w, c["freq"], None, c["lab_ptr"], c["lab_idx"])
    g.close()
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA)
    bad_lab = c["lab_idx"].copy()
    bad_lab[0] = c["K"]
    with pytest.raises(RuntimeError, match="lab_idx"):
        g.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], bad_lab)
    dup = c["lab_idx"].copy()
    dup[1] = dup[0]
    with pytest.raises(RuntimeError, match="ascending"):
        g.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], dup)
    g.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], c["lab_idx"])     # a failed load leaves it usable
    g.sweep(1)
    g.close()
    with pytest.raises(RuntimeError, match="not implemented"):
        gibbs.GibbsSampler(10, 10, 10, 1.0, 1.0, kind=gibbs.KIND_HSLDA)
    # > 512 active topics is outside the supported range
    K = 600
    lab_ptr = np.array([0, 513], dtype=np.int64)
    g = gibbs.GibbsSampler(1, 10, K, ALPHA, BETA)
    with pytest.raises(RuntimeError, match="active topics"):
        g.load(np.array([0, 1]), np.array([1], dtype=np.int32), None, None, lab_ptr,
               np.arange(513, dtype=np.int32))
    g.close()


def test_large_scale_properties(gibbs):
    """Size-independent properties at a size the oracle would take too long for (1M draws, K=500)."""
    from lda_thesis_b200 import synth
    c = synth.labeled_corpus(D=5000, mean_pairs=200, K=500, V=20000, seed=3)
    runs = []
    for n_refresh in (1, 4):
        g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA, seed=5, mode="snapshot", n_refresh=n_refresh)
        g.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], c["lab_idx"])
        g.sweep(3)
        st = g.get_state()
        assert_invariants(st, c)
        runs.append(st)
        # determinism: the same handle parameters give the same chain
        g2 = gibbs.GibbsSampler(c["D"], c["V"], c["K"], ALPHA, BETA, seed=5, mode="snapshot", n_refresh=n_refresh)
        g2.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], c["lab_idx"])
        g2.sweep(3)
        st2 = g2.get_state()
        for k in ("z", "n_wk", "n_dk_act", "n_k"):
            assert np.array_equal(st[k], st2[k])
        phi = g.emit_phi()
        assert np.allclose(phi.sum(axis=1), 1.0, rtol=1e-12)
        g.close()
        g2.close()
