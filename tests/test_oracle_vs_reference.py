"""oracle/gibbs_oracle.c pinned against the UNMODIFIED reference.

tests/golden/*.npz were produced by oracle/make_golden.py: the reference's own
LabeledLDA.training_iteration (LabeledLDA.py:101-125) / SubLDA.training_iteration
(CascadeLDA.py:397-421) with only `multinom_draw` swapped for the Philox inverse-CDF draw.
Integer state must match bit-for-bit; phi/theta to 1e-12 (they are functions of the counts).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden

N_SWEEPS = 3


def extra_counts(g, o):
    """COO of (reference initial n_k_v) - (histogram of z_init): SubLDA's spurious counts, empty for LabeledLDA."""
    if "n_k_v_init" not in g:
        return None
    diff = g["n_k_v_init"].T - o.n_wk
    w, k = np.nonzero(diff)
    return w.astype(np.int32), k.astype(np.int32), diff[w, k].astype(np.int32)


def _oracle_from(g, oracle):
    o = oracle.LldaOracle(g["doc_ptr"], g["word"], g["freq"], g["lab_ptr"], g["lab_idx"], int(g["K"]), int(g["V"]),
                          float(g["alpha"]), float(g["beta"]), seed=int(g["seed"]), z=g["z_init"])
    extra = extra_counts(g, o)
    if extra is not None:
        o.add_word_topic_counts(*extra)
    return o


@pytest.mark.parametrize("name", ["llda_abstracts300.npz", "sublda_abstracts300.npz"])
def test_exact_sweep_matches_reference(oracle, name):
    g = load_golden(name)
    o = _oracle_from(g, oracle)
    for s in range(N_SWEEPS):
        o.exact_sweep()
        assert np.array_equal(o.z, g["s%d_z" % s]), "z differs after sweep %d" % s
        assert np.array_equal(o.n_wk.T, g["s%d_n_k_v" % s])
        assert np.array_equal(o.n_dk_dense(), g["s%d_n_d_k" % s])
        assert np.array_equal(o.n_k, g["s%d_n_zk" % s])
        # invariants of SURVEY.md §4
        tot = int(g["freq"].sum())
        assert o.n_k.sum() == tot and o.n_dk_act.sum() == tot
        if "n_k_v_init" not in g:
            assert o.n_wk.sum() == tot and np.array_equal(o.n_wk.sum(axis=0), o.n_k)


def test_outputs_match_reference(oracle):
    g = load_golden("llda_abstracts300.npz")
    o = _oracle_from(g, oracle)
    o.exact_sweep(N_SWEEPS)
    phi, theta = o.phi(), o.theta()
    assert np.allclose(phi[:, g["phi_cols"]], g["phi_sub"], rtol=1e-12, atol=0)
    assert np.allclose(phi.sum(axis=1), g["phi_rowsum"], rtol=1e-12, atol=0)
    assert np.allclose(theta, g["theta"], rtol=1e-12, atol=0)
    s = load_golden("sublda_abstracts300.npz")
    o = _oracle_from(s, oracle)
    o.exact_sweep(N_SWEEPS)
    assert np.allclose(o.phi(smoothed=False), s["ph"], rtol=1e-12, atol=0, equal_nan=True)


def test_init_z_is_uniform_over_labels(oracle):
    g = load_golden("llda_abstracts300.npz")
    o = oracle.LldaOracle(g["doc_ptr"], g["word"], g["freq"], g["lab_ptr"], g["lab_idx"], int(g["K"]), int(g["V"]),
                          0.1, 0.01, seed=5)
    lab_of_draw = np.repeat(np.arange(o.D), np.diff(o.doc_ptr))
    for n in range(0, o.N, 97):
        d = lab_of_draw[n]
        assert o.z[n] in o.lab_idx[o.lab_ptr[d]:o.lab_ptr[d + 1]]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_golden_regenerates_byte_identically(tmp_path):
    import hashlib
    before = {f: hashlib.md5(open(os.path.join(GOLDEN, f), "rb").read()).hexdigest()
              for f in ("llda_abstracts300.npz", "sublda_abstracts300.npz")}
    env = dict(os.environ, LDA_GOLDEN_DIR=str(tmp_path))
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "make_golden.py")], env=env,
                          stdout=subprocess.DEVNULL)
    for f, h in before.items():
        assert hashlib.md5(open(os.path.join(str(tmp_path), f), "rb").read()).hexdigest() == h
