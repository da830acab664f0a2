"""Small seeded corpora + comparison helpers shared by the parity tests."""
import numpy as np


def make_corpus(seed, D, K, V, label_lens, mean_pairs=12, with_freq=True, empty_docs=(), one_draw_docs=(),
                max_f=4):
    """label_lens: per-document label-list length (root included), cycled over the D documents."""
    rng = np.random.default_rng(seed)
    doc_ptr, lab_ptr = [0], [0]
    words, freqs, labs = [], [], []
    for d in range(D):
        if d in empty_docs:
            n = 0
        elif d in one_draw_docs:
            n = 1
        else:
            n = int(min(V, max(1, rng.poisson(mean_pairs))))
        w = np.sort(rng.choice(V, size=n, replace=False))
        words.append(w)
        freqs.append(rng.integers(1, max_f + 1, size=n))
        a = int(min(K, label_lens[d % len(label_lens)]))
        lab = np.concatenate(([0], 1 + np.sort(rng.choice(K - 1, size=a - 1, replace=False)))) if a > 1 else np.array([0])
        labs.append(lab)
        doc_ptr.append(doc_ptr[-1] + n)
        lab_ptr.append(lab_ptr[-1] + a)
    c = dict(D=D, K=K, V=V, doc_ptr=np.array(doc_ptr, dtype=np.int64),
             word=np.concatenate(words).astype(np.int32) if words else np.zeros(0, np.int32),
             freq=np.concatenate(freqs).astype(np.int32) if with_freq else None,
             lab_ptr=np.array(lab_ptr, dtype=np.int64), lab_idx=np.concatenate(labs).astype(np.int32))
    return c


def assert_state_equal(st, o, what=""):
    assert np.array_equal(st["z"], o.z), what + ": z differs"
    assert np.array_equal(st["n_wk"], o.n_wk), what + ": n_wk differs"
    assert np.array_equal(st["n_dk_act"], o.n_dk_act), what + ": n_dk differs"
    assert np.array_equal(st["n_k"], o.n_k), what + ": n_k differs"


def assert_invariants(st, c, extra_counts=0):
    """count conservation and mask respect."""
    f = c["freq"] if c["freq"] is not None else np.ones(c["word"].shape[0], dtype=np.int64)
    tot = int(np.sum(f))
    assert int(st["n_k"].sum()) == tot
    assert int(st["n_dk_act"].sum()) == tot
    assert int(st["n_wk"].sum()) == tot + extra_counts
    if extra_counts == 0:
        assert np.array_equal(st["n_wk"].sum(axis=0), st["n_k"])
    doc_of = np.repeat(np.arange(c["D"]), np.diff(c["doc_ptr"]))
    per_doc = np.bincount(doc_of, weights=f, minlength=c["D"]).astype(np.int64)
    lab_doc = np.repeat(np.arange(c["D"]), np.diff(c["lab_ptr"]))
    assert np.array_equal(np.bincount(lab_doc, weights=st["n_dk_act"], minlength=c["D"]).astype(np.int64), per_doc)
    # every z inside its document's label list
    D, K = c["D"], c["K"]
    mask = np.zeros((D, K), dtype=bool)
    mask[lab_doc, c["lab_idx"]] = True
    assert mask[doc_of, st["z"]].all()
