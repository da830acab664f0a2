"""Synthetic corpora in pair space for the BASELINE.json configs.

No strings, no gensim: documents are generated directly as the (word id, frequency) pairs that
`Dictionary.doc2bow` would produce (LabeledLDA.py:64), in the CSR layout of include/gibbs_b200.h.

    word ids   ~ Zipf(s = 1.07) over V, unique inside a document, ascending (doc2bow order)
    pairs/doc  ~ Poisson(mean_pairs) thinned from an oversampled Zipf draw (mean within 1 % of mean_pairs)
    f          = Geometric(0.8) on {1, 2, ...}  (mean 1.25; abstracts: 280 645 / 225 152)
    labels/doc = root (topic 0, LabeledLDA.py:96) + clip(1 + Poisson(2.5), 1, 7) distinct non-root topics
"""
import numpy as np

CONFIGS = {
    # name: D, mean pairs per doc, K, V, seed        
    "C2": dict(D=100_000, mean_pairs=200, K=100, V=100_000, seed=20260201),
    "C4": dict(D=1_000_000, mean_pairs=250, K=500, V=100_000, seed=20260401),
    "C4shard": dict(D=125_000, mean_pairs=250, K=500, V=100_000, seed=20260401),   # one rank's share of C4 at 8 GPUs
}


def zipf_cdf(V, s=1.07):
    w = np.arange(1, V + 1, dtype=np.float64) ** (-s)
    c = np.cumsum(w)
    return c / c[-1]


def _pairs_block(rng, cdf, n_docs, mean_pairs, oversample=2.2):
    """(doc, word) unique pairs for n_docs documents, sorted by (doc, word); returns doc ids, word ids."""
    V = cdf.shape[0]
    want = np.maximum(rng.poisson(mean_pairs, size=n_docs).astype(np.int64), 1)
    raw = np.ceil(want * oversample).astype(np.int64) + 8
    doc = np.repeat(np.arange(n_docs, dtype=np.int64), raw)
    word = np.searchsorted(cdf, rng.random(doc.shape[0]), side="right").astype(np.int64)
    np.minimum(word, V - 1, out=word)
    key = doc * V + word
    key.sort()
    key = key[np.concatenate(([True], key[1:] != key[:-1]))]      # distinct ids per document
    doc = key // V
    # thin each document to ~want[d] of its distinct ids (Bernoulli, keeps the ascending order)
    have = np.bincount(doc, minlength=n_docs)
    keep = rng.random(key.shape[0]) * have[doc] < want[doc]
    key = key[keep]
    return key // V, key % V


def labeled_corpus(D, mean_pairs, K, V, seed, block_docs=20_000, with_freq=True, max_extra_labels=7):
    """Returns dict(doc_ptr, word, freq, lab_ptr, lab_idx, D, K, V)."""
    rng = np.random.default_rng(seed)
    cdf = zipf_cdf(V)
    words, counts = [], []
    for d0 in range(0, D, block_docs):
        n = min(block_docs, D - d0)
        doc, word = _pairs_block(rng, cdf, n, mean_pairs)
        words.append(word.astype(np.int32))
        counts.append(np.bincount(doc, minlength=n))
    word = np.concatenate(words) if words else np.zeros(0, np.int32)
    lens = np.concatenate(counts) if counts else np.zeros(0, np.int64)
    doc_ptr = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    freq = rng.geometric(0.8, size=word.shape[0]).astype(np.int32) if with_freq else None
    # labels: root + 1..max_extra distinct non-root topics
    n_extra = np.clip(1 + rng.poisson(2.5, size=D), 1, min(max_extra_labels, K - 1)).astype(np.int64)
    cand = rng.integers(1, K, size=(D, max_extra_labels), dtype=np.int64)
    cand[np.arange(max_extra_labels)[None, :] >= n_extra[:, None]] = 0           # unused slots -> root (dedupes away)
    full = np.concatenate([np.zeros((D, 1), dtype=np.int64), cand], axis=1)
    full.sort(axis=1)
    first = np.ones_like(full, dtype=bool)
    first[:, 1:] = full[:, 1:] != full[:, :-1]
    lab_ptr = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(first.sum(axis=1), out=lab_ptr[1:])
    lab_idx = full[first].astype(np.int32)
    return dict(doc_ptr=doc_ptr, word=word, freq=freq, lab_ptr=lab_ptr, lab_idx=lab_idx, D=D, K=K, V=V)


def config_corpus(name, D=None):
    cfg = dict(CONFIGS[name])
    if D is not None:
        cfg["D"] = int(D)
    return labeled_corpus(**cfg)


def shard_bounds(doc_ptr, n_shards):
    """Contiguous document ranges balanced by the number of draws, aligned to nothing."""
    doc_ptr = np.asarray(doc_ptr)
    N = int(doc_ptr[-1])
    cuts = [0]
    for r in range(1, n_shards):
        cuts.append(int(np.searchsorted(doc_ptr, N * r // n_shards, side="left")))
    cuts.append(doc_ptr.shape[0] - 1)
    return cuts


def take_docs(c, d0, d1):
    """Sub-corpus of documents [d0, d1) re-based to start at 0."""
    p0, p1 = int(c["doc_ptr"][d0]), int(c["doc_ptr"][d1])
    l0, l1 = int(c["lab_ptr"][d0]), int(c["lab_ptr"][d1])
    out = dict(c)
    out.update(D=d1 - d0, doc_ptr=c["doc_ptr"][d0:d1 + 1] - p0, word=c["word"][p0:p1],
               freq=None if c["freq"] is None else c["freq"][p0:p1],
               lab_ptr=c["lab_ptr"][d0:d1 + 1] - l0, lab_idx=c["lab_idx"][l0:l1])
    return out
