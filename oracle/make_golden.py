"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden.py            # needs /root/reference; deterministic, byte-stable output

Each fixture holds the flattened inputs (CSR corpus, label lists, initial z) and, per sweep, the
reference's own state (z, n_k_v, n_d_k, n_zk) after `training_iteration()` ran with only
`multinom_draw` replaced (oracle/patched_reference.py).  tests/test_oracle_vs_reference.py checks
oracle/gibbs_oracle.c against them on the CPU; tests/test_llda_gpu.py checks the device path.
"""
import io
import os
import sys
import zipfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import patched_reference as pr  # noqa: E402

GOLDEN = os.environ.get("LDA_GOLDEN_DIR") or os.path.join(os.path.dirname(_HERE), "tests", "golden")
SEED = 20261017
N_SWEEPS = 3


def save_npz(path, arrays):
    """np.savez_compressed with fixed zip timestamps so regenerated files are byte-identical."""
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as zf:
        for name in sorted(arrays):
            buf = io.BytesIO()
            np.lib.format.write_array(buf, np.asanyarray(arrays[name]), allow_pickle=False)
            info = zipfile.ZipInfo(name + ".npy", date_time=(1980, 1, 1, 0, 0, 0))
            info.compress_type = zipfile.ZIP_DEFLATED
            zf.writestr(info, buf.getvalue())


def load_slice(mod, n_docs, depth):
    """First n_docs rows of abstracts_data.csv through the reference's own load_corpus."""
    import csv
    import tempfile
    src = os.path.join(pr.REFERENCE_DIR, "abstracts_data.csv")
    with open(src, "r") as f:
        rows = [row for _, row in zip(range(n_docs), csv.reader(f))]
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False, newline="") as tmp:
        csv.writer(tmp).writerows(rows)
        name = tmp.name
    try:
        return mod.load_corpus(name, depth) if depth is not None else mod.load_corpus(name)
    finally:
        os.unlink(name)


def golden_llda(n_docs=300):
    mod = pr.import_reference("LabeledLDA")
    docs, labs, labelset = load_slice(mod, n_docs, 3)
    dicti = mod.prune_dict(docs, lower=0.01, upper=0.9)
    keep = [i for i, d in enumerate(docs) if len(dicti.doc2bow(d)) > 0]
    docs, labs = [docs[i] for i in keep], [labs[i] for i in keep]
    np.random.seed(0)
    model = mod.LabeledLDA(docs, labs, labelset, dicti, 0.1, 0.01)      # README.md:50 alpha, beta
    out = pr.flatten_llda(model)
    out["z_init"] = out.pop("z")
    out.update(K=np.int64(model.K), V=np.int64(model.V), alpha=np.float64(model.alpha), beta=np.float64(model.beta),
               seed=np.uint64(SEED))
    draw = pr.patch(mod, SEED)
    states = pr.run_llda_sweeps(mod, model, draw, N_SWEEPS)
    for s, st in enumerate(states):
        for k, v in st.items():
            out["s%d_%s" % (s, k)] = v
    phi, theta = model.get_phi(), model.get_theta()                     # LabeledLDA.py:231-239 after the last sweep
    out["phi_cols"] = np.arange(0, model.V, 7, dtype=np.int64)
    out["phi_sub"] = phi[:, ::7]
    out["phi_rowsum"] = phi.sum(axis=1)
    out["theta"] = theta
    save_npz(os.path.join(GOLDEN, "llda_abstracts%d.npz" % n_docs), out)
    print("llda: D=%d K=%d V=%d pairs=%d" % (model.D, model.K, model.V, len(out["word"])))


def golden_sublda(n_docs=300):
    """Root-scope SubLDA of CascadeLDA.go_down_tree (CascadeLDA.py:137-141) on the same slice."""
    mod = pr.import_reference("CascadeLDA")
    docs, labs, labelset = load_slice(mod, n_docs, 3)
    dicti = mod.prune_dict(docs, lower=0.01, upper=0.9)
    keep = [i for i, d in enumerate(docs) if len(dicti.doc2bow(d)) > 0]
    docs, labs = [docs[i] for i in keep], [labs[i] for i in keep]
    cas = mod.CascadeLDA(docs, labs, labelset, dicti, alpha=0.1, beta=0.01)
    np.random.seed(1)
    sub = mod.SubLDA(cas.doc_tups, cas.l1, list(cas.lablist_l1), dicti, alpha=0.1, beta=0.01)
    out = pr.flatten_llda(sub)
    out["z_init"] = out.pop("z")
    # SubLDA.__init__ iterates `zip(doc, zets, freqs)` over the (id, freq) TUPLES (CascadeLDA.py:382-385), so
    # `n_k_v[z, (id, freq)] += f` also credits column `freq`: the reference's initial table is the histogram
    # plus those spurious counts.  Parity means reproducing it, so the fixture carries the reference's table.
    out["n_k_v_init"] = np.asarray(sub.n_k_v, dtype=np.int32).copy()
    out.update(K=np.int64(sub.K), V=np.int64(sub.V), alpha=np.float64(sub.alpha), beta=np.float64(sub.beta),
               seed=np.uint64(SEED + 1))
    draw = pr.patch(mod, SEED + 1)
    states = pr.run_llda_sweeps(mod, sub, draw, N_SWEEPS)
    for s, st in enumerate(states):
        for k, v in st.items():
            out["s%d_%s" % (s, k)] = v
    out["ph"] = sub.get_ph()                                            # CascadeLDA.py:394-395 (unsmoothed)
    save_npz(os.path.join(GOLDEN, "sublda_abstracts%d.npz" % n_docs), out)
    print("sublda: D=%d K=%d V=%d pairs=%d" % (sub.D, sub.K, sub.V, len(out["word"])))


if __name__ == "__main__":
    if os.environ.get("PYTHONHASHSEED") != "0":       # CascadeLDA.load_corpus orders labels through set()
        os.environ["PYTHONHASHSEED"] = "0"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    os.makedirs(GOLDEN, exist_ok=True)
    golden_llda()
    golden_sublda()
