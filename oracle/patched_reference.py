"""Run the UNMODIFIED reference samplers with a counter-addressed draw (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference exists (the build container).  It never travels to the GPU box:
what travels are the fixtures oracle/make_golden.py writes under tests/golden/.

What is patched, and why
    The reference draws with `multinom_draw(1, prob).argmax()` (LabeledLDA.py:119, CascadeLDA.py:415,
    HSLDA.py:261) from NumPy's legacy global MT19937.  `np.random.multinomial` consumes a data-dependent
    number of uniforms per draw, so its stream cannot be addressed by (sweep, draw index) and no
    counter-based GPU generator can follow it.  The modules look the name `multinom_draw` up in their
    globals at call time, so replacing that ONE module attribute swaps the generator and nothing else:
    every other statement of training_iteration / sample_z runs as written, on the reference's own
    NumPy arrays.  The replacement is an inverse-CDF draw

        k = first index with cumsum(prob)[k] > u * cumsum(prob)[-1]

    with u = (word + 0.5) * 2^-32 and word = Philox4x32-10 addressed as in oracle/philox.py
    (stream 0, the sweep number, the document's index and the draw's position inside the document).
"""
import importlib
import os
import sys
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import philox  # noqa: E402

REFERENCE_DIR = os.environ.get("LDA_REFERENCE_DIR", "/root/reference")
STUB_DIR = os.path.join(_HERE, "gensim_stub")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "LabeledLDA.py"))


def import_reference(name):
    """Import /root/reference/<name>.py unmodified (gensim resolved to oracle/gensim_stub)."""
    if not reference_available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_DIR)
    for p in (REFERENCE_DIR, STUB_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")           # '\d' SyntaxWarning at LabeledLDA.py:24 etc.
        mod = importlib.import_module(name)
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_DIR)):
        raise RuntimeError("%s resolved to %s, not the reference" % (name, mod.__file__))
    return mod


class PhiloxDraw(object):
    """Drop-in for `multinom_draw`: returns a one-hot vector so `.argmax()` yields the drawn index."""

    def __init__(self, seed, stream=philox.STREAM_SWEEP):
        self.seed = int(seed)
        self.stream = stream
        self.sweep = 0
        self.t = 0
        self._words = None
        self.calls = 0

    def begin_sweep(self, sweep, doc_lens, doc_base=0):
        """doc_lens[d] = draws of document d; the draw at position p of document d uses word (doc_base + d, p)."""
        self.sweep = int(sweep)
        self.t = 0
        doc_lens = np.asarray(doc_lens, dtype=np.int64)
        doc = np.repeat(np.arange(doc_lens.shape[0], dtype=np.int64), doc_lens)
        starts = np.cumsum(doc_lens) - doc_lens
        pos = np.arange(doc.shape[0], dtype=np.int64) - starts[doc]
        self._words = philox.draw_words(self.seed, self.stream, self.sweep, doc + doc_base, pos)

    def __call__(self, n, prob):
        assert n == 1
        u = (float(self._words[self.t]) + 0.5) * (1.0 / 4294967296.0)
        self.t += 1
        self.calls += 1
        cs = np.cumsum(np.asarray(prob, dtype=np.float64))
        k = int(np.argmax(cs > u * cs[-1]))
        out = np.zeros(len(cs), dtype=np.int64)
        out[k] = 1
        return out


def patch(module, seed):
    draw = PhiloxDraw(seed)
    module.multinom_draw = draw
    return draw


# ---------------------------------------------------------------------------------- state flattening
def flatten_llda(model):
    """LabeledLDA / SubLDA instance -> CSR arrays in the layout of include/gibbs_b200.h."""
    lens = np.array([len(d) for d in model.docs], dtype=np.int64)
    doc_ptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    word = np.array([v for d in model.docs for v in d], dtype=np.int32)
    freq = np.array([f for d in model.freqs for f in d], dtype=np.int32)
    z = np.array([k for d in model.z_dn for k in d], dtype=np.int32)
    rows, cols = np.nonzero(np.asarray(model.labs))
    lab_ptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=len(lens)), out=lab_ptr[1:])
    return dict(doc_ptr=doc_ptr, word=word, freq=freq, z=z, lab_ptr=lab_ptr, lab_idx=cols.astype(np.int32))


def llda_state(model):
    return dict(z=np.array([k for d in model.z_dn for k in d], dtype=np.int32),
                n_k_v=np.asarray(model.n_k_v, dtype=np.int32).copy(),
                n_d_k=np.asarray(model.n_d_k, dtype=np.int32).copy(),
                n_zk=np.asarray(model.n_zk, dtype=np.int32).copy())


def run_llda_sweeps(module, model, draw, n_sweeps, first_sweep=0):
    """training_iteration() x n_sweeps under the patched draw; returns the state after each sweep."""
    n_draws = sum(len(d) for d in model.docs)
    out = []
    for s in range(first_sweep, first_sweep + n_sweeps):
        draw.begin_sweep(s, [len(d) for d in model.docs])
        model.training_iteration()
        assert draw.t == n_draws, "exactly one draw per pair per sweep"
        out.append(llda_state(model))
    return out
