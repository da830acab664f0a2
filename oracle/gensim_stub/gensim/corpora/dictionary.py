"""Dictionary stand-in: the subset of gensim.corpora.dictionary.Dictionary the reference uses
(LabeledLDA.py:59-64, 156, 282-284; CascadeLDA.py:67-72, 187, 361, 451-453)."""
from collections import Counter


class Dictionary(object):
    def __init__(self, documents=None):
        self.token2id = {}
        self.id2token = {}
        self.dfs = {}
        self.num_docs = 0
        if documents is not None:
            for doc in documents:
                self.doc2bow(doc, allow_update=True)

    def __len__(self):
        return len(self.token2id)

    def values(self):
        return [self.id2token[i] for i in sorted(self.id2token)]

    def doc2bow(self, document, allow_update=False):
        counter = Counter(document)
        if allow_update:
            for w in sorted(w for w in counter if w not in self.token2id):
                i = len(self.token2id)
                self.token2id[w] = i
                self.id2token[i] = w
            self.num_docs += 1
            for w in counter:
                i = self.token2id[w]
                self.dfs[i] = self.dfs.get(i, 0) + 1
        t2i = self.token2id
        return sorted((t2i[w], c) for w, c in counter.items() if w in t2i)

    def filter_extremes(self, no_below=5, no_above=0.5, keep_n=100000):
        hi = no_above * self.num_docs
        good = [i for i in self.id2token if no_below <= self.dfs.get(i, 0) <= hi]
        good = sorted(good, key=lambda i: -self.dfs.get(i, 0))[:keep_n]
        old = sorted(good)
        remap = {o: n for n, o in enumerate(old)}
        self.token2id = {self.id2token[o]: remap[o] for o in old}
        self.dfs = {remap[o]: self.dfs[o] for o in old}
        self.id2token = {n: w for w, n in self.token2id.items()}
