"""Porter (1980) suffix-stripping stemmer, written from the published algorithm description.

Stands in for gensim's ``stem_text`` (used by ``preprocess_documents`` at LabeledLDA.py:45,
CascadeLDA.py:48, HSLDA.py:78 of the reference): gensim is not installable here,
and the drop-in modules must tokenise on their own.  Token ids need not match a gensim run --
the parity oracle and the GPU path always consume the same tokenised corpus.
"""

_VOWELS = frozenset("aeiou")


def _is_cons(w, i):
    c = w[i]
    if c in _VOWELS:
        return False
    if c == "y":
        return i == 0 or not _is_cons(w, i - 1)
    return True


def _measure(w):
    """Number of VC sequences in w ([C](VC)^m[V])."""
    m, i, n = 0, 0, len(w)
    while i < n and _is_cons(w, i):
        i += 1
    while i < n:
        while i < n and not _is_cons(w, i):
            i += 1
        if i >= n:
            break
        m += 1
        while i < n and _is_cons(w, i):
            i += 1
    return m


def _has_vowel(w):
    return any(not _is_cons(w, i) for i in range(len(w)))


def _double_cons(w):
    return len(w) >= 2 and w[-1] == w[-2] and _is_cons(w, len(w) - 1)


def _cvc(w):
    n = len(w)
    if n < 3:
        return False
    if not (_is_cons(w, n - 3) and not _is_cons(w, n - 2) and _is_cons(w, n - 1)):
        return False
    return w[-1] not in "wxy"


_STEP2 = (
    ("ational", "ate"), ("tional", "tion"), ("enci", "ence"), ("anci", "ance"), ("izer", "ize"),
    ("bli", "ble"), ("alli", "al"), ("entli", "ent"), ("eli", "e"), ("ousli", "ous"),
    ("ization", "ize"), ("ation", "ate"), ("ator", "ate"), ("alism", "al"), ("iveness", "ive"),
    ("fulness", "ful"), ("ousness", "ous"), ("aliti", "al"), ("iviti", "ive"), ("biliti", "ble"),
    ("logi", "log"),
)
_STEP3 = (
    ("icate", "ic"), ("ative", ""), ("alize", "al"), ("iciti", "ic"), ("ical", "ic"),
    ("ful", ""), ("ness", ""),
)
_STEP4 = (
    "al", "ance", "ence", "er", "ic", "able", "ible", "ant", "ement", "ment", "ent", "ion", "ou",
    "ism", "ate", "iti", "ous", "ive", "ize",
)


def _replace_longest(w, table, min_m):
    best = None
    for suf, rep in table:
        if w.endswith(suf) and (best is None or len(suf) > len(best[0])):
            best = (suf, rep)
    if best is None:
        return w
    stem = w[: len(w) - len(best[0])]
    if _measure(stem) > min_m:
        return stem + best[1]
    return w


def stem(word):
    w = word
    if len(w) <= 2:
        return w
    # step 1a
    if w.endswith("sses"):
        w = w[:-2]
    elif w.endswith("ies"):
        w = w[:-2]
    elif w.endswith("ss"):
        pass
    elif w.endswith("s"):
        w = w[:-1]
    # step 1b
    fired = False
    if w.endswith("eed"):
        if _measure(w[:-3]) > 0:
            w = w[:-1]
    elif w.endswith("ed") and _has_vowel(w[:-2]):
        w, fired = w[:-2], True
    elif w.endswith("ing") and _has_vowel(w[:-3]):
        w, fired = w[:-3], True
    if fired:
        if w.endswith(("at", "bl", "iz")):
            w += "e"
        elif _double_cons(w) and w[-1] not in "lsz":
            w = w[:-1]
        elif _measure(w) == 1 and _cvc(w):
            w += "e"
    # step 1c
    if w.endswith("y") and _has_vowel(w[:-1]):
        w = w[:-1] + "i"
    # steps 2, 3
    w = _replace_longest(w, _STEP2, 0)
    w = _replace_longest(w, _STEP3, 0)
    # step 4
    best = None
    for suf in _STEP4:
        if w.endswith(suf) and (best is None or len(suf) > len(best)):
            best = suf
    if best is not None:
        stem_ = w[: len(w) - len(best)]
        if _measure(stem_) > 1 and (best != "ion" or (stem_ and stem_[-1] in "st")):
            w = stem_
    # step 5
    if w.endswith("e"):
        m = _measure(w[:-1])
        if m > 1 or (m == 1 and not _cvc(w[:-1])):
            w = w[:-1]
    if _measure(w) > 1 and _double_cons(w) and w.endswith("l"):
        w = w[:-1]
    return w
