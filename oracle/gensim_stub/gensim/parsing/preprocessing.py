"""preprocess_documents stand-in (LabeledLDA.py:45, CascadeLDA.py:47, HSLDA.py:78)."""
import re

_STOP = frozenset("""a about above after again against all also am an and any are as at be because been before being
below between both but by can could did do does doing down during each few for from further had has have having he her
here hers him his how however i if in into is it its itself just may me more most my no nor not of off on once only or
other our out over own same she should so some such than that the their them then there these they this those through to
too under until up us very was we were what when where which while who whom why will with would you your paper find
using use used show results study""".split())
_TOKEN = re.compile(r"[a-z]+")


def preprocess_string(s):
    return [t for t in _TOKEN.findall(s.lower()) if len(t) >= 3 and t not in _STOP]


def preprocess_documents(docs):
    return [preprocess_string(d) for d in docs]
