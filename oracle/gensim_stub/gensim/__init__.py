"""Minimal stand-in for the five gensim names the reference touches (TEST INFRASTRUCTURE ONLY).

gensim 2.3.0 (requirements.txt:5 of the reference) is not installed in this image and cannot be
fetched.  oracle/patched_reference.py puts this directory on sys.path so the UNMODIFIED reference
modules import and run.  Tokenisation differs from real gensim (no Porter stemming, short stop
list); that is irrelevant for parity because the reference and this repository's sampler consume
the same tokenised corpus and the same Dictionary object.
"""
