"""Document sharding: a chain must not depend on how documents are split over ranks.

CPU (gloo, world_size 2): the host-side protocol -- contiguous shards balanced by draws, `doc_base` RNG addressing,
sum all-reduce of the word-topic deltas -- replayed with the oracle, must equal the unsharded oracle run.
GPU (needs 2 devices): the same through libgibbs_b200's in-library NCCL all-reduce, bit-identical to one GPU.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_corpus

WORKER = os.path.join(ROOT, "tests", "shard_worker.py")


def _run_workers(mode, world, tmp_path, extra=()):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29653",
                   LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, WORKER, mode, str(tmp_path)] + list(extra), env=env))
    rcs = [p.wait(timeout=600) for p in procs]
    assert rcs == [0] * world, "worker exit codes %s" % rcs


def test_two_rank_gloo_protocol_matches_unsharded(oracle, tmp_path):
    from lda_thesis_b200 import synth
    c = make_corpus(31, D=700, K=40, V=300, label_lens=[1, 2, 3, 5, 8, 9, 12], mean_pairs=15)
    np.savez(os.path.join(str(tmp_path), "corpus.npz"), **{k: v for k, v in c.items() if v is not None})
    _run_workers("gloo", 2, tmp_path)
    o = oracle.LldaOracle(c["doc_ptr"], c["word"], c["freq"], c["lab_ptr"], c["lab_idx"], c["K"], c["V"], 0.1, 0.01,
                          seed=17)
    o.snapshot_sweep(3)
    cuts = synth.shard_bounds(c["doc_ptr"], 2)
    z = np.concatenate([np.load(os.path.join(str(tmp_path), "z_%d.npy" % r)) for r in range(2)])
    assert np.array_equal(z, o.z)
    for r in range(2):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "n_wk_%d.npy" % r)), o.n_wk)
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "n_k_%d.npy" % r)), o.n_k)
    assert cuts[0] == 0 and cuts[-1] == c["D"] and 0 < cuts[1] < c["D"]


@pytest.mark.gpu
@pytest.mark.parametrize("n_refresh", [1, 3])
def test_two_gpus_bit_identical_to_one(gibbs, oracle, tmp_path, n_refresh):
    if gibbs.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    c = make_corpus(32, D=900, K=64, V=500, label_lens=[1, 2, 4, 7, 8, 9, 16, 20, 33, 40], mean_pairs=20)
    np.savez(os.path.join(str(tmp_path), "corpus.npz"), **{k: v for k, v in c.items() if v is not None})
    _run_workers("nccl", 2, tmp_path, extra=[str(n_refresh)])
    g = gibbs.GibbsSampler(c["D"], c["V"], c["K"], 0.1, 0.01, seed=17, mode="snapshot", n_refresh=n_refresh,
                           tile_docs=32)
    g.load(c["doc_ptr"], c["word"], c["freq"], None, c["lab_ptr"], c["lab_idx"])
    g.sweep(3)
    st = g.get_state()
    o = oracle.LldaOracle(c["doc_ptr"], c["word"], c["freq"], c["lab_ptr"], c["lab_idx"], c["K"], c["V"], 0.1, 0.01,
                          seed=17)
    o.snapshot_sweep(3, n_refresh=n_refresh, tile_docs=32)
    assert np.array_equal(st["z"], o.z) and np.array_equal(st["n_wk"], o.n_wk)
    z = np.concatenate([np.load(os.path.join(str(tmp_path), "z_%d.npy" % r)) for r in range(2)])
    assert np.array_equal(z, st["z"])
    for r in range(2):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "n_wk_%d.npy" % r)), st["n_wk"])
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "n_k_%d.npy" % r)), st["n_k"])
    g.close()
