"""One rank of tests/test_multi_shard.py.  argv: mode (gloo|nccl), scratch dir[, n_refresh]."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from lda_thesis_b200 import synth  # noqa: E402


def main():
    mode, tmp = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    with np.load(os.path.join(tmp, "corpus.npz")) as f:
        c = {k: f[k] for k in f.files}
    c["D"], c["K"], c["V"] = int(c["D"]), int(c["K"]), int(c["V"])
    c.setdefault("freq", None)
    cuts = synth.shard_bounds(c["doc_ptr"], world)
    d0, d1 = cuts[rank], cuts[rank + 1]
    s = synth.take_docs(c, d0, d1)
    if mode == "gloo":
        import torch
        import torch.distributed as dist
        import oracle_lib
        dist.init_process_group("gloo", rank=rank, world_size=world)
        o = oracle_lib.LldaOracle(s["doc_ptr"], s["word"], s["freq"], s["lab_ptr"], s["lab_idx"], s["K"], s["V"],
                                  0.1, 0.01, seed=17, doc_base=d0)
        # initial histograms: sum over shards (gibbs_load does this with ncclAllReduce)
        t = torch.from_numpy(o.n_wk_pad)
        dist.all_reduce(t)
        tk = torch.from_numpy(o.n_k)
        dist.all_reduce(tk)
        for _ in range(3):
            before_wk, before_k = o.n_wk_pad.copy(), o.n_k.copy()
            o.snapshot_sweep(1)                       # samples this shard against the replicated table
            d_wk = torch.from_numpy(o.n_wk_pad - before_wk)
            d_k = torch.from_numpy(o.n_k - before_k)
            dist.all_reduce(d_wk)                     # the one exchange step of the sweep
            dist.all_reduce(d_k)
            o.n_wk_pad[:] = before_wk + d_wk.numpy()
            o.n_k[:] = before_k + d_k.numpy()
        np.save(os.path.join(tmp, "z_%d.npy" % rank), o.z)
        np.save(os.path.join(tmp, "n_wk_%d.npy" % rank), o.n_wk)
        np.save(os.path.join(tmp, "n_k_%d.npy" % rank), o.n_k)
        dist.destroy_process_group()
        return 0
    # ---- nccl: the product path, communicator id passed through a file
    from lda_thesis_b200 import _lib
    n_refresh = int(sys.argv[3])
    idf = os.path.join(tmp, "nccl_id_%d.bin" % n_refresh)
    if rank == 0:
        uid = _lib.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        for _ in range(600):
            if os.path.exists(idf):
                break
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    g = _lib.GibbsSampler(s["D"], s["V"], s["K"], 0.1, 0.01, seed=17, mode="snapshot", device=rank,
                          n_refresh=n_refresh, tile_docs=32, doc_base=d0)
    g.comm_init(world, rank, uid)
    g.load(s["doc_ptr"], s["word"], s["freq"], None, s["lab_ptr"], s["lab_idx"])
    g.sweep(3)
    st = g.get_state()
    np.save(os.path.join(tmp, "z_%d.npy" % rank), st["z"])
    np.save(os.path.join(tmp, "n_wk_%d.npy" % rank), st["n_wk"])
    np.save(os.path.join(tmp, "n_k_%d.npy" % rank), st["n_k"])
    g.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
