"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous row blocks, int64 limb all-reduces and the
claim the CUDA path relies on — every reduced quantity is bit-identical for any number of ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    from conftest import swiss
    from flgp_b200.datasets import shard_bounds
    from flgp_b200.sharding import exchange_unique_id

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, s, r = 3001, 37, 3
    X, _ = swiss(n, 21)
    lo, hi = shard_bounds(n, world, rank)
    Xl = np.asfortranarray(X[lo:hi])
    init = np.sort(np.random.default_rng(3).choice(n, s, replace=False)).astype(np.int32)
    # the unique-id exchange used for the NCCL communicator (here with a fake 128-byte id)
    uid = exchange_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    # global max |x| (all-reduce max), then Lloyd with all-reduced int64 accumulators
    m = torch.tensor([O.lib().orc_maxabs(Xl.ctypes.data_as(O._p(Xl).__class__), O.I64(Xl.size))])
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    maxabs = float(m[0])
    # initial centres: owner contributes the bit pattern, others zero (int64 sum is exact)
    Cb = np.zeros((s, 3), np.int64, order="F")
    for j, g in enumerate(init):
        if lo <= g < hi:
            Cb[j] = Xl[g - lo].view(np.int64)
    t = torch.from_numpy(np.ascontiguousarray(Cb))
    dist.all_reduce(t)
    Cc = np.asfortranarray(t.numpy().view(np.float64))
    assign = np.full(hi - lo, -1, np.int32)
    words = O.kmeans_acc_words(s, 3)
    iters = 0
    sizes = None
    for it in range(100):
        iters += 1
        acc = np.zeros(words, np.int64)
        O.kmeans_step(Xl, Cc, maxabs, n, assign, acc)
        ta = torch.from_numpy(acc)
        dist.all_reduce(ta)
        Cn, sizes = O.kmeans_update(acc, Cc, maxabs, n)
        if acc[-1] == 0:
            break
        Cc = Cn
    U = np.asfortranarray(np.c_[Cc, sizes])
    # KNN + LAE are row-local; column sums and Gram are reduced as limbs
    Zj, Zx, _ = O.lae(Xl, U[:, :3], r)
    hi1 = np.zeros(s, np.int64)
    lo1 = np.zeros(s, np.int64)
    O.colsum(Zj, Zx, s, 1, n, (hi1, lo1))
    for a in (hi1, lo1):
        dist.all_reduce(torch.from_numpy(a))
    c1 = O.fx_decode(hi1, lo1, 1.0, n)
    Zg = O.graph_laplacian_apply(Zj, Zx, s, "cluster-normalized", c1, U[:, 3])
    hi2 = np.zeros(s, np.int64)
    lo2 = np.zeros(s, np.int64)
    O.colsum(Zj, Zg, s, 1, n, (hi2, lo2))
    for a in (hi2, lo2):
        dist.all_reduce(torch.from_numpy(a))
    w = O.spectrum_scale(O.fx_decode(hi2, lo2, 1.0, n))
    gh = np.zeros(s * s, np.int64)
    gl = np.zeros(s * s, np.int64)
    O.gram(Zj, Zg, w, s, 1, n, (gh, gl))
    for a in (gh, gl):
        dist.all_reduce(torch.from_numpy(a))
    G = O.fx_decode(gh, gl, 1.0, n).reshape(s, s)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), U=U, assign=assign, Zj=Zj, Zg=Zg, G=G, iters=iters, lo=lo,
             hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_shard_invariance(tmp_path, oracle):
    from conftest import swiss

    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    n, s, r = 3001, 37, 3
    X, _ = swiss(n, 21)
    init = np.sort(np.random.default_rng(3).choice(n, s, replace=False)).astype(np.int32)
    U, assign, iters = oracle.kmeans_lloyd(X, s, init)
    Zj, Zx, _ = oracle.lae(X, U[:, :3], r)
    Zg = oracle.graph_laplacian(Zj, Zx, s, "cluster-normalized", U[:, 3], 1)
    c2 = oracle.colsum(Zj, Zg, s, 1)
    G = oracle.gram(Zj, Zg, oracle.spectrum_scale(c2), s, 1)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(2)]
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == n
    for p in parts:
        assert int(p["iters"]) == iters
        assert np.array_equal(p["U"], U)  # bit-exact centres and sizes on every rank
        assert np.array_equal(p["G"], np.ascontiguousarray(G).reshape(s, s))
    assert np.array_equal(np.concatenate([p["assign"] for p in parts]), assign)
    assert np.array_equal(np.vstack([p["Zj"] for p in parts]), Zj)
    assert np.array_equal(np.vstack([p["Zg"] for p in parts]), Zg)


def test_shard_bounds_cover_rows():
    from flgp_b200.datasets import shard_bounds

    for n in (1, 7, 100, 10_000_000):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, w, k) for k in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[k][1] == b[k + 1][0] for k in range(w - 1))
            assert max(h - l for l, h in b) == -(-n // w)
