"""Multi-GPU invariance check, launched by torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank runs its shard through the sharded C ABI; rank 0 also runs the whole matrix on one GPU (no
communicator) and checks: k-means centres/sizes, Z pattern and values BIT-EXACT for any rank count (integer
limb all-reduces); eigenvalues and predictions to 1e-10 (fp64 K x K all-reduce order differs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make, shard_bounds  # noqa: E402
from flgp_b200.sharding import init_context_comm  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, s, r, K = 200_003, 700, 300, 3, 40
    X, Y, _ = make("C4", 77, n=n)
    lo, hi = shard_bounds(n, world, rank)
    init = F.default_init(n, s, 5)
    ctx = F.Context(local)
    init_context_comm(ctx, rank, world)
    ep = F.heat_kernel_spectrum_sharded(np.asfortranarray(X[lo:hi]), n, lo, s, r, K, init_idx=init, iter_max=30, ctx=ctx)
    m_local = max(0, min(hi - lo, m - lo))
    y, cov = F.regression_fixed(ep, Y[lo:lo + m_local], m, K, (10.0, 0.01), 1e-5)
    U = ep.anchors()
    Z = ep.Z()
    vals = ep.values
    iters = ep.kmeans_iters
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(lo=lo, hi=hi, U=U, Zj=Z.indices, Zx=Z.data, vals=vals, y=y, cov=cov,
                                          iters=iters))
    ok = True
    if rank == 0:
        ctx1 = F.Context(local)  # a second context in this process: no communicator, whole matrix
        ep1 = F.heat_kernel_spectrum_sharded(X, n, 0, s, r, K, init_idx=init, iter_max=30, ctx=ctx1)
        y1, cov1 = F.regression_fixed(ep1, Y[:m], m, K, (10.0, 0.01), 1e-5)
        U1, Z1, v1 = ep1.anchors(), ep1.Z(), ep1.values
        for g in gathered:
            ok &= g["iters"] == ep1.kmeans_iters
            ok &= np.array_equal(g["U"], U1)
            ok &= np.allclose(g["vals"], v1, rtol=1e-10, atol=0)
        Zj = np.concatenate([g["Zj"] for g in gathered])
        Zx = np.concatenate([g["Zx"] for g in gathered])
        ok &= np.array_equal(Zj, Z1.indices) and np.array_equal(Zx, Z1.data)
        yy = np.concatenate([g["y"] for g in gathered])
        cc = np.concatenate([g["cov"] for g in gathered])
        e_y = np.abs(yy - y1).max() / np.abs(y1).max()
        e_c = np.abs(cc - cov1).max() / max(1.0, np.abs(cov1).max())
        ok &= e_y < 1e-9 and e_c < 1e-9
        print("multi_gpu_check world=%d: iters=%d U/Z bit-exact=%s, dy=%.2e dcov=%.2e -> %s" %
              (world, ep1.kmeans_iters, np.array_equal(Zx, Z1.data), e_y, e_c, "OK" if ok else "FAIL"))
    ok &= large_d_case(rank, local, world, ctx)
    ok &= large_s_case(rank, local, world, ctx)
    ok &= nystrom_case(rank, local, world, ctx)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def large_d_case(rank, local, world, ctx):
    """d = 16 (config 5's dimension): the tensor-core distance path with persistent integer accumulators, sharded."""
    n, m, s, r, K, d = 60_001, 300, 128, 5, 30, 16
    rng = np.random.default_rng(123)
    a, b = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    T4 = np.c_[np.cos(a), np.sin(a), np.cos(b), np.sin(b)]
    Q, _ = np.linalg.qr(rng.standard_normal((d, 4)))
    X = np.asfortranarray(T4 @ Q.T + 0.01 * rng.standard_normal((n, d)))
    lo, hi = shard_bounds(n, world, rank)
    init = F.default_init(n, s, 6)
    ep = F.heat_kernel_spectrum_sharded(np.asfortranarray(X[lo:hi]), n, lo, s, r, K, init_idx=init, iter_max=12, ctx=ctx)
    Z = ep.Z()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(U=ep.anchors(), Zj=Z.indices, Zx=Z.data, vals=ep.values, iters=ep.kmeans_iters))
    ok = True
    if rank == 0:
        ctx1 = F.Context(local)
        ep1 = F.heat_kernel_spectrum_sharded(X, n, 0, s, r, K, init_idx=init, iter_max=12, ctx=ctx1)
        Z1 = ep1.Z()
        for g in gathered:
            ok &= g["iters"] == ep1.kmeans_iters and np.array_equal(g["U"], ep1.anchors())
            ok &= np.allclose(g["vals"], ep1.values, rtol=1e-10, atol=0)
        ok &= np.array_equal(np.concatenate([g["Zj"] for g in gathered]), Z1.indices)
        ok &= np.array_equal(np.concatenate([g["Zx"] for g in gathered]), Z1.data)
        print("multi_gpu_check world=%d large-d (d=16): U/Z bit-exact -> %s" % (world, "OK" if ok else "FAIL"))
    return ok


def large_s_case(rank, local, world, ctx):
    """s = 1100, K = 100: the iterative eigensolver (chfsi.cu) with its filter columns sharded over the ranks and
    all-gathered; every column is computed by one rank with the single-GPU instruction sequence, so eigenvalues and
    predictions must be BIT-identical to the one-GPU run (only the K x K tail all-reduce can differ in the last bits)."""
    n, m, s, r, K = 150_001, 900, 1100, 3, 100
    X, Y, _ = make("C4", 99, n=n)
    lo, hi = shard_bounds(n, world, rank)
    init = F.default_init(n, s, 9)
    ep = F.heat_kernel_spectrum_sharded(np.asfortranarray(X[lo:hi]), n, lo, s, r, K, init_idx=init, iter_max=15, ctx=ctx)
    m_local = max(0, min(hi - lo, m - lo))
    y, cov = F.regression_fixed(ep, Y[lo:lo + m_local], m, K, (10.0, 0.01), 1e-5)
    rows = ep.rows(np.arange(0, hi - lo, 997, dtype=np.int32))
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(vals=ep.values, y=y, rows=rows))
    ok = True
    if rank == 0:
        ctx1 = F.Context(local)
        ep1 = F.heat_kernel_spectrum_sharded(X, n, 0, s, r, K, init_idx=init, iter_max=15, ctx=ctx1)
        y1, _ = F.regression_fixed(ep1, Y[:m], m, K, (10.0, 0.01), 1e-5)
        idx1 = np.concatenate([np.arange(lo_, hi_, 997) for lo_, hi_ in (shard_bounds(n, world, q) for q in range(world))])
        rows1 = ep1.rows(idx1.astype(np.int32))
        bit_vals = all(np.array_equal(g["vals"], ep1.values) for g in gathered)
        bit_rows = np.array_equal(np.concatenate([g["rows"] for g in gathered]), rows1)
        yy = np.concatenate([g["y"] for g in gathered])
        e_y = np.abs(yy - y1).max() / np.abs(y1).max()
        ok = bit_vals and bit_rows and e_y < 1e-9
        print("multi_gpu_check world=%d large-s (s=%d, iterative eigensolver, sharded filter): eigenvalues bit-exact=%s, "
              "eigenvector rows bit-exact=%s, dy=%.2e -> %s" % (world, s, bit_vals, bit_rows, e_y, "OK" if ok else "FAIL"))
    return ok


def nystrom_case(rank, local, world, ctx):
    """fit_nystrom_regression sharded over the ranks against the single-process entry on the whole matrix (fixed pars:
    the comparison is of the extension and the tail, not of the optimiser path)."""
    n, m, s, K = 40_001, 400, 200, 30
    X, Y, _ = make("C4", 55, n=n)
    lo, hi = shard_bounds(n, world, rank)
    init = F.default_init(n, s, 2)
    a2s = np.array([0.5, 2.0])
    m_local = max(0, min(hi - lo, m - lo))
    res = F.fit_nystrom_regression_sharded(np.asfortranarray(X[lo:hi]), n, lo, Y[lo:lo + m_local], m, s, K, a2s=a2s,
                                           pars=(10.0, 0.05), init_idx=init, iter_max=20, ctx=ctx)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(mean=res["mean"], cov=res["cov"], a2=res["a2"], obj=res["obj"]))
    ok = True
    if rank == 0:
        ctx1 = F.Context(local)
        one = F.fit_nystrom_regression_gp_rcpp(X[:m], Y[:m], X[m:], s, K, a2s=a2s, pars=(10.0, 0.05), init_idx=init,
                                               iter_max=20, ctx=ctx1)
        mean = np.concatenate([g["mean"] for g in gathered])
        cov = np.concatenate([g["cov"] for g in gathered])
        ref_mean = np.r_[one["Y_pred"]["train"], one["Y_pred"]["test"]]
        e_m = np.abs(mean - ref_mean).max() / np.abs(ref_mean).max()
        e_c = np.abs(cov[m:] - one["posterior"]["cov"]).max() / max(1.0, np.abs(one["posterior"]["cov"]).max())
        ok = all(g["a2"] == one["a2"] for g in gathered) and e_m < 1e-9 and e_c < 1e-9
        print("multi_gpu_check world=%d Nystrom sharded: same a2=%s, dmean=%.2e dcov=%.2e -> %s" %
              (world, all(g["a2"] == one["a2"] for g in gathered), e_m, e_c, "OK" if ok else "FAIL"))
    return ok


if __name__ == "__main__":
    main()
