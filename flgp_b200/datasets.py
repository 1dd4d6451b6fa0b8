"""Synthetic inputs of the five BASELINE.json configs, fixed seeds, fp64.

Rows are generated in blocks of 2^16 from a (seed, block) keyed generator so that any shard of any
config can be produced without generating the rest, and the union is independent of the number of
ranks.  There is no network in the build environment: all data is synthetic by construction.
"""
from __future__ import annotations

import numpy as np

BLOCK = 1 << 16

CONFIGS = {
    "C1": dict(name="rings", n=4800, d=2, m=100, s=600, r=3, K=100),
    "C2": dict(name="spiral", n=4000, d=2, m=200, s=500, r=3, K=100),
    "C3": dict(name="clusters784", n=70000, d=784, m=1000, s=1000, r=5, K=200),
    "C4": dict(name="swissroll", n=10_000_000, d=3, m=5000, s=2000, r=3, K=200),
    "C5": dict(name="torus16", n=100_000_000, d=16, m=5000, s=4000, r=5, K=300),
}


def _blocks(lo, hi):
    b = lo // BLOCK
    while b * BLOCK < hi:
        yield b, max(lo, b * BLOCK) - b * BLOCK, min(hi, (b + 1) * BLOCK) - b * BLOCK
        b += 1


def _gen(kind, seed, lo, hi, d):
    X = np.empty((hi - lo, d), order="F")
    Y = np.empty(hi - lo)
    at = 0
    aux = None
    if kind == "clusters784":
        aux = np.random.default_rng([seed, 1 << 30]).standard_normal((10, d)) * 3.0
    if kind == "torus16":
        q, _ = np.linalg.qr(np.random.default_rng([seed, 1 << 30]).standard_normal((16, 4)))
        aux = q.T  # 4 x 16, orthonormal rows
    for b, a, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        if kind == "swissroll":  # C4
            t = rng.uniform(1.5 * np.pi, 4.5 * np.pi, BLOCK)
            h = rng.uniform(0.0, 21.0, BLOCK)
            nz = rng.standard_normal(BLOCK) * 0.1
            xb = np.stack([t * np.cos(t), h, t * np.sin(t)], axis=1)
            yb = np.sin(t) + h / 21.0 + nz
        elif kind == "spiral":  # C2, README.md:115-133
            th = rng.uniform(0.0, 8.0 * np.pi, BLOCK)
            nz = rng.standard_normal(BLOCK)
            rad = (th + 4.0) ** 0.7
            xb = np.stack([rad * np.cos(th), rad * np.sin(th)], axis=1)
            yb = 3 * np.sin(th / 10) + 3 * np.cos(th / 2) + 4 * np.sin(4 * th / 5) + nz
        elif kind == "clusters784":  # C3
            lab = rng.integers(0, 10, BLOCK)
            xb = aux[lab] + rng.standard_normal((BLOCK, d))
            xb = np.clip(xb / 8.0 + 0.5, 0.0, 1.0)
            yb = lab.astype(np.float64)
        elif kind == "torus16":  # C5
            a_ = rng.uniform(0, 2 * np.pi, BLOCK)
            b_ = rng.uniform(0, 2 * np.pi, BLOCK)
            core = np.stack([np.cos(a_), np.sin(a_), np.cos(b_), np.sin(b_)], axis=1)
            xb = core @ aux + rng.standard_normal((BLOCK, 16)) * 0.01
            yb = np.sin(a_) * np.cos(b_)
        else:
            raise ValueError(kind)
        X[at:at + e - a] = xb[a:e]
        Y[at:at + e - a] = yb[a:e]
        at += e - a
    return X, Y


def rings(seed=1234):
    """C1: six concentric rings, 800 points each, standardised columns / sqrt(d) (README.md:44-55)."""
    rng = np.random.default_rng(seed)
    xs, ys = [], []
    for i in range(6):
        th = rng.uniform(0, 2 * np.pi, 800)
        rad = 0.5 + 0.1 * i
        xs.append(np.stack([rad * np.cos(th), rad * np.sin(th)], axis=1))
        ys.append(np.full(800, 1.0 if (-1) ** i > 0 else 0.0))
    X = np.vstack(xs)
    Y = np.concatenate(ys)
    X = (X - X.mean(0)) / X.std(0, ddof=1) / np.sqrt(2.0)
    perm = rng.permutation(len(X))  # training rows first: m sampled without replacement
    return np.asfortranarray(X[perm]), Y[perm]


def make(config: str, seed: int = 1234, lo: int = 0, hi: int | None = None, n: int | None = None):
    """Rows [lo, hi) of a config's X_all (training rows are rows 0..m-1) and the labels of those rows."""
    cfg = dict(CONFIGS[config])
    if n is not None:
        cfg["n"] = n
    hi = cfg["n"] if hi is None else hi
    if cfg["name"] == "rings":
        X, Y = rings(seed)
        return np.asfortranarray(X[lo:hi]), Y[lo:hi], cfg
    X, Y = _gen(cfg["name"], seed, lo, hi, cfg["d"])
    return X, Y, cfg


def shard_bounds(n_total: int, nranks: int, rank: int):
    """Contiguous row blocks of ceil(n / nranks) rows."""
    per = -(-n_total // nranks)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)
