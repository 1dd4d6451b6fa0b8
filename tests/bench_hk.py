"""Stand-alone timing of HK_from_spectrum_cpp's GEMM (gpurun helper, not a pytest file):
python tests/bench_hk.py [n0] [n1] [K]   -- default: the 5000 x 5000 training block of BASELINE config 4, K = 200"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402
from flgp_b200.datasets import make  # noqa: E402

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
n = 200_000
X, Y, cfg = make("C4", 3, n=n)
ctx = F.default_ctx()
ep = F.heat_kernel_spectrum_cpp(X[:5000], X[5000:], 2000, 3, K, seed=1, iter_max=10)
idx0 = np.arange(n0, dtype=np.int32)
idx1 = np.arange(n1, dtype=np.int32)
peak = ctx.dfma_peak_tflops()
ctx.set_timing(True)
for it in range(3):
    ctx.stage_reset()
    H = F.HK_from_spectrum_cpp(ep, K, 10.0, idx0, idx1)
    st = {d["name"]: d for d in ctx.stages()}
    g = st["hk_gemm"]
    print("run %d: hk_gemm %.3f ms = %.2f TFLOP/s (%.0f%% of the measured FP64 FMA peak %.1f), whole call %.1f ms" %
          (it, g["ms"], g["flops"] / g["ms"] / 1e9, 100 * g["flops"] / g["ms"] / 1e9 / peak, peak,
           st["hk_from_spectrum"]["ms"]))
V = ep.rows(np.arange(max(n0, n1), dtype=np.int32))
lam = np.exp(-10.0 * (1.0 - ep.values[:K]))
Ho = (V[:n0, :K] * lam) @ V[:n1, :K].T
print("max |H - Ho| = %.3e (scale %.3e)" % (np.abs(H - Ho).max(), np.abs(Ho).max()))
