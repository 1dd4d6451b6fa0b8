"""Stand-alone timing of the dense top-K eigensolver (gpurun helper, not a pytest file):
python tests/bench_eigh.py [s] [K]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.standard_normal((s, s)))
lam = 1.0 / (1.0 + 0.02 * np.arange(s)) ** 2
A = np.asfortranarray((Q * lam) @ Q.T)
A = (A + A.T) / 2
ctx = F.default_ctx()
ctx.set_timing(True)
for it in range(3):
    ctx.stage_reset()
    t0 = time.perf_counter()
    res = F.eigs_sym(A, K)
    t1 = time.perf_counter()
    st = {d["name"]: round(d["ms"], 3) for d in ctx.stages()}
    print("run %d: wall %.2f ms, stages %s" % (it, (t1 - t0) * 1e3, st))
w = np.linalg.eigvalsh(A)[::-1][:K]
print("max |dlam| = %.3e" % np.abs(res["values"] - w).max())
Y = res["vectors"]
print("resid = %.3e, orth = %.3e" % (np.abs(A @ Y - Y * res["values"]).max(), np.abs(Y.T @ Y - np.eye(K)).max()))

# context: the library eigensolver of the platform (cuSOLVER through torch.linalg.eigh, all s eigenpairs) on the same
# matrix, timed with CUDA events after warm-up -- not used by the product path
try:
    import torch

    Ad = torch.tensor(np.ascontiguousarray(A), device="cuda")
    for _ in range(2):
        torch.linalg.eigh(Ad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        w_t, _ = torch.linalg.eigh(Ad)
    e1.record()
    torch.cuda.synchronize()
    print("cuSOLVER (torch.linalg.eigh, full spectrum): %.2f ms per call; max |dlam| vs ours = %.3e" %
          (e0.elapsed_time(e1) / 3, np.abs(np.sort(w_t.cpu().numpy())[::-1][:K] - res["values"]).max()))
except Exception as ex:  # pragma: no cover
    print("cuSOLVER comparison unavailable:", ex)
