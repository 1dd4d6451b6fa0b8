"""Round-trip bandwidth of the library's host copy path from pageable memory (gpurun helper, not a pytest file):
python tests/bench_hostcopy.py [MB]; FLGP_HC_LANES / FLGP_HC_CHUNK_MB / FLGP_NO_STAGED_COPY select the variant."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flgp_b200 as F  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 240
ctx = F.default_ctx()
src = np.random.default_rng(0).integers(0, 1 << 62, size=mb * (1 << 20) // 8, dtype=np.int64)
ctx.copy_roundtrip(src)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    out = ctx.copy_roundtrip(src)
    best = min(best, time.perf_counter() - t0)
assert np.array_equal(out, src)
print("lanes=%s chunk=%s MB staged=%s: %d MB up + down in %.2f ms = %.1f GB/s" % (
    os.environ.get("FLGP_HC_LANES", "default"), os.environ.get("FLGP_HC_CHUNK_MB", "4"),
    "no" if os.environ.get("FLGP_NO_STAGED_COPY") else "yes", mb, best * 1e3, 2 * mb / 1024 / best))
