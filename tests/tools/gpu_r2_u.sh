set -x
timeout 1200 python -m pytest tests/test_gpu_cc.py tests/test_gpu_dropin.py -x -q -m gpu --tb=short 2>&1 | tail -5
python - <<'PY'
import time, numpy as np
from squid_b200 import api
from tests.test_gpu_cc import by_smallest_node
rng = np.random.default_rng(1)
for n, m in ((2_000_000, 1_500_000), (2_000_000, 6_000_000)):
    a = rng.integers(0, n, m).astype(np.int32); b = rng.integers(0, n, m).astype(np.int32)
    api.ConnectedComponent(n, a, b)
    t = time.perf_counter(); lab = api.ConnectedComponent(n, a, b); t1 = time.perf_counter() - t
    t = time.perf_counter(); ref = by_smallest_node(n, a, b); t2 = time.perf_counter() - t
    print("cc n=%d m=%d components=%d: device call %.2f ms (host arrays in and out), scipy on one core %.1f ms, equal=%s" % (n, m, lab.max() + 1, 1e3 * t1, 1e3 * t2, np.array_equal(lab, ref)))
PY
