import sys, ctypes as C
sys.path.insert(0, '/root/repo')
from squid_b200 import api
L = api.lib()
L.sqg_selftest_gpu_sort.argtypes = [C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_float)]
for n in (100000, 1070000, 4000000):
    for rng in (n // 3, 1 << 40):
        for it in range(3):
            ms = C.c_float(-1)
            v = L.sqg_selftest_gpu_sort(0, n, 11 + it, rng, 0, C.byref(ms))
        print("n", n, "range", rng, "verdict", v, "device ms %.2f" % ms.value, flush=True)
