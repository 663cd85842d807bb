"""RS-encode NTT 2^log_n -> 2^(log_n+1) timed with the library's CUDA-event brackets (ml_profile_*); variants through env:
MLB_NTT_NO_WTAB=1 (inter-pass twiddles from the two-level tables instead of the N-entry matrix)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
from multilinear_b200 import load

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = load()
ml.set_device(0)
n = 1 << log_n
c = ml.synthetic_elements_dev(0xB200, n)
out = ml.DeviceBuffer(32 * n)
g = ml.pow_2_generator(log_n + 1)
gb = (C.c_uint8 * 16)(*g.to_bytes(16, "little"))
for _ in range(3):
    ml.check(L.ml_reed_solomon_dev(c.ptr, C.c_size_t(n), gb, out.ptr, None))
ml.synchronize()
L.ml_profile_reset(); L.ml_profile_enable(1)
for _ in range(reps):
    ml.check(L.ml_reed_solomon_dev(c.ptr, C.c_size_t(n), gb, out.ptr, None))
ml.synchronize()
L.ml_profile_enable(0)
t, cnt, by = C.c_double(0), C.c_uint64(0), C.c_double(0)
L.ml_profile_get(C.c_int(0), C.byref(t), C.byref(cnt), C.byref(by))
ms = t.value / cnt.value
print(json.dumps({"workload": "rs_encode_ntt", "log_n": log_n, "variant": "no_wtab" if os.environ.get("MLB_NTT_NO_WTAB") else "wtab",
                  "ms": ms, "alg_gbs": by.value / cnt.value / (ms * 1e-3) / 1e9, "reps": int(cnt.value)}))
