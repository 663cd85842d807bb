"""Host -> device upload of one 2^24-coefficient polynomial (256 MiB) from the three kinds of host memory a caller can hand to the
host-pointer entry points: library-pinned, caller-allocated + ml_host_register, plain pageable; and the cost of registering."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from multilinear_b200 import api as ml
from multilinear_b200 import load

L = load()
ml.set_device(0)
n = 1 << 24
nbytes = 16 * n
dev = ml.DeviceBuffer(nbytes)

def upload(ptr, reps=5):
    best = None
    for _ in range(reps):
        ml.synchronize()
        t0 = time.perf_counter()
        ml.check(L.ml_dev_upload(dev.ptr, C.c_void_p(ptr), C.c_size_t(nbytes)))
        ml.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return nbytes / best / 1e9, best * 1e3

out = {}
hp = C.c_void_p()
ml.check(L.ml_host_alloc_pinned(C.c_size_t(nbytes), C.byref(hp)))
C.memset(hp, 1, nbytes)
out["pinned_gbs"], out["pinned_ms"] = upload(hp.value)
a = np.ones(nbytes, dtype=np.uint8)
out["pageable_gbs"], out["pageable_ms"] = upload(a.ctypes.data)
t0 = time.perf_counter(); ml.check(L.ml_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(nbytes))); out["register_ms"] = (time.perf_counter() - t0) * 1e3
out["registered_gbs"], out["registered_ms"] = upload(a.ctypes.data)
t0 = time.perf_counter(); ml.check(L.ml_host_unregister(C.c_void_p(a.ctypes.data))); out["unregister_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); ml.check(L.ml_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(nbytes))); out["register_again_ms"] = (time.perf_counter() - t0) * 1e3
ml.check(L.ml_host_unregister(C.c_void_p(a.ctypes.data)))
print(json.dumps(out))
