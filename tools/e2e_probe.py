"""Diagnostic: end-to-end (host pointer) vs device-resident prove calls at several levels of concurrency."""
import ctypes as C, json, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
from multilinear_b200 import load
L = load()
ml.set_device(0)
n = 1 << 24
PEMAX = 8
coeffs = [ml.synthetic_elements_dev(0xB200 + j, n) for j in range(2)]
streams, pinned = [], []
for j in range(PEMAX):
    h = C.c_void_p(); ml.check(L.ml_stream_create(C.byref(h))); streams.append(h)
    hp = C.c_void_p(); ml.check(L.ml_host_alloc_pinned(C.c_size_t(16 * n), C.byref(hp)))
    ml.check(L.ml_dev_download(hp, coeffs[j % 2].ptr, C.c_size_t(16 * n))); pinned.append(hp)
ml.synchronize()
def worker(j, steps, dev, times):
    ml.set_device(0)
    ml.check(L.ml_set_thread_stream(streams[j], C.c_int(1)))
    for _ in range(steps):
        t = ml.Transcript(); h = C.c_void_p(); t0 = time.perf_counter()
        if dev: ml.check(L.ml_rs_fri_prove_dev(coeffs[j % 2].ptr, C.c_size_t(n), t.h, streams[j], C.byref(h)))
        else: ml.check(L.ml_rs_fri_prove(pinned[j], C.c_size_t(n), t.h, C.byref(h)))
        times.append(time.perf_counter() - t0)
        p = ml.FriProof(h); p.serialize(); del p
def run(pe, steps, dev):
    times = []
    ts = [threading.Thread(target=worker, args=(j, steps, dev, times)) for j in range(pe)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    ml.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3, sum(times) / len(times) * 1e3
run(PEMAX, 1, False)
for dev in (True, False, False, False):
    for pe in ((1, 2, 4, 8) if dev else (4, 8, 8, 8)):
        step, call = run(pe, 3, dev)
        print("dev=%d PE=%d  step %.1f ms  (%.1f ms per commit)  call mean %.1f ms" % (dev, pe, step, step / pe, call), flush=True)
