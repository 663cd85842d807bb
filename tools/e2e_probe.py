"""Diagnostic: end-to-end (host pointer) vs device-resident prove calls at several levels of concurrency."""
import ctypes as C, json, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
MODE = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "torch" in MODE:
    import torch
    torch.cuda.init(); _x = torch.zeros(1, device="cuda"); torch.cuda.synchronize()
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
if "prelude" in MODE:  # what bench.py does first: 6 device-resident commits in flight, 10 steps
    def dev_worker(j):
        ml.set_device(0)
        for _ in range(10):
            f = ml.FriProverData.fold_from_coeffs_dev(coeffs[j % 2], n, ml.Transcript(), streams[j].value); f.fold_roots(); del f
    ts = [threading.Thread(target=dev_worker, args=(j,)) for j in range(6)]
    for t in ts: t.start()
    for t in ts: t.join()
    ml.synchronize()
if "events" in MODE:
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
run(PEMAX, 1, False)
print("MODE", MODE)
L.ml_trace_dump.restype = None
for dev in (False, False):
    for pe in (8, 8, 8):
        if os.environ.get("MLB_TRACE"): L.ml_trace_dump(); print("---- run", flush=True)
        step, call = run(pe, 3, dev)
        print("dev=%d PE=%d  step %.1f ms  (%.1f ms per commit)  call mean %.1f ms" % (dev, pe, step, step / pe, call), flush=True)
