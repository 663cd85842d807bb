"""Serial commit latency (one 2^log_n commit at a time, CUDA events) and launches per commit; knobs through the environment
(MLB_WALK_MIN_THREADS, MLB_TOP_MAX)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from multilinear_b200 import api as ml
from multilinear_b200 import load
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = load()
ml.set_device(0)
n = 1 << log_n
c = ml.synthetic_elements_dev(0xB200, n)
roots = None
for _ in range(3):
    f = ml.FriProverData.fold_from_coeffs_dev(c, n, ml.Transcript(), None)
    roots = f.fold_roots(); del f
torch.cuda.synchronize()
l0 = ml.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    f = ml.FriProverData.fold_from_coeffs_dev(c, n, ml.Transcript(), None)
    del f
e1.record()
torch.cuda.synchronize()
print(json.dumps({"log_n": log_n, "ms_per_commit": e0.elapsed_time(e1) / reps, "launches_per_commit": (ml.kernel_launches() - l0) / reps,
                  "walk_min_threads": os.environ.get("MLB_WALK_MIN_THREADS"), "top_max": os.environ.get("MLB_TOP_MAX"), "root0": roots[0].hex()[:16], "root_last": roots[-1].hex()[:16]}))
