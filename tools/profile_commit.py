"""One warm-up + one commit at 2^24 (reed_solomon + FriProverData::fold) — the target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
ml.set_device(0)
coeffs = ml.synthetic_elements_dev(0xB200, n)
for _ in range(2):
    f = ml.FriProverData.fold_from_coeffs_dev(coeffs, n, ml.Transcript(), None)
    print(f.fold_roots()[0].hex(), f.last_element)
    del f
ml.synchronize()
