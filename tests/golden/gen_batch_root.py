"""Generates tests/golden/batch_root_64x2p22.json: the Merkle::batch_commit root (src/merkle_tree/mod.rs:92-131) over the RS
codes of bench.py's batched workload (BASELINE configs[4]: 64 polynomials of 2^22 evaluations, synthetic seeds 5000 + j),
computed by the CPU oracle.  About two minutes on 8 cores; ~10 GiB of host memory.

    python tests/golden/gen_batch_root.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def batch_root(O, seeds, n, n_ranges=8):
    g = O.pow2_generator((n.bit_length() - 1) + 1)
    codes = [O.reed_solomon(O.bit_reverse(O.to_coefficient(O.synthetic(s, n))), g) for s in seeds]
    rows = n // n_ranges
    roots = []
    for r in range(n_ranges):
        datas = [np.concatenate([c[r * rows:(r + 1) * rows], c[n + r * rows:n + (r + 1) * rows]], axis=1) for c in codes]
        roots.append(O.merkle_batch_commit(datas).root())
    subtree_roots = [r.hex() for r in roots]
    while len(roots) > 1:
        roots = [hashlib.sha256(roots[i] + roots[i + 1]).digest() for i in range(0, len(roots), 2)]
    return roots[0], subtree_roots


if __name__ == "__main__":
    from oracle import binding
    binding.build()
    O = binding.Oracle(threads=os.cpu_count() or 1)
    log_n, B = 22, 64
    root, sub = batch_root(O, [5000 + j for j in range(B)], 1 << log_n)
    out = {"workload": "batched_commit_64x2^22", "n_polys": B, "log_n": log_n, "seeds": "5000 + j", "root": root.hex(),
           "subtree_roots_8_ranges": sub, "generator": "tests/golden/gen_batch_root.py (CPU oracle, oracle/oracle.c)"}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "batch_root_64x2p22.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
