"""Integer-pipe speed-of-light numbers on the attached B200 (ml_microbench)."""
import json
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml

res = {}
threads = 148 * 2048
for what, n, iters in (("modmul", threads, 256), ("butterfly", threads, 256), ("sha_leaf_alu", threads, 64), ("sha_node_alu", threads, 64),
                       ("sha_leaf_fma", threads, 64), ("sha_node_fma", threads, 64), ("copy", 1 << 30, 1)):
    ms, work = ml.microbench(what, n, iters)
    res[what] = {"ms": ms, "work": work, "per_s": work / (ms * 1e-3)}
    print("%-14s %8.3f ms  %.3e /s" % (what, ms, work / (ms * 1e-3)))
print(json.dumps(res))
