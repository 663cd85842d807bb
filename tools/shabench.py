import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
threads = 148 * 2048
for m in (0, 63, 3, 1, 2, 19, 11, 27, 59, 43):
    ms, w = ml.microbench("sha_leaf_m%d" % m, threads, 64)
    ms2, w2 = ml.microbench("sha_node_m%d" % m, threads, 64)
    print("mask %2d  leaf %.3e/s  node %.3e/s" % (m, w / ms * 1e3, w2 / ms2 * 1e3))
