import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
threads = 148 * 2048
names = ["m0", "m2", "m2_r1", "m2_r11", "m2_r111", "m2_r101", "m2_r21", "m2_r121", "m2_r211", "m2_r221", "m0_r1", "m0_r11", "m0_r111", "m0_r121", "m0_r211"]
for nm in names:
    ms, w = ml.microbench("sha_leaf_" + nm, threads, 64)
    ms2, w2 = ml.microbench("sha_node_" + nm, threads, 64)
    print("%-10s leaf %.3e/s  node %.3e/s" % (nm, w / ms * 1e3, w2 / ms2 * 1e3))
