import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
names = ["IADD3", "IMAD", "IMAD.WIDE", "SHF", "LOP3", "IADD3+IMAD", "SHF+IMAD", "SHF+IMAD.WIDE", "SHF+LOP3", "SHF+LOP3+IMAD"]
threads = 148 * 2048
names += ["IMAD.HI", "IMAD.HI+SHF", "DFMA", "DFMA+IMAD.WIDE", "DFMA+SHF"]
modes = list(range(10)) + [10, 11, 12, 13, 14]
for mode, name in zip(modes, names):
    ms, work = ml.microbench("pipe%d" % mode, threads, 512)
    per_clk_sm = work / (ms * 1e-3) / 148 / 1.965e9
    print("%-16s %7.3f ms  %6.1f thread-instr/clk/SM  (%.2f warp-instr/clk/SMSP)" % (name, ms, per_clk_sm, per_clk_sm / 128))
