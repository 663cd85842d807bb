"""Small end-to-end case for compute-sanitizer: sharded prove (4 virtual ranks), width-w sumcheck chain, PCS prove, pageable upload."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from multilinear_b200 import api as ml
from oracle.binding import Oracle, fe_arr
ml.set_device(0)
O = Oracle(threads=4)
nv, B = 8, 8
polys = [O.synthetic(100 + j, 1 << nv) for j in range(B)]
inp = O.synthetic(7, nv)
outs = fe_arr([O.mle_evals_evaluate(p, inp) for p in polys])
sh = ml.ShardedBatchedProver.single_process([0] * 4, B, nv)
t, ot = ml.Transcript(), O.transcript()
proof = sh.prove(inp, outs, polys, t)
op, st = O.batched_pcs_prove(inp, outs, polys, ot)
assert st == 0 and proof.fri_proof.serialize() == op.fri.blob and t.random() == ot.random()
sh.free()
ev = O.synthetic(5, 1 << 13)
i2 = O.synthetic(6, 13)
out = O.mle_evals_evaluate(ev, i2)
p = ml.PCSProof.prove(i2, out, ev, ml.Transcript())
assert p.verify(ml.Transcript()) == 0
w = 3
matrix = O.synthetic(31, w << 13)
g = ml.WideSumcheckTables.build(i2, matrix, w)
o = O.wsumcheck_build(i2, matrix, w)
terms = [(7, [0, 1]), (5, [2]), (3, [])]
g.set_composition(terms); o.set_composition(terms)
assert g.compute_sumcheck_polynomials(2, ml.Transcript(), 11) == o.compute_sumcheck_polynomials(2, O.transcript(), 11)
big = O.synthetic(9, 1 << 21)  # 32 MiB pageable: staged upload path
pr = ml.FriProof.prove_from_coeffs(big, ml.Transcript())
assert pr.verify() == 0
print("sanitize case ok")
