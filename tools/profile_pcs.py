"""Times the PCS prove path (config 1 / 4 shapes) on the GPU with per-kernel-group CUDA events."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from multilinear_b200 import api as ml
from multilinear_b200 import load
L = load()
NAMES = ["ntt_rs_encode", "merkle_leaf_subtree", "merkle_nodes", "merkle_top", "fri_fold", "sumcheck_sums", "sumcheck_fold",
         "mobius", "eq_table", "bit_reverse", "query_gather", "fused_tail", "transcript_step"]
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << nv
ml.set_device(0)
evals = ml.synthetic_elements_dev(0xB200, n)
inputs = ml.from_i64(range(5, 5 + nv))
def timed(fn, reps=3):
    fn()
    ml.synchronize()
    L.ml_profile_reset(); L.ml_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ml.synchronize()
    dt = (time.perf_counter() - t0) / reps
    L.ml_profile_enable(0)
    ks = {}
    for i, name in enumerate(NAMES):
        t, cnt, by = C.c_double(0), C.c_uint64(0), C.c_double(0)
        L.ml_profile_get(C.c_int(i), C.byref(t), C.byref(cnt), C.byref(by))
        if cnt.value:
            ks[name] = (round(t.value / reps, 3), cnt.value // reps)
    return dt * 1e3, ks
out = [None]
def claim():
    o = (C.c_uint8 * 16)()
    ml.check(L.ml_mle_evals_evaluate_dev(evals.ptr, C.c_size_t(n), C.c_void_p(inputs.ctypes.data), C.c_size_t(nv), o, None))
    out[0] = int.from_bytes(bytes(o), "little")
ms, ks = timed(claim)
print("evaluate (claim)      %8.3f ms" % ms, ks)
def prove():
    p = ml.PCSProof.prove_dev(inputs, out[0], evals, n, ml.Transcript(), None)
    return p
ms, ks = timed(prove)
print("PCSProof::prove 2^%d  %8.3f ms" % (nv, ms), ks)
p = prove()
t0 = time.perf_counter(); ok = p.verify(ml.Transcript()); print("verify", ok, "%.1f ms" % ((time.perf_counter() - t0) * 1e3))
def sumcheck():
    h = C.c_void_p()
    ml.check(L.ml_sumcheck_build_tables_for_pcs_dev(C.c_void_p(inputs.ctypes.data), C.c_size_t(nv), evals.ptr, C.c_size_t(n), None, C.byref(h)))
    s = ml.SumcheckTables(h)
    return s.compute_sumcheck_polynomials(1, ml.Transcript(), out[0])
ms, ks = timed(sumcheck)
print("sumcheck all rounds   %8.3f ms" % ms, ks)
