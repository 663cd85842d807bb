"""Secondary BASELINE configs (parity-test cases, not the bench line): device-resident timings with CUDA events.
  config 2: reed_solomon 2^20 -> 2^21 then intt 2^21
  config 4: sumcheck over 2^24 evaluations, all 24 rounds with transcript challenges
  config 1: PCSProof::prove at n_vars = 20 (the reference's own test inputs)
Prints one JSON object; the CPU column is the oracle with all host threads on the same inputs."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from multilinear_b200 import api as ml
from multilinear_b200 import load
from oracle.binding import Oracle, fe_arr
L = load()
ml.set_device(0)
O = Oracle(threads=os.cpu_count() or 1)
res = {}

def gpu_time(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def cpu_time(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3

# ---- config 2
n = 1 << 20
coeffs = ml.synthetic_elements_dev(7, n)
code = ml.DeviceBuffer(32 * n)
back = ml.DeviceBuffer(32 * n)
g21 = ml.pow_2_generator(21)
gb = (C.c_uint8 * 16)(*g21.to_bytes(16, "little"))
def rs():
    ml.check(L.ml_reed_solomon_dev(coeffs.ptr, C.c_size_t(n), gb, code.ptr, None))
def inv():
    ml.check(L.ml_intt_dev(code.ptr, C.c_size_t(2 * n), gb, back.ptr, None))
t_rs, t_inv = gpu_time(rs), gpu_time(inv)
hc = coeffs.elems()
assert (back.elems()[:n] == hc).all() and not back.elems()[n:].any()  # intt(reed_solomon(c)) == c padded with zeros
c_rs = cpu_time(lambda: O.reed_solomon(hc, g21))
hcode = code.elems()
c_inv = cpu_time(lambda: O.intt(hcode, g21))
res["config2_ntt_2p20_blowup2"] = {"gpu_rs_encode_ms": t_rs, "gpu_intt_ms": t_inv, "cpu_rs_encode_ms": c_rs, "cpu_intt_ms": c_inv,
                                   "rs_alg_gbs": 48 * n / t_rs / 1e6, "intt_alg_gbs": 32 * 2 * n / t_inv / 1e6, "roundtrip_exact": True}
# ---- config 4
nv = 24
n = 1 << nv
evals = ml.synthetic_elements_dev(0xB200, n)
inputs = ml.from_i64(range(5, 5 + nv))
o = (C.c_uint8 * 16)()
ml.check(L.ml_mle_evals_evaluate_dev(evals.ptr, C.c_size_t(n), C.c_void_p(inputs.ctypes.data), C.c_size_t(nv), o, None))
claim = int.from_bytes(bytes(o), "little")
def sumcheck():
    h = C.c_void_p()
    ml.check(L.ml_sumcheck_build_tables_for_pcs_dev(C.c_void_p(inputs.ctypes.data), C.c_size_t(nv), evals.ptr, C.c_size_t(n), None, C.byref(h)))
    return ml.SumcheckTables(h).compute_sumcheck_polynomials(1, ml.Transcript(), claim)
t_sc = gpu_time(sumcheck, reps=5)
nv_c = 20  # CPU oracle sample (its eq-table build is O(n v) like the reference)
ev_c = O.synthetic(0xB200, 1 << nv_c)
in_c = fe_arr(range(5, 5 + nv_c))
cl_c = O.mle_evals_evaluate(ev_c, in_c)
c_sc = cpu_time(lambda: O.sumcheck_build(in_c, ev_c).compute_sumcheck_polynomials(1, O.transcript(), cl_c))
res["config4_sumcheck_2p24"] = {"gpu_ms_2p24_incl_table_build": t_sc, "gpu_melem_per_s": n / t_sc / 1e3,
                                "cpu_ms_2p20": c_sc, "cpu_melem_per_s": (1 << nv_c) / c_sc / 1e3, "alg_gbs": 112 * n / t_sc / 1e6}
# ---- config 1
nv = 20
n = 1 << nv
ev1 = ml.from_i64([7 * i + 3 for i in range(n)])
in1 = ml.from_i64(range(nv))
out1 = ml.MultilinearPolynomialEvals(ev1).evaluate(ml.to_ints(in1))
d1 = ml.DeviceBuffer.from_host(ev1)
t_p = gpu_time(lambda: ml.PCSProof.prove_dev(in1, out1, d1, n, ml.Transcript(), None), reps=5)
t0 = time.perf_counter()
op, st = O.pcs_prove(in1, out1, ev1, O.transcript())
c_p = (time.perf_counter() - t0) * 1e3
gp = ml.PCSProof.prove_dev(in1, out1, d1, n, ml.Transcript(), None)
res["config1_pcs_prove_nvars20"] = {"gpu_ms": t_p, "cpu_ms": c_p, "bit_exact_vs_oracle": gp.fri_proof.serialize() == op.fri.blob,
                                    "verifies": gp.verify(ml.Transcript()) == 0}
# ---- the reference's sumcheck_high_bench (src/constraint_system/sumcheck.rs:368-398): pythagorean trace, width 4, 2^20 rows
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from test_oracle import pythagorean_system
M_ = 2**128 - 45 * 2**40 + 1
def wide(log_h, use_gpu):
    matrix, row_point, terms, t = pythagorean_system(O, log_h)
    if use_gpu:
        g = ml.WideSumcheckTables.build(row_point, matrix, 4); g.set_composition(terms)
        tt = ml.Transcript()
        f = lambda: None
        t0 = time.perf_counter(); out = g.compute_sumcheck_polynomials(2, tt, 0); dt = time.perf_counter() - t0
    else:
        o = O.wsumcheck_build(row_point, matrix, 4); o.set_composition(terms)
        t0 = time.perf_counter(); out = o.compute_sumcheck_polynomials(2, t, 0); dt = time.perf_counter() - t0
    return dt * 1e3, out
wide(12, True)  # warm-up
g_ms, g_out = wide(20, True)
c_ms, c_out = wide(16, False)
g16_ms, g16_out = wide(16, True)
res["sumcheck_high_bench_2p20_x4"] = {"gpu_ms_2p20_proof_only": g_ms, "gpu_ms_2p16": g16_ms, "cpu_ms_2p16_single_thread": c_ms,
                                      "bit_exact_vs_oracle_2p16": g16_out == c_out}
res["cpu_threads"] = os.cpu_count()
print(json.dumps(res, indent=1))
