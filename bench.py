#!/usr/bin/env python
"""bench.py — PCS commit throughput on B200 (BASELINE.json metric), one process per GPU.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): commit to a polynomial of
2^24 coefficients = reed_solomon (RS-encode NTT to 2^25, src/fri/mod.rs:19-28) + FriProverData::fold
(Merkle commit of every FRI layer, transcript challenges, all 24 folds; src/fri/mod.rs:136-145).
A "step" is one such commit on synthetic random field elements.

  value : Melem/s = n_gpus * 2^24 / t, inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e   : the same metric through the host-pointer C ABI (ml_rs_fri_prove: pinned host coefficients -> H2D ->
          RS-encode + FriProof::prove incl. the 128 query openings -> proof back on the host)
  --impl reference : the CPU oracle (C restatement of the reference; the Rust crate cannot be built here)
          on all host threads, same metric, bounded sample.

N > 1: launched by torchrun; every rank commits its own polynomial (weak scaling, independent units, no
data-path collective); only timing is reduced over NCCL.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pcs_commit_melem_per_s"
UNIT = "Melem/s"
PROF_NAMES = ["ntt_rs_encode", "merkle_leaf_subtree", "merkle_nodes", "merkle_top", "fri_fold", "sumcheck_sums", "sumcheck_fold",
              "mobius", "eq_table", "bit_reverse", "query_gather"]


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_commit_rate(oracle_mod, log_n, threads, steps, warmup, seed=0xB200):
    """Melem/s of the CPU oracle for reed_solomon + FriProverData::fold at 2^log_n coefficients"""
    from oracle.binding import fe_ints
    O = oracle_mod.Oracle(threads=threads)
    n = 1 << log_n
    coeffs = O.synthetic(seed, n)
    gp = O.pow2_generator_powers(log_n + 1)  # the reference's callers pass gen_pows in (src/fri/mod.rs:136)
    gen = fe_ints(gp[1:2])[0]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        code = O.reed_solomon(coeffs, gen)
        f, st = O.fri_fold(gp, code, O.transcript())
        dt = time.perf_counter() - t0
        assert st == 0 and f.last_element() is not None
        del f
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return n / mean / 1e6, mean


def pick_cpu_sample(oracle_mod, threads, total_steps, budget_s):
    """largest log_n in {20, 22, 24} whose (warmup+steps) run fits the time budget, from one timed 2^18 commit"""
    _, t18 = cpu_commit_rate(oracle_mod, 18, threads, 1, 0)
    for log_n in (24, 22, 20):
        est = t18 * (1 << (log_n - 18)) * (log_n + 1) / 19.0
        if est * total_steps <= budget_s:
            return log_n
    return 18


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host threads; rank 0 only"""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import binding
    binding.build()
    threads = os.cpu_count() or 1
    log_n = pick_cpu_sample(binding, threads, args.steps + args.warmup, 150.0)
    rate, mean = cpu_commit_rate(binding, log_n, threads, args.steps, args.warmup)
    sample = "reed_solomon + FriProverData::fold of one 2^%d-coefficient polynomial per step (workload is 2^%d)" % (log_n, args.log_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u128", "data": "synthetic",
        "config": {"workload": "pcs_commit_rs_merkle_fri_fold", "log_n": args.log_n, "sample_log_n": log_n, "blowup": 2,
                   "note": "CPU oracle = C restatement of the reference (Rust toolchain and winter-math/sha2 sources unavailable)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    from multilinear_b200 import api as ml
    from multilinear_b200 import load
    L = load()
    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ml.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = 1 << args.log_n
    N = 2 * n
    coeffs = ml.synthetic_elements_dev(0xB200 + rank, n)  # resident in HBM before the timed region
    ml.synchronize()

    def step():
        f = ml.FriProverData.fold_from_coeffs_dev(coeffs, n, ml.Transcript(), None)
        roots, last = f.fold_roots(), f.last_element
        del f  # the handle (all layers, ~3.5 GB at 2^24) returns to the stream-ordered pool
        return roots, last

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    L.ml_profile_enable(0)
    for _ in range(args.warmup):
        roots, last = step()
    L.ml_profile_reset()
    L.ml_profile_enable(1)
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = ml.kernel_launches()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        roots, last = step()
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1) / args.steps
    launches = ml.kernel_launches() - launches0
    L.ml_profile_enable(0)

    # per-kernel device time inside the timed region
    kernels = {}
    for i, name in enumerate(PROF_NAMES):
        t, cnt, by = C.c_double(0), C.c_uint64(0), C.c_double(0)
        L.ml_profile_get(C.c_int(i), C.byref(t), C.byref(cnt), C.byref(by))
        if cnt.value:
            kernels[name] = {"ms_per_step": t.value / args.steps, "launches_per_step": cnt.value / args.steps,
                             "alg_bytes_per_step": by.value / args.steps,
                             "achieved_gbs": by.value / (t.value * 1e-3) / 1e9 if t.value > 0 else None}
    L.ml_profile_reset()

    # ---- e2e through the host-pointer C ABI (pinned host input, proof back on the host)
    e2e_steps = max(1, min(args.steps, 5))
    pinned = C.c_void_p()
    ml.check(L.ml_host_alloc_pinned(C.c_size_t(16 * n), C.byref(pinned)))
    ml.check(L.ml_dev_download(pinned, coeffs.ptr, C.c_size_t(16 * n)))

    def e2e_step():
        t = ml.Transcript()
        h = C.c_void_p()
        ml.check(L.ml_rs_fri_prove(pinned, C.c_size_t(n), t.h, C.byref(h)))
        proof = ml.FriProof(h)
        blob = proof.serialize()
        return proof, len(blob)

    proof, blob_len = e2e_step()  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        proof, blob_len = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_ok = proof.commitments == roots and proof.last_elem == last
    L.ml_host_free_pinned(pinned)

    # max over ranks
    if dist is not None:
        tt = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        ll = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ll)
        launches = int(ll[0])

    if rank == 0:
        value = world * n / (ms * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        dom = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"]) if kernels else (None, None)
        roofline = None
        if dom[0]:
            k = dom[1]
            per_launch_bytes = k["alg_bytes_per_step"] / k["launches_per_step"]
            per_launch_ms = k["ms_per_step"] / k["launches_per_step"]
            ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": None, "peak_source": peak_src, "share_of_step": k["ms_per_step"] / ms,
                        "note": "SHA-256 hashing is integer-pipe (alu) bound, not HBM bound; see int_pipe and profiles/"}
        ntt = kernels.get("ntt_rs_encode")
        extra = {}
        if ntt:
            extra["ntt_roofline"] = {"bound": "hbm", "achieved": ntt["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                                     "frac": ntt["achieved_gbs"] / hbm_peak, "alg_bytes": ntt["alg_bytes_per_step"], "ms": ntt["ms_per_step"]}
        # integer-pipe speed of light measured on this GPU (no HBM traffic)
        try:
            sol = {}
            for what, nn, it in (("sha_leaf", 148 * 2048, 32), ("sha_node", 148 * 2048, 32), ("butterfly", 148 * 2048, 128)):
                mms, work = ml.microbench(what, nn, it)
                sol[what + "_per_s"] = work / (mms * 1e-3)
            extra["int_pipe"] = sol
            if dom[0] == "merkle_leaf_subtree":
                leaves = 2.0 * n  # sum over the fold chain of leaves handled by the leaf/subtree kernel
                ideal_ms = (leaves / sol["sha_leaf_per_s"] + leaves * 0.875 / sol["sha_node_per_s"]) * 1e3
                roofline["int_pipe_frac"] = ideal_ms / dom[1]["ms_per_step"]
        except Exception as e:  # noqa: BLE001
            extra["int_pipe_error"] = str(e)

        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding
            binding.build()
            threads = os.cpu_count() or 1
            log_s = pick_cpu_sample(binding, threads, 2, 25.0)
            rate, mean = cpu_commit_rate(binding, log_s, threads, 1, 1)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": "reed_solomon + FriProverData::fold at 2^%d coefficients, %.2f s per commit, all host threads (OpenMP)" % (log_s, mean)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u128", "data": "synthetic",
            "config": {"workload": "pcs_commit_rs_merkle_fri_fold", "log_n": args.log_n, "blowup": 2, "polys_per_gpu": 1,
                       "l2": "inputs larger than L2 (256 MiB coefficients, 512 MiB code per step)", "parallelism": "independent commits per GPU"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": world * n / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": blob_len,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "includes": "128 query openings + proof serialisation", "matches_device_run": bool(e2e_ok)},
            "gpu_launches": launches, "clocks": sampler.result(), "kernels": kernels,
        }
        line.update(extra)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24, dest="log_n")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
