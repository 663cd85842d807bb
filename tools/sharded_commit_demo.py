"""Config 5 on N GPUs (torchrun): batched commit of B polynomials of 2^v evaluations, sharded by polynomial for the
encode and by leaf range for the hashing (one NCCL all-to-all + one 32-byte all-gather).  Every rank checks the
batch root against a single-GPU computation of the same tree on rank-local data."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torch.distributed as dist
from multilinear_b200 import api as ml
from multilinear_b200 import load
from multilinear_b200.sharded import CudaBackend, sharded_batch_commit

v = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["serial", "pipelined", "p2p"]
n = 1 << v
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
ml.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = load()
be = CudaBackend()

def poly(j):
    t = torch.empty(16 * n, dtype=torch.uint8, device="cuda")
    ml.check(L.ml_synthetic_elements_dev(C.c_uint64(5000 + j), C.c_size_t(n), C.c_void_p(t.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return t

mine = [poly(j) for j in range(B) if j % world == rank]
d = dist if world > 1 else None
# single-GPU reference of the same tree (all polynomials on this GPU), only for moderate sizes
want = None
if B * n <= (1 << 26):
    allp = [poly(j) for j in range(B)]
    want = sharded_batch_commit(allp, n, B, be, None)
    del allp
for mode in modes:
    kw = {}
    if mode.startswith("p2p:"):  # p2p:<max CTAs of the store pass>
        kw["p2p_ctas"] = int(mode.split(":")[1])
        mode_name, mode = mode, "p2p"
    else:
        mode_name = mode
    root = sharded_batch_commit(mine, n, B, be, d, mode=mode, **kw)  # warm-up (tables, pool, peer mappings)
    torch.cuda.synchronize()
    if d: d.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        root = sharded_batch_commit(mine, n, B, be, d, mode=mode, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if d:
        t = torch.tensor([ms], device="cuda"); d.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
    if rank == 0:
        print(json.dumps({"workload": "batched_commit", "mode": mode_name, "polys": B, "log_n": v, "n_gpus": world, "ms": ms,
                          "melem_per_s": B * n / (ms * 1e-3) / 1e6, "root": root.hex(),
                          "matches_single_gpu": None if want is None else want == root}), flush=True)
be.release_peer_buffers()
if d:
    d.barrier(); d.destroy_process_group()
