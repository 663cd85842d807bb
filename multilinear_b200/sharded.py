"""Batched commit sharded over GPUs (BASELINE config 5).

`Merkle::batch_commit` over B codes (src/merkle_tree/mod.rs:110-131, src/fri/batched_fri.rs:62-77): leaf i hashes the
ReedSolomonPairs (code_j[i], code_j[i+N/2]) of ALL codes j, so the work re-partitions from "by polynomial" (encoding)
to "by leaf range" (hashing):

  1. rank g encodes its polynomials {j : j mod G == g}                      (no communication)
  2. all-to-all: rank g receives rows [g*Lr, (g+1)*Lr) of every code        (the only data-path collective)
  3. rank g hashes its Lr batched leaves and reduces its subtree             (no communication)
  4. all-gather of the G subtree roots (32 B each); every rank hashes the top log2(G) levels

One process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in the CPU tests).  The compute steps go
through a backend object: `CudaBackend` calls the C ABI on device tensors; the tests inject a CPU backend to check
the exchange logic with world_size 2.
"""
import ctypes as C

import torch


class CudaBackend:
    """compute steps through libmultilinear_b200.so on cuda tensors (uint8)"""

    def __init__(self):
        from . import api
        from ._lib import load
        self.api, self.L = api, load()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def empty(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def encode(self, evals_t, n):
        code = self.empty(32 * n)
        self.api.check(self.L.ml_pcs_encode_dev(C.c_void_p(evals_t.data_ptr()), C.c_size_t(n), C.c_void_p(code.data_ptr()), self._stream()))
        return code

    def pack_pairs(self, code_t, n_code, n_ranks, n_local, pl, send_t):
        self.api.check(self.L.ml_pack_pairs_dev(C.c_void_p(code_t.data_ptr()), C.c_size_t(n_code), C.c_size_t(n_ranks), C.c_size_t(n_local),
                                                 C.c_size_t(pl), C.c_void_p(send_t.data_ptr()), self._stream()))

    def leaf_subtree_root(self, recv_t, ptr_offsets, rows):
        import numpy as np
        ptrs = (C.c_void_p * len(ptr_offsets))(*[recv_t.data_ptr() + o for o in ptr_offsets])
        out = np.empty(32, dtype=np.uint8)
        self.api.check(self.L.ml_batched_leaf_subtree_dev(ptrs, C.c_size_t(len(ptr_offsets)), C.c_size_t(rows), self._stream(),
                                                           C.c_void_p(out.ctypes.data)))
        return torch.from_numpy(out.copy())

    def top(self, roots_bytes, n_roots):
        import numpy as np
        buf = np.frombuffer(roots_bytes, dtype=np.uint8).copy()
        out = np.empty(32, dtype=np.uint8)
        self.api.check(self.L.ml_merkle_top_from_roots(C.c_void_p(buf.ctypes.data), C.c_size_t(n_roots), C.c_void_p(out.ctypes.data)))
        return out.tobytes()


def owner(j, world):
    return j % world


def sharded_batch_commit(local_evals, n, n_polys, backend, dist=None):
    """local_evals: this rank's polynomials (uint8 tensors of 16*n bytes), in increasing global index.
    Returns the batch root (identical on every rank)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_code, leaves = 2 * n, n
    assert n_polys % world == 0 and leaves % world == 0, "polynomials and leaves must split evenly over ranks"
    n_local, rows = n_polys // world, leaves // world
    assert len(local_evals) == n_local
    chunk = rows * 32
    send = backend.empty(world * n_local * chunk)
    for pl, ev in enumerate(local_evals):  # phase 1 + packing in exchange order
        code = backend.encode(ev, n)
        backend.pack_pairs(code, n_code, world, n_local, pl, send)
    if world > 1:  # phase 2
        recv = backend.empty(world * n_local * chunk)
        dist.all_to_all_single(recv, send)
    else:
        recv = send
    # recv layout: [src rank][src's local poly][rows][32]; global poly j = src + pl*world -> leaf hash order j = 0..B-1
    offsets = [((owner(j, world) * n_local + j // world) * chunk) for j in range(n_polys)]
    root = backend.leaf_subtree_root(recv, offsets, rows)  # phase 3
    if world == 1:
        return bytes(root.numpy().tobytes())
    gathered = [torch.empty(32, dtype=torch.uint8) for _ in range(world)]  # phase 4 (32 bytes per rank)
    if recv.is_cuda:
        g_dev = [torch.empty(32, dtype=torch.uint8, device=recv.device) for _ in range(world)]
        dist.all_gather(g_dev, root.to(recv.device))
        gathered = [g.cpu() for g in g_dev]
    else:
        dist.all_gather(gathered, root)
    return backend.top(b"".join(bytes(g.numpy().tobytes()) for g in gathered), world)
