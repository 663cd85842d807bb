"""Batched commit sharded over GPUs (BASELINE config 5).

`Merkle::batch_commit` over B codes (src/merkle_tree/mod.rs:110-131, src/fri/batched_fri.rs:62-77): leaf i hashes the
ReedSolomonPairs (code_j[i], code_j[i+N/2]) of ALL codes j, so the work re-partitions from "by polynomial" (encoding)
to "by leaf range" (hashing):

  1. rank g encodes its polynomials {j : j mod G == g}                      (no communication)
  2. exchange: rank g receives rows [g*Lr, (g+1)*Lr) of every code          (the only data-path exchange)
  3. rank g hashes its Lr batched leaves and reduces its subtree             (no communication)
  4. all-gather of the G subtree roots (32 B each); every rank hashes the top log2(G) levels

Three ways to run step 2 (`mode`):
  "serial"    one all-to-all after every local polynomial is encoded (the simplest schedule; CPU/gloo tests use it)
  "pipelined" one all-to-all per polynomial on a side stream, hidden behind the next polynomial's NTT
  "p2p"       no collective at all: the pack pass stores each pair straight into the owner's receive buffer through
              NVLink peer mappings (CUDA IPC), also on a side stream behind the next NTT; one barrier before hashing.
              The store pass runs on a small grid (p2p_ctas, default 32 CTAs): it is NVLink-latency bound, and a wide
              grid only takes SM slots from the NTT it hides behind (2 GPUs, 64 x 2^22: 29.8 ms at 32 CTAs, 33.3 ms at
              1184; NCCL pipelined 30.5 ms, serial 33.4 ms; profiles/r1_sharded_batched_commit.txt)

One process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in the CPU tests).  The compute steps go
through a backend object: `CudaBackend` calls the C ABI on device tensors; the tests inject a CPU backend to check
the exchange logic with world_size 2.
"""
import ctypes as C

import torch


class CudaBackend:
    """compute steps through libmultilinear_b200.so on cuda tensors (uint8)"""

    is_cuda = True

    def __init__(self):
        from . import api
        from ._lib import load
        self.api, self.L = api, load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._side = None
        self._peer = None  # (key, own_ptr, [mapped base per rank], rank)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def encode_streams(self):
        """two streams for the per-polynomial encodes: consecutive polynomials overlap, so the partial last wave of one NTT
        pass (2^23 points = 4.6 waves of CTAs) is filled by the next polynomial's kernels"""
        if getattr(self, "_enc2", None) is None:
            self._enc2 = torch.cuda.Stream(device=self.device)
        return [torch.cuda.current_stream(), self._enc2]

    def empty(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def encode(self, evals_t, n, out=None):
        code = self.empty(32 * n) if out is None else out
        self.api.check(self.L.ml_pcs_encode_dev(C.c_void_p(evals_t.data_ptr()), C.c_size_t(n), C.c_void_p(code.data_ptr()), self._stream()))
        return code

    def pack_pairs(self, code_t, n_code, n_ranks, n_local, pl, send_t):
        self.api.check(self.L.ml_pack_pairs_dev(C.c_void_p(code_t.data_ptr()), C.c_size_t(n_code), C.c_size_t(n_ranks), C.c_size_t(n_local),
                                                 C.c_size_t(pl), C.c_void_p(send_t.data_ptr()), self._stream()))

    def leaf_subtree_root(self, recv_t, ptr_offsets, rows):
        return self.leaf_subtree_root_ptr(recv_t.data_ptr(), ptr_offsets, rows)

    def leaf_subtree_root_ptr(self, base_ptr, ptr_offsets, rows):
        """subtree root as a 32-byte device tensor (feeds the all-gather without a host round trip)"""
        ptrs = (C.c_void_p * len(ptr_offsets))(*[base_ptr + o for o in ptr_offsets])
        root = self.empty(32)
        self.api.check(self.L.ml_batched_leaf_subtree_root_dev(ptrs, C.c_size_t(len(ptr_offsets)), C.c_size_t(rows),
                                                                C.c_void_p(root.data_ptr()), self._stream()))
        return root

    def top(self, roots_bytes, n_roots):
        import numpy as np
        buf = np.frombuffer(roots_bytes, dtype=np.uint8).copy()
        out = np.empty(32, dtype=np.uint8)
        self.api.check(self.L.ml_merkle_top_from_roots(C.c_void_p(buf.ctypes.data), C.c_size_t(n_roots), C.c_void_p(out.ctypes.data)))
        return out.tobytes()

    # ---- peer-mapped receive buffers (mode "p2p")
    def peer_buffers(self, nbytes, dist):
        """allocate this rank's receive buffer, exchange CUDA IPC handles, map every peer's buffer; cached per size"""
        world, rank = dist.get_world_size(), dist.get_rank()
        key = (nbytes, world)
        if self._peer is not None and self._peer[0] == key:
            return self._peer[1], self._peer[2]
        self.release_peer_buffers()
        own, handle = C.c_void_p(), (C.c_uint8 * 64)()
        self.api.check(self.L.ml_ipc_alloc(C.c_size_t(nbytes), C.byref(own), handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        allh = [torch.empty(64, dtype=torch.uint8, device=self.device) for _ in range(world)]
        dist.all_gather(allh, mine)
        bases, ok = [], 1
        for g in range(world):
            if g == rank:
                bases.append(own.value)
            else:
                hb = (C.c_uint8 * 64)(*allh[g].cpu().tolist())
                p = C.c_void_p()
                if self.L.ml_ipc_open(hb, C.byref(p)) != 0:  # no peer access to that GPU from here
                    ok = 0
                    bases.append(None)
                else:
                    bases.append(p.value)
        # every rank must take the same path: agree on whether all mappings exist
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self._peer = (key, own.value, bases, rank)
        if int(flag[0]) == 0:
            self.release_peer_buffers()
            return None, None
        return own.value, bases

    def release_peer_buffers(self):
        if self._peer is None:
            return
        _, own, bases, rank = self._peer
        torch.cuda.synchronize()
        for g, b in enumerate(bases):
            if g != rank and b is not None:
                self.L.ml_ipc_close(C.c_void_p(b))
        self.L.ml_ipc_free(C.c_void_p(own))
        self._peer = None

    def pack_pairs_peer(self, code_t, n_code, n_ranks, global_poly, bases, max_ctas=0):
        arr = (C.c_void_p * len(bases))(*bases)
        self.api.check(self.L.ml_pack_pairs_peer_dev(C.c_void_p(code_t.data_ptr()), C.c_size_t(n_code), C.c_size_t(n_ranks), C.c_size_t(global_poly),
                                                      arr, C.c_uint(max_ctas), self._stream()))


def owner(j, world):
    return j % world


def _finish(root, backend, dist, world):
    """phase 4: all-gather the subtree roots, hash the top log2(world) levels on every rank"""
    if world == 1:
        return bytes(root.cpu().numpy().tobytes())
    if root.is_cuda:
        g_dev = torch.empty(32 * world, dtype=torch.uint8, device=root.device)
        dist.all_gather_into_tensor(g_dev, root)
        gathered = bytes(g_dev.cpu().numpy().tobytes())
    else:
        parts = [torch.empty(32, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, root)
        gathered = b"".join(bytes(g.numpy().tobytes()) for g in parts)
    return backend.top(gathered, world)


def sharded_batch_commit(local_evals, n, n_polys, backend, dist=None, mode="serial", p2p_ctas=32):
    """local_evals: this rank's polynomials (uint8 tensors of 16*n bytes), in increasing global index.
    Returns the batch root (identical on every rank)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_code, leaves = 2 * n, n
    assert n_polys % world == 0 and leaves % world == 0, "polynomials and leaves must split evenly over ranks"
    n_local, rows = n_polys // world, leaves // world
    assert len(local_evals) == n_local
    chunk = rows * 32
    if mode == "p2p" and (world == 1 or not getattr(backend, "is_cuda", False)):
        mode = "pipelined"
    if mode == "serial":
        send = backend.empty(world * n_local * chunk)
        if getattr(backend, "is_cuda", False):
            enc = backend.encode_streams()
            enc[1].wait_stream(enc[0])
            for pl, ev in enumerate(local_evals):  # phase 1 + packing in exchange order, alternating between two streams
                with torch.cuda.stream(enc[pl & 1]):
                    code = backend.encode(ev, n)
                    backend.pack_pairs(code, n_code, world, n_local, pl, send)
            enc[0].wait_stream(enc[1])
        else:
            for pl, ev in enumerate(local_evals):
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
        return _finish(root, backend, dist, world)

    # ---- pipelined / p2p: the receive buffer is [global poly j][rows][32] (j = pl*world + src), one exchange per polynomial
    offsets = [j * chunk for j in range(n_polys)]
    cuda = getattr(backend, "is_cuda", False)
    side = backend.side_stream() if cuda else None
    main = torch.cuda.current_stream() if cuda else None
    if mode == "p2p":
        own, bases = backend.peer_buffers(n_polys * chunk, dist)
        if own is None:  # some pair of GPUs has no peer access: every rank falls back to the NCCL exchange
            mode = "pipelined"
    if mode == "p2p":
        codes = [backend.empty(32 * n) for _ in range(2)]
        freed = [None, None]  # event: the store pass has finished reading code buffer b
        enc = backend.encode_streams()
        enc[1].wait_stream(enc[0])
        for pl, ev in enumerate(local_evals):
            b = pl & 1
            with torch.cuda.stream(enc[b]):  # code buffer b belongs to encode stream b
                if freed[b] is not None:
                    enc[b].wait_event(freed[b])
                backend.encode(ev, n, out=codes[b])
                done = torch.cuda.Event()
                done.record(enc[b])
            with torch.cuda.stream(side):  # NVLink stores overlap the next polynomial's NTT
                side.wait_event(done)
                backend.pack_pairs_peer(codes[b], n_code, world, pl * world + rank, bases, max_ctas=p2p_ctas)
                freed[b] = torch.cuda.Event()
                freed[b].record(side)
        side.synchronize()  # my stores have landed in every peer's HBM
        dist.barrier()      # ... and everybody else's in mine
        root = backend.leaf_subtree_root_ptr(own, offsets, rows)
        # the all-gather below also fences the next call: nobody overwrites a receive buffer that is still being hashed
        return _finish(root, backend, dist, world)

    assert mode == "pipelined", mode
    recv = backend.empty(n_polys * chunk)
    sends = [backend.empty(world * chunk) for _ in range(2 if cuda else 1)]
    freed = [None, None]
    enc = backend.encode_streams() if cuda else None
    if cuda:
        enc[1].wait_stream(enc[0])
    for pl, ev in enumerate(local_evals):
        b = pl & 1 if cuda else 0
        dst = recv[pl * world * chunk:(pl + 1) * world * chunk]   # [src][rows][32] = polys pl*world .. pl*world+world-1
        if cuda:
            with torch.cuda.stream(enc[b]):  # send buffer b belongs to encode stream b
                if freed[b] is not None:
                    enc[b].wait_event(freed[b])
                code = backend.encode(ev, n)
                backend.pack_pairs(code, n_code, world, 1, 0, sends[b])  # [dest][rows][32]
                if world == 1:
                    dst.copy_(sends[b])
                done = torch.cuda.Event()
                done.record(enc[b])
        else:
            code = backend.encode(ev, n)
            backend.pack_pairs(code, n_code, world, 1, 0, sends[b])
            if world == 1:
                dst.copy_(sends[b])
        if world == 1:
            pass
        elif cuda:
            with torch.cuda.stream(side):
                side.wait_event(done)
                dist.all_to_all_single(dst, sends[b])
                freed[b] = torch.cuda.Event()
                freed[b].record(side)
        else:
            dist.all_to_all_single(dst, sends[b])
    if cuda:
        main.wait_stream(enc[1])
        if world > 1:
            main.wait_stream(side)
    root = backend.leaf_subtree_root(recv, offsets, rows)
    return _finish(root, backend, dist, world)
