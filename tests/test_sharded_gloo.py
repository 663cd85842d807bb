"""N>1 path on CPU: world_size-2 gloo run of the sharded batched commit (exchange + root gather logic), with the CPU
oracle standing in for the CUDA kernels.  The result must equal the single-process `Merkle::batch_commit` root."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """same interface as multilinear_b200.sharded.CudaBackend, computed by the oracle on CPU tensors"""

    def __init__(self):
        from oracle.binding import Oracle
        self.O = Oracle(threads=1)

    def empty(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def encode(self, evals_t, n):
        ev = evals_t.numpy().reshape(n, 16)
        coeffs = self.O.bit_reverse(self.O.to_coefficient(ev))
        code = self.O.reed_solomon(coeffs, self.O.pow2_generator(int(np.log2(n)) + 1))
        return torch.from_numpy(np.ascontiguousarray(code).reshape(-1).copy())

    def pack_pairs(self, code_t, n_code, n_ranks, n_local, pl, send_t):
        code = code_t.numpy().reshape(n_code, 16)
        half = n_code // 2
        pairs = np.concatenate([code[:half], code[half:]], axis=1)  # (half, 32)
        rows = half // n_ranks
        view = send_t.numpy().reshape(n_ranks, n_local, rows, 32)
        view[:, pl] = pairs.reshape(n_ranks, rows, 32)

    def leaf_subtree_root(self, recv_t, ptr_offsets, rows):
        buf = recv_t.numpy()
        datas = [buf[o:o + rows * 32].reshape(rows, 32) for o in ptr_offsets]
        return torch.from_numpy(np.frombuffer(self.O.merkle_batch_commit(datas).root(), dtype=np.uint8).copy())

    def top(self, roots_bytes, n_roots):
        import hashlib
        layer = [roots_bytes[32 * i:32 * i + 32] for i in range(n_roots)]
        while len(layer) > 1:
            layer = [hashlib.sha256(layer[i] + layer[i + 1]).digest() for i in range(0, len(layer), 2)]
        return layer[0]


def _polys(n, n_polys):
    from oracle.binding import Oracle
    O = Oracle(threads=1)
    return [O.synthetic(1000 + j, n) for j in range(n_polys)]


def _worker(rank, world, port, n, n_polys, out_q, mode="serial"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multilinear_b200.sharded import sharded_batch_commit
    polys = _polys(n, n_polys)
    local = [torch.from_numpy(polys[j].reshape(-1).copy()) for j in range(n_polys) if j % world == rank]
    root = sharded_batch_commit(local, n, n_polys, OracleBackend(), dist, mode=mode)
    out_q.put((rank, root))
    dist.barrier()
    dist.destroy_process_group()


def _reference_root(n, n_polys):
    from oracle.binding import Oracle
    O = Oracle(threads=1)
    be = OracleBackend()
    datas = []
    for p in _polys(n, n_polys):
        code = be.encode(torch.from_numpy(p.reshape(-1).copy()), n).numpy().reshape(2 * n, 16)
        datas.append(np.concatenate([code[:n], code[n:]], axis=1))
    return O.merkle_batch_commit(datas).root()


@pytest.mark.parametrize("mode", ["serial", "pipelined"])
def test_sharded_batch_commit_world2_matches_single_process(mode):
    n, n_polys, world = 256, 6, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, n_polys, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _reference_root(n, n_polys)
    assert results[0] == results[1] == want


@pytest.mark.parametrize("mode", ["serial", "pipelined"])
def test_sharded_batch_commit_single_rank(mode):
    sys.path.insert(0, ROOT)
    from multilinear_b200.sharded import sharded_batch_commit
    n, n_polys = 128, 3
    local = [torch.from_numpy(p.reshape(-1).copy()) for p in _polys(n, n_polys)]
    assert sharded_batch_commit(local, n, n_polys, OracleBackend(), None, mode=mode) == _reference_root(n, n_polys)
