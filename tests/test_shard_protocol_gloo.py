"""N>1 host logic of the sharded batched prover (csrc/shard.cu) on CPU, world_size 2 over gloo:
  * the IPC-record all-gather of ShardedBatchedProver.connect_over hands every rank all records in rank order;
  * the fingerprint decomposition of phases S2/S3 — Horner partial sums of the rank's polynomials in rho^G scaled by
    rho^(G-1-rank), exchanged slice-wise and added by the slice owner — equals the reference's fingerprint over the whole batch
    (src/fri/batched_fri.rs:30-38, batched_pcs.rs:55-63), checked against the oracle."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    from multilinear_b200.api import ShardedBatchedProver
    from oracle import pyref as P
    from oracle.binding import Oracle, fe_ints
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        # ---- record exchange (no GPU: a stand-in handle that exports a recognisable record and captures what connect gets)
        class Fake(ShardedBatchedProver):
            def __init__(self):
                self.local_ranks, self.got, self.h = [rank], None, None

            def export(self):
                return bytes([rank]) + bytes(3) + bytes(4) + bytes([0xA0 + rank]) * 64

            def connect(self, records):
                self.got = bytes(records)

        f = Fake()
        f.connect_over(dist, torch.device("cpu"))
        ok_records = len(f.got) == 72 * world and all(f.got[72 * g] == g and f.got[72 * g + 8] == 0xA0 + g for g in range(world))

        # ---- fingerprint decomposition
        O = Oracle(threads=1)
        M = P.M
        B, n = 6, 64
        polys = [fe_ints(O.synthetic(900 + j, n)) for j in range(B)]
        rho = fe_ints(O.synthetic(77, 1))[0]
        slice_len = n // world
        rho_g, scale = pow(rho, world, M), pow(rho, world - 1 - rank, M)
        mine = [polys[j] for j in range(rank, B, world)]
        partial = []
        for i in range(n):
            acc = 0
            for p in mine:
                acc = (acc * rho_g + p[i]) % M
            partial.append(acc * scale % M)
        # slice s of my partial goes to rank s (the stage area of the slice owner)
        hi = torch.tensor([int(x >> 96) for x in partial], dtype=torch.int64)
        m1 = torch.tensor([int((x >> 64) & 0xFFFFFFFF) for x in partial], dtype=torch.int64)
        m0 = torch.tensor([int((x >> 32) & 0xFFFFFFFF) for x in partial], dtype=torch.int64)
        lo = torch.tensor([int(x & 0xFFFFFFFF) for x in partial], dtype=torch.int64)
        packed = torch.stack([hi, m1, m0, lo], dim=1).contiguous()  # (n, 4) 32-bit limbs
        recv = torch.empty_like(packed)
        dist.all_to_all_single(recv, packed)  # recv[g*slice_len + r] = rank g's partial at position rank*slice_len + r
        total = []
        for r in range(slice_len):
            acc = 0
            for g in range(world):
                a, b, c, d = [int(v) for v in recv[g * slice_len + r]]
                acc = (acc + ((a << 96) | (b << 64) | (c << 32) | d)) % M
            total.append(acc)
        want = [O.fingerprint(rho, np.array([np.frombuffer(int(polys[j][rank * slice_len + r]).to_bytes(16, "little"), dtype=np.uint8) for j in range(B)]))
                for r in range(slice_len)]
        out_q.put((rank, ok_records, total == want))
    finally:
        dist.destroy_process_group()


def test_shard_record_exchange_and_fingerprint_decomposition_world2():
    from oracle import binding
    binding.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]
