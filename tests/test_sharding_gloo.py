"""CPU tier, world_size 2 over gloo: the multi-GPU host logic (contiguous slices, one all-gather of the
384-byte Miller partial products, one final exponentiation) reproduces the unsharded result.  The oracle
stands in for the GPU engine here - this is a test of the plumbing, not of the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """duck-typed stand-in for sylow_b200.Engine backed by oracle/sylow_oracle.c (test infrastructure)."""

    def __init__(self):
        from oracle import c_oracle

        self.c = c_oracle

    def verify_batch_partial(self, pks, msgs, sigs, dst=None, weight_seed=None, first_index=0):
        # prod miller(r_i sig, G2gen) * miller(-r_i H(m), pk) over the slice, via the oracle's primitives
        c = self.c
        n = msgs[1].size - 1
        if n == 0:
            one = np.zeros(384, np.uint8)
            one[0] = 1
            return one
        hm, _ = c.hash_to_g1_batch(msgs, threads=1)
        if weight_seed is not None:
            from oracle import bn254_py as ob

            r = np.zeros((n, 32), np.uint8)
            for i in range(n):
                wt = int.from_bytes(ob.keccak256(bytes(weight_seed) + (first_index + i).to_bytes(8, "little"))[:8], "little") | 1
                r[i, :8] = np.frombuffer(wt.to_bytes(8, "little"), np.uint8)
            hm, _ = c.g1_mul_batch(hm, r, threads=1)
            sigs, _ = c.g1_mul_batch(np.ascontiguousarray(sigs), r, threads=1)
        P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
        neg = hm.copy()
        for i in range(n):
            y = int.from_bytes(bytes(hm[i, 32:]), "little")
            neg[i, 32:] = np.frombuffer(((P - y) % P).to_bytes(32, "little"), dtype=np.uint8)
        from tests import wire as w
        from oracle import bn254_py as o

        gen = np.tile(np.frombuffer(w.g2_b(o.G2_GEN), dtype=np.uint8), (n, 1))
        g1 = np.concatenate([np.ascontiguousarray(sigs), neg])
        g2 = np.concatenate([gen, np.ascontiguousarray(pks)])
        return c.miller_product(g1, g2, threads=1)

    def miller_product(self, g1, g2):
        if g1.shape[0] == 0:
            one = np.zeros(384, np.uint8)
            one[0] = 1
            return one
        return self.c.miller_product(g1, g2, threads=1)

    def fp12_product(self, f):
        from oracle import bn254_py as o
        from tests import wire as w

        acc = o.FP12_ONE
        for row in f:
            acc = o.fp12_mul(acc, w.b_fp12(bytes(row)))
        return np.frombuffer(w.fp12_b(acc), dtype=np.uint8).copy()

    def verify_batch_finish(self, partials):
        prod = self.fp12_product(partials)
        gt = self.c.final_exp_batch(prod.reshape(1, 384), threads=1)
        one = np.zeros(384, np.uint8)
        one[0] = 1
        return bool((gt[0] == one).all())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, corrupt, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import c_oracle as c
    from sylow_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(5)
        sks = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
        sks[:, 31] &= 0x0F
        msgs = [bytes(rs.randint(0, 256, size=rs.randint(0, 70), dtype=np.uint8)) for _ in range(n)]
        offs = np.zeros(n + 1, np.uint64)
        offs[1:] = np.cumsum([len(m) for m in msgs])
        buf = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
        sigs = c.sign_batch(sks, (buf, offs), threads=2)
        from oracle import bn254_py as o
        from tests import wire as w

        gen = np.tile(np.frombuffer(w.g2_b(o.G2_GEN), dtype=np.uint8), (n, 1))
        pks, _ = c.g2_mul_batch(gen, sks, threads=2)
        if corrupt:
            sigs[n - 1] = sigs[0]
        eng = OracleEngine()
        ok = sharding.verify_batch_sharded(eng, pks, (buf, offs), sigs)
        seed = bytes(range(32))
        okw = sharding.verify_batch_sharded(eng, pks, (buf, offs), sigs, weight_seed=seed)
        forged_u = forged_w = None
        if n >= 2 and not corrupt:
            # a forgery that cancels ACROSS the two ranks: sig_0 + D on rank 0's slice, sig_(n-1) - D on rank 1's.  The
            # unweighted product (the reference example's check) accepts it; the weighted form rejects it on every rank.
            F = o.FpOps
            D = o.affine_to_proj(F, w.rand_g1(__import__("random").Random(3)))
            f = sigs.copy()
            for i, d in ((0, D), (n - 1, o.proj_neg(F, D))):
                pt = o.proj_to_affine(F, o.proj_add(F, o.affine_to_proj(F, w.b_g1(bytes(sigs[i]))), d))
                f[i] = np.frombuffer(w.g1_b(pt), np.uint8)
            forged_u = sharding.verify_batch_sharded(eng, pks, (buf, offs), f)
            forged_w = sharding.verify_batch_sharded(eng, pks, (buf, offs), f, weight_seed=seed)
        g1 = np.concatenate([sigs, sigs])[:n]
        prod = sharding.miller_product_sharded(eng, g1, pks)
        q.put((rank, ok, bytes(prod), bytes(c.miller_product(g1, pks, threads=1)), okw, forged_u, forged_w))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,corrupt", [(5, False), (5, True), (1, False)])
def test_world_size_2_gloo(n, corrupt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, corrupt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, prod, ref, okw, forged_u, forged_w in res:
        assert ok is (not corrupt), (rank, ok)
        assert okw is (not corrupt), (rank, okw)  # random-weight form: same verdict on valid / corrupted batches
        assert prod == ref  # sharded glued product == unsharded, bit for bit
        if n >= 2 and not corrupt:
            assert forged_u is True and forged_w is False, (rank, forged_u, forged_w)


def test_rank_slice_covers_batch():
    from sylow_b200.sharding import rank_slice, slice_messages

    for n in (0, 1, 7, 8, 1 << 20):
        for world in (1, 2, 4, 8):
            idx = np.arange(n)
            parts = [idx[rank_slice(n, r, world)] for r in range(world)]
            assert np.array_equal(np.concatenate(parts), idx)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    buf = np.arange(10, dtype=np.uint8)
    offs = np.array([0, 3, 3, 7, 10], dtype=np.uint64)
    b, o2 = slice_messages(buf, offs, slice(1, 3))
    assert b.tolist() == [3, 4, 5, 6] and o2.tolist() == [0, 0, 4]
    with pytest.raises(ValueError):
        rank_slice(4, 2, 2)
