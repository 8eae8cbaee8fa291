"""world_size-2 gloo test of the multi-GPU orchestration (graph_b200/multigpu.py) on CPUs.

The CUDA shard backend needs a GPU, so this test injects an oracle-based backend (numpy, test-only)
with the same interface; what is covered here is the host logic the N>1 path adds: the cyclic deal
of 32-row slices, the per-sweep exchange of the dealt out_scores slices (packed all-gather), the
all-reduced error / stop rule and the assembly of the final score vector.  The GPU backend itself is covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleShardBackend:
    """Jacobi sweep over this rank's dealt 32-row slices on the CPU (identity renumbering)."""

    def __init__(self, in_off, in_tgt, out_off, rank, world):
        from graph_b200.multigpu import owner_of_rows
        self.in_off, self.in_tgt = in_off.astype(np.int64), in_tgt
        self.outdeg = np.diff(out_off.astype(np.int64)).astype(np.float32)
        self.n = len(in_off) - 1
        self.n_active = self.n
        self.rank, self.world = rank, world
        self.rows = np.nonzero(owner_of_rows(np.arange(self.n), world) == rank)[0]
        self.device = torch.device("cpu")
        self.launches = 0

    def init(self, damping, x0, x1, scores):
        init = np.float32(1.0) / np.float32(self.n)
        with np.errstate(divide="ignore"):
            x0.numpy()[:] = init / self.outdeg
        sc = scores.numpy()
        sc[:] = 0.0               # rows of other ranks stay 0: the score vectors are summed at the end
        sc[self.rows] = init

    def step(self, damping, sweep_no, x_cur, x_next, peers, scores, err):
        xc, xn, sc = x_cur.numpy(), x_next.numpy(), scores.numpy()
        base = (np.float32(1.0) - np.float32(damping)) / np.float32(self.n)
        e = 0.0
        d = np.float32(damping)
        for u in self.rows:
            tot = np.float32(xc[self.in_tgt[self.in_off[u]:self.in_off[u + 1]]].astype(np.float64).sum())
            new = np.float32(base + np.float32(d * tot))
            e += abs(float(np.float32(new - sc[u])))
            sc[u] = new
            with np.errstate(divide="ignore"):
                xn[u] = new / self.outdeg[u]
        err.numpy()[0] = e

    def finish(self, scores_internal):
        return scores_internal.clone()


def _worker(rank, world, port, scale, cfgs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from graph_b200.multigpu import ShardedPageRank
        src, dst = oracle.rmat_edges(scale, seed=42)
        n = 1 << scale
        out_off, _ = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
        in_off, in_tgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
        spr = ShardedPageRank(backend=OracleShardBackend(in_off, in_tgt, out_off, rank, world), exchange="nccl")
        results = []
        for maxit, tol in cfgs:
            spr.run(maxit, 0.85, tol)
            results.append((spr.ran_iterations, spr.error, spr.scores_host()))
        if rank == 0:
            q.put((len(spr.backend.rows), results))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_sharded_page_rank_two_ranks_gloo():
    import oracle
    scale, world = 10, 2
    cfgs = [(20, 0.0), (50, 1e-4), (3, 1.0)]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_shared, scale, cfgs, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows0, results = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    src, dst = oracle.rmat_edges(scale, seed=42)
    n = 1 << scale
    out_off, _ = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    in_off, in_tgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    assert rows0 == n // 2   # slices dealt round-robin: rank 0 owns every other slice
    for (maxit, tol), (it, err, scores) in zip(cfgs, results):
        want, wit, werr = oracle.page_rank_jacobi(in_off, in_tgt, out_off, maxit, tol, 0.85, acc64=True)
        assert it == wit
        assert np.max(np.abs(scores - want) / want) <= 1e-6
        assert abs(err - werr) <= 2e-6


_free_port_shared = _free_port()


def test_owner_of_rows_is_a_cyclic_deal_of_slices():
    from graph_b200.multigpu import owner_of_rows
    o = owner_of_rows(np.arange(32 * 7 + 5), 3)
    assert (o[:32] == 0).all() and (o[32:64] == 1).all() and (o[64:96] == 2).all() and (o[96:128] == 0).all()
    assert o[-1] == (7 % 3)


# ---- multi-GPU WCC orchestration (phases + all-gather + merge) with a numpy backend ----------------
class NumpyWccBackend:
    """The phases of gb_wcc_shard_phase on the CPU (Afforest link rule, afforest.rs:22-39)."""

    def __init__(self, out_off, out_tgt, in_off, in_tgt, rounds=2, samples=64):
        self.oo, self.ot, self.io, self.it = out_off.astype(np.int64), out_tgt, in_off.astype(np.int64), in_tgt
        self.n = len(out_off) - 1
        self.rounds, self.samples = rounds, samples
        self.device = torch.device("cpu")

    def new_parent(self):
        return torch.empty(self.n, dtype=torch.int32)

    @staticmethod
    def _link(p, u, v):
        p1, p2 = p[u], p[v]
        while p1 != p2:
            hi, lo = max(p1, p2), min(p1, p2)
            ph = p[hi]
            if ph == lo:
                break
            if ph == hi:
                p[hi] = lo
                break
            p1, p2 = p[p[hi]], p[lo]

    def phase(self, which, parent, vb=0, ve=0, skip=0, use_skip=0, other=None):
        from graph_b200 import _capi
        p = parent.numpy()
        if which == _capi.WCC_INIT:
            p[:] = np.arange(self.n)
        elif which == _capi.WCC_SAMPLE:
            for u in range(vb, ve):
                for t in self.ot[self.oo[u]:self.oo[u] + min(self.rounds, self.oo[u + 1] - self.oo[u])]:
                    self._link(p, u, int(t))
        elif which == _capi.WCC_COMPRESS:
            for x in range(self.n):
                while p[x] != p[p[x]]:
                    p[x] = p[p[x]]
        elif which == _capi.WCC_MERGE:
            o = other.numpy()
            for v in range(self.n):
                if o[v] != v:
                    self._link(p, v, int(o[v]))
        elif which == _capi.WCC_LINK_REMAINING:
            for u in range(vb, ve):
                if use_skip and p[u] == skip:
                    continue
                for t in self.ot[self.oo[u] + min(self.rounds, self.oo[u + 1] - self.oo[u]):self.oo[u + 1]]:
                    self._link(p, u, int(t))
                for t in self.it[self.io[u]:self.io[u + 1]]:
                    self._link(p, u, int(t))

    def sample_label(self, parent):
        p = parent.numpy()
        idx = (np.arange(self.samples, dtype=np.int64) * 2654435761) % self.n
        vals, counts = np.unique(p[idx], return_counts=True)
        return int(vals[np.argmax(counts)]), True


def _wcc_worker(rank, world, port, scale, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from graph_b200.multigpu import ShardedWcc
        src, dst = oracle.rmat_edges(scale, seed=9)
        n = 1 << scale
        oo, ot = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
        io, it = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
        comp = ShardedWcc(backend=NumpyWccBackend(oo, ot, io, it)).run().numpy().astype(np.uint32)
        q.put((rank, comp))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_wcc_two_ranks_gloo():
    import oracle
    scale, world = 9, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_wcc_worker, args=(r, world, port, scale, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    src, dst = oracle.rmat_edges(scale, seed=9)
    oo, ot = oracle.csr_build(src, dst, 1 << scale, oracle.OUTGOING, oracle.SORTED)
    want = oracle.wcc_min_label(oo, ot)
    assert (got[0] == want).all() and (got[1] == want).all()   # every rank: min node id per component


def test_vertex_ranges_cover_and_align():
    from graph_b200.multigpu import vertex_ranges
    for n, w in ((1000, 3), (64, 8), (5, 2), (1 << 20, 8)):
        r = vertex_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 32 == 0 or lo == n for lo, _ in r)
