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
