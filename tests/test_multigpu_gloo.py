"""world_size-2 gloo test of the multi-GPU orchestration (graph_b200/multigpu.py) on CPUs.

The CUDA shard backend needs a GPU, so this test injects an oracle-based backend (numpy, test-only)
with the same interface; what is covered here is the host logic the N>1 path adds: the in-degree
partition, the per-sweep exchange of out_scores slices, the all-reduced error / stop rule and the
assembly of the final score vector.  The GPU backend itself is covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleShardBackend:
    """Jacobi sweep over a destination range on the CPU (identity renumbering)."""

    def __init__(self, in_off, in_tgt, out_off, rank, world):
        import oracle
        self.in_off, self.in_tgt = in_off.astype(np.int64), in_tgt
        self.outdeg = np.diff(out_off.astype(np.int64)).astype(np.float32)
        self.n = len(in_off) - 1
        self.n_active = self.n
        r = oracle.in_degree_partition(in_off, world).tolist()
        self.ranges = r + [self.n] * (world + 1 - len(r))
        self.rb, self.re = self.ranges[rank], self.ranges[rank + 1]
        self.device = torch.device("cpu")
        self.launches = 0

    def init(self, damping, x0, x1, scores):
        init = np.float32(1.0) / np.float32(self.n)
        with np.errstate(divide="ignore"):
            x0.numpy()[:] = init / self.outdeg
        scores.numpy()[:] = init

    def step(self, damping, sweep_no, x_cur, x_next, peers, scores, err):
        xc, xn, sc = x_cur.numpy(), x_next.numpy(), scores.numpy()
        base = (np.float32(1.0) - np.float32(damping)) / np.float32(self.n)
        e = 0.0
        d = np.float32(damping)
        for u in range(self.rb, self.re):
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
            q.put((spr.ranges, results))
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
    ranges, results = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    src, dst = oracle.rmat_edges(scale, seed=42)
    n = 1 << scale
    out_off, _ = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    in_off, in_tgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    assert ranges[0] == 0 and ranges[-1] == n and 0 < ranges[1] < n
    # both shards carry about half of the edges (greedy_node_map_partition)
    m = int(in_off[-1])
    assert abs(int(in_off[ranges[1]]) - m / 2) < 0.1 * m
    for (maxit, tol), (it, err, scores) in zip(cfgs, results):
        want, wit, werr = oracle.page_rank_jacobi(in_off, in_tgt, out_off, maxit, tol, 0.85, acc64=True)
        assert it == wit
        assert np.max(np.abs(scores - want) / want) <= 1e-6
        assert abs(err - werr) <= 2e-6


_free_port_shared = _free_port()


def test_rebalance_cuts_pure():
    from graph_b200.multigpu import rebalance_cuts
    # balanced times: cuts stay put
    c = rebalance_cuts([0.25, 0.5, 0.75], [1.0, 1.0, 1.0, 1.0])
    assert np.allclose(c, [0.25, 0.5, 0.75])
    # the last rank is the straggler: every cut moves right (its interval shrinks), damped by 1/2
    c = rebalance_cuts([0.25, 0.5, 0.75], [1.0, 1.0, 1.0, 3.0])
    assert all(b > a for a, b in zip([0.25, 0.5, 0.75], c)) and c == sorted(c) and c[-1] < 1.0
    # exact for a piecewise-constant density when applied without damping twice the step
    full = [2 * n - o for n, o in zip(c, [0.25, 0.5, 0.75])]
    dens = [1 / 0.25, 1 / 0.25, 1 / 0.25, 3 / 0.25]
    bounds = [0.0] + full + [1.0]
    seg_time = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        tt, edges = 0.0, [0.0, 0.25, 0.5, 0.75, 1.0]
        for d, (a, b) in zip(dens, zip(edges[:-1], edges[1:])):
            tt += d * max(0.0, min(hi, b) - max(lo, a))
        seg_time.append(tt)
    assert np.allclose(seg_time, [1.5] * 4)
    # degenerate inputs stay strictly inside (0, 1) and increasing
    c = rebalance_cuts([0.1, 0.2, 0.3], [0.0, 0.0, 0.0, 5.0])
    assert 0 < c[0] < c[1] < c[2] < 1
