"""Multi-GPU PageRank: 1-D edge-cut across the GPUs of one box, one process per GPU.

The reference has no distributed code at all; this is the one exchange step
BASELINE.json's north_star adds: every rank sweeps the destination rows of its shard
(`gb_pr_shard_step`) and the per-sweep out_scores slices are exchanged either

  * "peer"  — fused: the sweep kernel stores each finished out_score straight into every peer's
              next vector through NVLink peer mappings (torch symmetric memory supplies the
              pointers); the 8-byte all-reduce of the sweep error is the only collective and doubles
              as the inter-sweep barrier, or
  * "nccl"  — baseline: one NCCL broadcast per shard slice after the kernel.

Shards are ranges of INTERNAL rows (the JACOBI path's renumbering), chosen by the reference's own
greedy in-degree rule (crates/builder/src/graph_ops.rs:431-439, :479-509) so every rank derives the
same ranges without talking to anyone.  `torch.distributed` is plumbing only; the compute is the
same CUDA kernel as on one GPU.  The compute backend is injectable so that the orchestration
(partition, exchange, stop rule, assembly) is testable with gloo on CPUs (tests/test_multigpu_gloo.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import check, lib


def rebalance_cuts(cuts, times):
    """New cut fractions from the current ones and the measured per-rank times: the time density is
    taken as constant inside each rank's weight interval and the cuts move (half-way, damped) to
    where the cumulative time crosses k/P.  Pure function: every rank evaluates it on the same
    all-gathered numbers."""
    P = len(times)
    t = [max(float(v), 1e-6) for v in times]
    bounds = [0.0] + list(cuts) + [1.0]
    total = sum(t)
    new_cuts, acc, p = [], 0.0, 0
    for k in range(1, P):
        target = total * k / P
        while p < P - 1 and acc + t[p] < target:
            acc += t[p]
            p += 1
        frac = (target - acc) / t[p]
        new_cuts.append(bounds[p] + min(max(frac, 0.0), 1.0) * (bounds[p + 1] - bounds[p]))
    out = [0.5 * a + 0.5 * b for a, b in zip(cuts, new_cuts)]
    eps = 1e-6
    for k in range(P - 1):
        lo = (out[k - 1] + eps) if k else eps
        out[k] = min(max(out[k], lo), 1.0 - (P - 1 - k) * eps)
    return out


class CudaShardBackend:
    """The product backend: gb_pr_shard_* of libgraph_b200.so on this rank's GPU."""

    def __init__(self, graph, rank: int, world: int, row_cost: int = 3):
        self.graph = graph
        self.rank, self.world, self.row_cost = rank, world, row_cost
        self.n = graph.node_count()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._shard = C.c_void_p()
        self.launches = 0
        self.repartition(None)

    def repartition(self, cuts):
        """(Re)builds this rank's shard; cuts = None (greedy rule) or world-1 weight fractions."""
        ranges = np.zeros(self.world + 1, np.uint32)
        carr = None
        if cuts is not None:
            carr = (C.c_double * (self.world - 1))(*[float(c) for c in cuts])
        check(lib.gb_pr_shard_partition(self.graph._g, self.world, self.row_cost, carr,
                                        ranges.ctypes.data_as(C.c_void_p)))
        new_ranges = [int(v) for v in ranges]
        fresh = C.c_void_p()
        check(lib.gb_pr_shard_create(self.graph._g, new_ranges[self.rank], new_ranges[self.rank + 1],
                                     C.byref(fresh)))
        # the old shard is released only once the new one exists (a failed rebuild keeps the old state)
        old, self._shard, self.ranges = self._shard, fresh, new_ranges
        if old:
            check(lib.gb_pr_shard_free(old))
        rb, re, act = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        check(lib.gb_pr_shard_info(self._shard, C.byref(rb), C.byref(re), C.byref(act), None))
        self.n_active = int(act.value)

    def __del__(self):
        sh, self._shard = getattr(self, "_shard", None), None
        if sh:
            try:
                lib.gb_pr_shard_free(sh)
            except Exception:
                pass

    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def init(self, damping, x0, x1, scores):
        check(lib.gb_pr_shard_init(self._shard, damping, C.c_void_p(x0.data_ptr()), C.c_void_p(x1.data_ptr()),
                                   C.c_void_p(scores.data_ptr()), self._stream()))
        self.launches += 1

    def step(self, damping, sweep_no, x_cur, x_next, peer_ptrs, scores, err):
        arr = None
        if peer_ptrs:
            arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        check(lib.gb_pr_shard_step(self._shard, damping, sweep_no, C.c_void_p(x_cur.data_ptr()),
                                   C.c_void_p(x_next.data_ptr()), arr, len(peer_ptrs or ()),
                                   C.c_void_p(scores.data_ptr()), C.c_void_p(err.data_ptr()), self._stream()))
        self.launches += 3 if sweep_no == 1 else 2

    def finish(self, scores_internal):
        out = torch.empty_like(scores_internal)
        check(lib.gb_pr_shard_finish(self._shard, C.c_void_p(scores_internal.data_ptr()), C.c_void_p(out.data_ptr()),
                                     self._stream()))
        self.launches += 1
        return out


class ShardedPageRank:
    """page_rank over `world` shards; every rank ends up with the full score vector.

    run(max_iterations, damping, tolerance) follows page_rank.rs:88-110: sweep, error = sum over all
    ranks, stop when error < tolerance or the sweep count reaches max_iterations.
    """

    def __init__(self, graph=None, exchange: str = "auto", backend=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # per-row charge of the partition: ~2 edge-equivalents of vector traffic, plus (fused exchange)
        # one remote store per peer — measured 18 with 7 peers (profiles/r01_multigpu_diag.txt)
        want_peer = exchange in ("auto", "peer") and torch.cuda.is_available() and self.world > 1
        row_cost = 4 + 2 * (self.world - 1) if want_peer else 3
        self.backend = backend if backend is not None else CudaShardBackend(graph, self.rank, self.world, row_cost)
        b = self.backend
        self.n, self.n_active, self.ranges = b.n, b.n_active, b.ranges
        dev = b.device
        self.exchange = exchange if exchange in ("nccl", "allgather") else "nccl"
        self._peer_next = [None, None]
        self.diag = None
        self.x = None
        if exchange in ("auto", "peer") and dev.type == "cuda" and self.world > 1:
            try:
                self._setup_symmetric(dev)
                self.exchange = "peer"
            except Exception as exc:  # symmetric memory unavailable: fall back to the NCCL exchange
                if exchange == "peer":
                    raise
                self._symm_error = repr(exc)
        if self.x is None:
            self.x = [torch.empty(self.n, dtype=torch.float32, device=dev) for _ in range(2)]
        self.scores = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.err = torch.zeros(1, dtype=torch.float64, device=dev)
        self.ran_iterations = 0
        self.error = 0.0

    # -- symmetric memory (NVLink peer mappings) --
    def _setup_symmetric(self, dev):
        import torch.distributed._symmetric_memory as symm_mem
        group_name = (self.group or dist.group.WORLD).group_name
        buf = symm_mem.empty(2 * self.n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, group_name)
        self._symm = (buf, hdl)
        self.x = [buf[: self.n], buf[self.n:]]
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        for which in (0, 1):
            self._peer_next[which] = [ptrs[p] + which * self.n * 4 for p in range(self.world) if p != self.rank]

    @property
    def launches(self) -> int:
        return getattr(self.backend, "launches", 0)

    def _exchange(self, x_next):
        if self.exchange == "peer":
            return  # the kernel already stored the slice into every peer
        if self.exchange == "allgather":
            # one grouped collective: every rank's (uneven) slice lands in place in every x_next
            views = [x_next[min(self.ranges[p], self.n_active):min(self.ranges[p + 1], self.n_active)]
                     for p in range(self.world)]
            dist.all_gather(views, views[self.rank], group=self.group)
            return
        for p in range(self.world):
            lo, hi = min(self.ranges[p], self.n_active), min(self.ranges[p + 1], self.n_active)
            if hi > lo:
                dist.broadcast(x_next[lo:hi], src=dist.get_global_rank(self.group, p) if self.group else p,
                               group=self.group)

    def run(self, max_iterations: int = 20, damping: float = 0.85, tolerance: float = 0.0):
        b = self.backend
        b.init(damping, self.x[0], self.x[1], self.scores)
        if self.exchange == "peer":
            dist.barrier(group=self.group)  # nobody may store into a peer that is still initialising
        sweep = 0
        limit = max_iterations if max_iterations else 100000
        diag = self.diag
        while True:
            sweep += 1
            cur, nxt = self.x[(sweep - 1) & 1], self.x[sweep & 1]
            peers = self._peer_next[sweep & 1] if self.exchange == "peer" else None
            if diag is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            b.step(damping, sweep, cur, nxt, peers, self.scores, self.err)
            if diag is not None:
                ev[1].record()
            self._exchange(nxt)
            if diag is not None:
                ev[2].record()
                diag.append(ev)
            # total error of the sweep; also the barrier that orders the peer stores of sweep k before
            # any rank's reads in sweep k+1
            dist.all_reduce(self.err, op=dist.ReduceOp.SUM, group=self.group)
            if tolerance > 0.0:
                self.error = float(self.err.item())
                if self.error < tolerance:
                    break
            if sweep == limit:
                break
        self.ran_iterations = sweep
        if not tolerance > 0.0:
            self.error = float(self.err.item())
        return self

    def calibrate(self, rounds: int = 2, sweeps: int = 5, damping: float = 0.85):
        """Measured-time rebalancing (setup, untimed): run a few sweeps, all-gather every rank's mean
        kernel time, treat the time density as constant inside each rank's current weight interval and
        move the cut points to where the cumulative time crosses k/P.  Every rank computes the same
        cuts from the same gathered numbers, so no rank can disagree about the new ranges."""
        if self.world == 1 or not hasattr(self.backend, "repartition"):
            return None
        P = self.world
        cuts = [k / P for k in range(1, P)]
        history = []
        try:
            return self._calibrate(rounds, sweeps, damping, cuts, history)
        except Exception as exc:  # identical on every rank (same inputs): fall back to the static rule
            self.diag = None
            self.backend.repartition(None)
            self.ranges, self.n_active = self.backend.ranges, self.backend.n_active
            self.calibration = {"error": repr(exc)}
            return self.calibration

    def _calibrate(self, rounds, sweeps, damping, cuts, history):
        P = self.world
        for _ in range(rounds):
            self.diag = []
            self.run(sweeps, damping, 0.0)
            torch.cuda.synchronize() if self.scores.is_cuda else None
            k_ms = [a.elapsed_time(b) for a, b, _ in self.diag][1:]  # the first sweep also patches x0
            self.diag = None
            mine = torch.tensor([float(np.mean(k_ms))], dtype=torch.float64, device=self.scores.device)
            allt = [torch.zeros_like(mine) for _ in range(P)]
            dist.all_gather(allt, mine, group=self.group)
            t = [float(v.item()) for v in allt]
            history.append([round(v, 4) for v in t])
            cuts = rebalance_cuts(cuts, t)
            self.backend.repartition(cuts)
            self.ranges, self.n_active = self.backend.ranges, self.backend.n_active
        self.calibration = {"cuts": [round(c, 5) for c in cuts], "kernel_ms_per_rank": history}
        return self.calibration

    def diag_summary(self):
        """(mean kernel ms, mean exchange ms) per sweep of the recorded run (diagnostics only)."""
        torch.cuda.synchronize()
        k = [a.elapsed_time(b) for a, b, _ in self.diag]
        x = [b.elapsed_time(c) for _, b, c in self.diag]
        return float(np.mean(k)), float(np.mean(x))

    def scores_device(self):
        """Full score vector in original ids on this rank's device (exchanges the score slices)."""
        for p in range(self.world):
            lo, hi = min(self.ranges[p], self.n_active), min(self.ranges[p + 1], self.n_active)
            if hi > lo:
                dist.broadcast(self.scores[lo:hi], src=dist.get_global_rank(self.group, p) if self.group else p,
                               group=self.group)
        return self.backend.finish(self.scores)

    def scores_host(self) -> np.ndarray:
        return self.scores_device().cpu().numpy()
