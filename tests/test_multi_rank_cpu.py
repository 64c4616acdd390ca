"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: the table is broadcast once,
targets are dealt round-robin, each rank runs HITON-PC for its shard, rank 0 merges the neighbour lists into the
same graph a single process produces.  The per-rank engine here is a stand-in built on the oracle (there is no GPU
in this container); on the GPU box the same plumbing wraps fw.Engine (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fwload
from oracle import fwo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _OracleResult:
    def __init__(self, targets, lists, nt):
        self.targets, self._l, self.num_tests = np.asarray(targets, np.int64), lists, np.asarray(nt, np.int64)

    def pc(self, i):
        return self._l[i]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fw = fwload.load()
    par = fwload.load_sub("parallel")
    synth = fwload.load_sub("synth")
    p, n = 72, 300
    table = torch.zeros((p, n), dtype=torch.float32)
    if rank == 0:
        table.copy_(torch.from_numpy(np.concatenate([synth.clique(48, n, B=8, seed=5), synth.chain(24, n, B=8, seed=6)])))
    par.broadcast_table(dist, table, src=0)
    x = table.numpy()
    ora = fwo.Oracle(x.T, "fz")
    ora.compute_cor()
    off, nbr, st, ap = ora.pairwise(alpha=0.01, n_obs_min=20)
    uni = fw.NbrCSR(off, nbr, st, ap)
    order = fw.target_order(uni)
    shard = par.shard_targets(order, rank, world)
    lists, nts = [], []
    for T in shard:
        a, b = off[T], off[T + 1]
        wn, ws, wp, wt = ora.hiton_pc(int(T), nbr[a:b], st[a:b], ap[a:b], max_k=3, alpha=0.01, n_obs_min=20)
        lists.append((wn, ws, wp)); nts.append(wt)
    packed = par.pack_result(_OracleResult(shard, lists, nts))
    bucket = par.gather_results(dist, packed, dst=0)
    if rank == 0:
        merged = par.MergedResult(bucket)
        edges = fw.assemble_graph(merged, uni, "fz")
        q.put((sorted(int(t) for t in merged.targets), edges, int(merged.num_tests.sum()), x.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    targets, edges, ntests, x = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert targets == list(range(72))                       # every target processed exactly once
    single = fwo.Oracle(x.T, "fz").lgl(max_k=3, mode="single")
    assert [(a, b) for a, b, _ in edges] == [(a, b) for a, b, _ in single["edges"]]
    assert np.allclose([w for _, _, w in edges], [w for _, _, w in single["edges"]], rtol=0, atol=0)
    assert ntests == single["cond_tests"]


def test_shard_targets_partition():
    par = fwload.load_sub("parallel")
    order = np.random.default_rng(0).permutation(1001)
    for world in (1, 2, 4, 8):
        parts = [par.shard_targets(order, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(1001))
        assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1


def _records_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = fwload.load_sub("parallel")
    rng = np.random.default_rng(100 + rank)
    n = [5, 0, 17][rank]                                    # ragged, one rank with no records
    rec = {"x": rng.integers(0, 2 ** 31 - 2, n).astype(np.int32), "y": rng.integers(0, 2 ** 31 - 2, n).astype(np.int32),
           "p": rng.random(n) * 1e-300, "stat": rng.standard_normal(n), "n_reliable": 1000 + rank}
    got = par.allgather_records(dist, rec)
    q.put((rank, rec, got))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_records_ragged():
    """the exchange step of the sharded pairwise stage (parallel.allgather_records): every rank ends up with every rank's records,
    bit for bit (indices up to 2^31, denormal p-values), including an empty list"""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_records_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sent = {r: rec for r, rec, _ in out}
    for _, _, got in out:
        assert len(got) == world
        for r in range(world):
            for k in ("x", "y", "p", "stat"):
                assert got[r][k].dtype == sent[r][k].dtype and (got[r][k] == sent[r][k]).all()
            assert got[r]["n_reliable"] == sent[r]["n_reliable"]


def _table_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = fwload.load_sub("parallel")
    p, n = 37, 11                                           # 37 columns over 3 ranks: ragged slices
    full = torch.arange(p * n, dtype=torch.float32).reshape(p, n)
    c0, c1 = par.table_slice(p, rank, world)
    dev = torch.full((p, n), -1.0)
    par.upload_and_gather_table(dist, dev, full[c0:c1].clone(), rank, world)
    q.put((rank, bool((dev == full).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_upload_and_gather_table_ragged():
    """every rank contributes its column slice, every rank ends up with the whole table (parallel.upload_and_gather_table)"""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_table_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert sorted(r for r, _ in out) == [0, 1, 2] and all(ok for _, ok in out)
