"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise) of the library's group path (include/fwgpu.h "multi-GPU",
csrc/comm.cuh): every rank uploads its columns, fw_multi_cor computes its tile rows reading the peers' slices over NVLink and
leaves cor_mat row-sharded, fw_pairwise pulls the peers' candidate lists, fw_hiton_pc reads the owners' shards through peer
mappings.  Everything must be bit-identical to one GPU: cor_mat (fw_cor_gather), neighbour lists, PC sets, graph.
Two set-ups: one process per GPU (CUDA IPC handles exchanged through torch.distributed) and two contexts in one process (plain
peer access, one host thread per context)."""
import os
import socket
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P, N = 1100, 700                       # 9 tile rows -> groups of 3 (world 2); the last group is short


def _table(synth):
    return np.ascontiguousarray(np.concatenate([synth.clique(600, N, B=12, seed=1), synth.chain(500, N, B=20, seed=2)]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _single(fw, x, dev=0):
    eng = fw.Engine(dev)
    eng.set_data_colmajor(x, "fz")
    cor = eng.cor()
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    return cor, uni, order, res, fw.assemble_graph(res, uni, "fz"), eng.pairwise_stats()


def _rank_pass(fw, par, eng, x, rank, world, prefetch):
    """one pass of the group pipeline on this rank; returns (cor sub-matrix, uni, packed PC lists of the shard)"""
    c0, c1 = par.table_slice(P, rank, world)
    eng.pairwise_prefetch(0.01 if prefetch else 0.0, 20)
    xs = np.ascontiguousarray(x[c0:c1])
    eng.multi_set_data_ptr(xs.ctypes.data, N, P)
    eng.multi_cor()
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(par.shard_targets(order, rank, world), max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    eng.synchronize()
    return eng.cor_gather(np.arange(P)), uni, par.pack_result(res), eng.pairwise_stats()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import fwload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fw = fwload.load(); par = fwload.load_sub("parallel"); synth = fwload.load_sub("synth")
    x = _table(synth)
    eng = fw.Engine(rank)
    par.attach_group(dist, eng, N, P)
    out = []
    for prefetch in (True, False, True):                 # repeated passes: the barrier protocol must hold across passes
        cor, uni, packed, stats = _rank_pass(fw, par, eng, x, rank, world, prefetch)
        bucket = par.gather_results(dist, packed, dst=0)
        if rank == 0:
            merged = par.MergedResult(bucket)
            out.append((cor, uni, fw.assemble_graph(merged, uni, "fz"), int(merged.num_tests.sum()), stats))
    if rank == 0:
        cor1, uni1, order1, res1, edges1, stats1 = _single(fw, x, 0)
        ok = True
        for cor, uni, edges, nt, stats in out:
            ok &= bool((cor == cor1).all()) and bool((uni.offsets == uni1.offsets).all()) and bool((uni.nbr == uni1.nbr).all())
            ok &= bool((uni.stat == uni1.stat).all()) and bool((uni.pval == uni1.pval).all()) and edges == edges1
            ok &= nt == int(res1.num_tests.sum()) and stats == stats1
        q.put((ok, float(np.abs(out[0][0] - cor1).max()), len(edges1)))
    dist.barrier()
    eng.comm_detach()
    dist.destroy_process_group()


def test_group_two_processes_ipc():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    ok, err, ne = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok, "group result differs from the single-GPU one (max |cor diff| %g)" % err
    assert ne > 500


def test_group_one_process_two_contexts():
    import torch
    import fwload
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    fw = fwload.load(); par = fwload.load_sub("parallel"); synth = fwload.load_sub("synth")
    x = _table(synth)
    world = 2
    engs = [fw.Engine(r) for r in range(world)]
    blobs = [engs[r].comm_export(r, world, N, P) for r in range(world)]
    allh = np.concatenate(blobs)
    for e in engs:
        e.comm_attach(allh)
    outs, errs = [None] * world, []

    def run(r):
        try:
            outs[r] = _rank_pass(fw, par, engs[r], x, r, world, True)
        except Exception as ex:          # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errs, errs
    cor1, uni1, order1, res1, edges1, stats1 = _single(fw, x, 0)
    for r in range(world):
        cor, uni, packed, stats = outs[r]
        assert (cor == cor1).all() and (uni.offsets == uni1.offsets).all() and (uni.nbr == uni1.nbr).all() and (uni.stat == uni1.stat).all()
        assert stats == stats1
    merged = par.MergedResult([outs[r][2] for r in range(world)])
    assert fw.assemble_graph(merged, uni1, "fz") == edges1
    for e in engs:
        e.comm_detach()


def _worker_nz(rank, world, port, q):
    """table-based kinds: every rank holds the table, the pairwise stage is split by X (NCCL all-gather of the records on the
    device), targets are sharded"""
    import torch
    import torch.distributed as dist
    import fwload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fw = fwload.load(); par = fwload.load_sub("parallel"); synth = fwload.load_sub("synth")
    ok = True
    for kind in ("fz_nz", "mi"):
        x = synth.hetero(2600, 500, B=12, seed=21)[0] if kind == "fz_nz" else synth.binarize(synth.clique(2600, 400, B=10, seed=22))
        nom = 20 if kind == "fz_nz" else fw.auto_n_obs_min("mi", 3, 5, max_level=2)
        eng = fw.Engine(rank)
        if kind == "fz_nz":
            # the table once over NVLink: this rank uploads its column slice, the slices are broadcast from their owners
            c0, c1 = par.table_slice(x.shape[0], rank, world)
            dev = torch.empty(x.shape, dtype=torch.float32, device=torch.device("cuda", rank))
            par.upload_and_gather_table(dist, dev, torch.from_numpy(np.ascontiguousarray(x[c0:c1])), rank, world)
            ok &= bool((dev.cpu().numpy() == x).all())
            eng.adopt_data_device(dev.data_ptr(), x.shape[1], x.shape[0], kind)
        else:
            eng.set_data_colmajor(x, kind)
        uni = par.sharded_pairwise(dist, eng, kind, alpha=0.01, n_obs_min=nom, device=torch.device("cuda", rank), want_host=True)
        order = fw.target_order(uni)
        res = eng.si_HITON_PC(par.shard_targets(order, rank, world), max_k=3, alpha=0.01, n_obs_min=nom, want_tpc=False, kind=kind)
        bucket = par.gather_results(dist, par.pack_result(res), dst=0)
        if rank == 0:
            merged = par.MergedResult(bucket)
            edges = fw.assemble_graph(merged, uni, kind)
            e1 = fw.Engine(0)
            e1.set_data_colmajor(x, kind)
            uni1 = e1.pw_univar_neighbors(alpha=0.01, n_obs_min=nom)
            res1 = e1.si_HITON_PC(fw.target_order(uni1), max_k=3, alpha=0.01, n_obs_min=nom, want_tpc=False)
            ok &= bool((uni.offsets == uni1.offsets).all()) and bool((uni.nbr == uni1.nbr).all()) and bool((uni.stat == uni1.stat).all()) and bool((uni.pval == uni1.pval).all())
            ok &= edges == fw.assemble_graph(res1, uni1, kind) and int(merged.num_tests.sum()) == int(res1.num_tests.sum()) and len(edges) > 100
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pairwise_two_processes_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_nz, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    ok = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok, "sharded pairwise stage + sharded targets differ from the single-GPU result"
