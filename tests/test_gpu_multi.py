"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): the row-sharded cor_mat (balanced tile-row groups, two in-place
NCCL all-gathers, symmetrise) is bit-identical to the single-GPU fw_cor_matrix, and target-sharded HITON-PC merges to the
single-GPU graph."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import fwload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fw = fwload.load(); par = fwload.load_sub("parallel"); synth = fwload.load_sub("synth")
    p, n = 1100, 700                       # 9 tile rows -> padded to a multiple of 2*world
    d_x = torch.empty((p, n), dtype=torch.float32, device="cuda")
    if rank == 0:
        d_x.copy_(torch.from_numpy(np.concatenate([synth.clique(600, n, B=12, seed=1), synth.chain(500, n, B=20, seed=2)])))
    par.broadcast_table(dist, d_x, src=0)
    torch.cuda.synchronize()
    eng = fw.Engine(rank)
    eng.adopt_data_device(d_x.data_ptr(), n, p, "fz")
    h, nb_pad = par.cor_groups((p + 127) // 128, world)
    d_cor = torch.full((nb_pad * 128, p), float("nan"), dtype=torch.float32, device="cuda")
    eng.adopt_cor_device_rows(d_cor.data_ptr(), p, d_cor.shape[0])
    par.sharded_cor(dist, eng, d_cor)
    eng.synchronize()
    sharded = d_cor[:p].cpu().numpy()
    # single-GPU result on the same device
    eng1 = fw.Engine(rank)
    eng1.adopt_data_device(d_x.data_ptr(), n, p, "fz")
    single = eng1.cor()
    same = bool((sharded == single).all())
    # target-sharded HITON-PC on the sharded cor_mat
    uni = eng.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
    order = fw.target_order(uni)
    res = eng.si_HITON_PC(par.shard_targets(order, rank, world), max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
    bucket = par.gather_results(dist, par.pack_result(res), dst=0)
    if rank == 0:
        merged = par.MergedResult(bucket)
        edges = fw.assemble_graph(merged, uni, "fz")
        eng1.pw_univar_neighbors(alpha=0.01, n_obs_min=20)
        full = eng1.si_HITON_PC(order, max_k=3, alpha=0.01, n_obs_min=20, want_tpc=False)
        edges1 = fw.assemble_graph(full, uni, "fz")
        q.put((same, float(np.abs(sharded - single).max()), edges == edges1, int(merged.num_tests.sum()), int(full.num_tests.sum())))
    else:
        q.put((same, 0.0, True, 0, 0))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cor_and_hiton_two_gpus():
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
    outs = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for same, err, eq, nt, nt1 in outs:
        assert same, "sharded cor_mat differs from the single-GPU one (max |diff| %g)" % err
        assert eq and nt == nt1
