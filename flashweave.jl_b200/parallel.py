"""Multi-GPU plumbing: one process per GPU, static target sharding, one broadcast of the
table, no further collectives; per-target neighbour lists are gathered on rank 0 and merged with the OR rule.

Replaces the job/result RemoteChannels of src/interleaved.jl for parallel="single" semantics (targets are
independent given the pairwise stage, src/learning.jl:137-138).  `dist` is torch.distributed (nccl on GPUs,
gloo in the CPU tests); nothing here touches the compute path.
"""
import numpy as np


def shard_targets(order, rank, world):
    """target i of the degree-ordered list -> rank i mod world (load balance: neighbouring degrees spread over ranks)"""
    return np.ascontiguousarray(np.asarray(order)[rank::world])


def broadcast_table(dist, tensor, src=0):
    """the one collective of the data path: the OTU table, once, over NVLink (NCCL) — or gloo on CPU"""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def pack_result(res):
    """HitonResult -> plain dict of the valid entries (what travels back to the host that assembles the graph)"""
    out = {"targets": np.asarray(res.targets).copy(), "num_tests": np.asarray(res.num_tests).copy(), "pc": []}
    for i in range(len(res.targets)):
        nb, st, pv = res.pc(i)
        out["pc"].append((np.asarray(nb).copy(), np.asarray(st).copy(), np.asarray(pv).copy()))
    return out


def gather_results(dist, packed, dst=0):
    """gather the per-rank packed results on `dst` (KBs per target; not a data-path collective)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [packed]
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(packed, bucket, dst=dst)
    return bucket


class MergedResult:
    """per-target PC lists of all ranks, with the accessors assemble_graph() uses"""

    def __init__(self, packed_list):
        self.targets, self._pc, nt = [], [], []
        for pk in packed_list:
            self.targets.extend(int(t) for t in pk["targets"])
            self._pc.extend(pk["pc"])
            nt.extend(int(x) for x in pk["num_tests"])
        self.targets = np.asarray(self.targets, np.int64)
        self.num_tests = np.asarray(nt, np.int64)

    def pc(self, i):
        return self._pc[i]


# ---- the GPUs of one node as a group (include/fwgpu.h "multi-GPU"; csrc/comm.cuh) ------------------------------------------
def exchange_handles(dist, blob):
    """all-gather the ranks' fixed-size fw_comm_export blobs (a few hundred bytes; setup, not data path)"""
    world = dist.get_world_size()
    bucket = [None] * world
    dist.all_gather_object(bucket, bytes(bytearray(np.asarray(blob, np.uint8))))
    return np.frombuffer(b"".join(bucket), np.uint8).copy()


def attach_group(dist, eng, n, p):
    """fw_comm_export on every rank, exchange of the handles, fw_comm_attach: afterwards Engine.multi_set_data_ptr /
    multi_cor / pw_univar_neighbors / si_HITON_PC work on the row-sharded cor_mat without any further collective."""
    blob = eng.comm_export(dist.get_rank(), dist.get_world_size(), n, p)
    eng.comm_attach(exchange_handles(dist, blob))
    dist.barrier()                                   # every rank has mapped its peers before anyone starts a pass


def table_slice(p, rank, world):
    """columns [p*rank/world, p*(rank+1)/world) of the table belong to `rank` (the library's partition)"""
    return p * rank // world, p * (rank + 1) // world


def upload_and_gather_table(dist, dev_table, host_slice, rank, world):
    """The table once over NVLink instead of N times over PCIe (table-based kinds: every rank needs all of it): rank r copies its
    columns [p*r/N, p*(r+1)/N) from (pinned) host memory into its place in `dev_table` ([p, n] CUDA tensor), then every slice is
    broadcast from its owner (NCCL; slices may differ in length, which all_gather_into_tensor does not allow).  Returns after the
    table is complete on this rank."""
    import torch
    p = dev_table.shape[0]
    c0, c1 = table_slice(p, rank, world)
    dev_table[c0:c1].copy_(host_slice, non_blocking=True)
    for r in range(world):
        a, b = table_slice(p, r, world)
        dist.broadcast(dev_table[a:b], src=r)
    if dev_table.is_cuda:
        torch.cuda.synchronize(dev_table.device)


# ---- pairwise stage of the table-based kinds (mi, mi_nz, fz_nz) split over the ranks ---------------------------------------------
def pack_records(rec):
    """the raw-significant records of one rank as ONE float64 array [4, n] (x, y, p, stat; indices are exact in float64)"""
    return np.stack([rec["x"].astype(np.float64), rec["y"].astype(np.float64), rec["p"], rec["stat"]])


def unpack_records(arr, n_reliable):
    return {"x": arr[0].astype(np.int32), "y": arr[1].astype(np.int32), "p": np.ascontiguousarray(arr[2]), "stat": np.ascontiguousarray(arr[3]),
            "n_reliable": int(n_reliable)}


def allgather_records(dist, rec, device=None):
    """The one exchange step of the sharded pairwise stage (Benjamini-Hochberg needs every p-value, statfuns.jl:326-350): all ranks'
    records on every rank.  Two collectives: the counts, then the records padded to the longest list (NCCL on `device`, gloo on CPU)."""
    import torch
    world = dist.get_world_size()
    meta = torch.tensor([len(rec["x"]), rec["n_reliable"]], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    counts = [int(m[0]) for m in metas]
    nmax = max(max(counts), 1)
    mine = torch.zeros((4, nmax), dtype=torch.float64, device=device)
    mine[:, :counts[dist.get_rank()]] = torch.from_numpy(pack_records(rec)).to(mine.device)
    bucket = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(bucket, mine)
    return [unpack_records(b[:, :c].cpu().numpy(), int(m[1])) for b, c, m in zip(bucket, counts, metas)]


def sharded_pairwise(dist, eng, kind, alpha=0.01, hps=5, n_obs_min=0, FDR=True, correct_reliable_only=True, device=None, want_host=False):
    """pw_univar_neighbors by all ranks together (every rank holds the table): partial -> all-gather -> merge; identical lists on every
    rank.  With a CUDA `device` the records never leave the GPUs: they are copied device-to-device into torch buffers, all-gathered by
    NCCL over NVLink and handed to fw_pairwise_merge as device pointers."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None or torch.device(device).type != "cuda":
        rec = eng.pairwise_partial(rank, world, alpha=alpha, hps=hps, n_obs_min=n_obs_min, correct_reliable_only=correct_reliable_only, kind=kind)
        recs = allgather_records(dist, rec, device=device)
        return eng.pairwise_merge(recs, alpha=alpha, FDR=FDR, correct_reliable_only=correct_reliable_only, kind=kind, want_host=want_host)
    n_raw, n_rel = eng.pairwise_partial_run(rank, world, alpha=alpha, hps=hps, n_obs_min=n_obs_min, correct_reliable_only=correct_reliable_only, kind=kind)
    meta = torch.tensor([n_raw, n_rel], dtype=torch.int64, device=device)
    metas = torch.zeros((world, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.cpu().numpy()
    counts = [int(c) for c in metas[:, 0]]
    nmax = max(max(counts), 1)
    mine_i = torch.zeros((2, nmax), dtype=torch.int32, device=device)
    mine_f = torch.zeros((2, nmax), dtype=torch.float64, device=device)
    torch.cuda.synchronize(device)                   # the buffers exist (torch's stream) before the library's stream writes them
    eng.pairwise_partial_copy_ptrs(mine_i[0].data_ptr(), mine_i[1].data_ptr(), mine_f[0].data_ptr(), mine_f[1].data_ptr())
    all_i = torch.empty((world, 2, nmax), dtype=torch.int32, device=device)
    all_f = torch.empty((world, 2, nmax), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(all_i, mine_i)
    dist.all_gather_into_tensor(all_f, mine_f)
    xs = torch.cat([all_i[r, 0, :counts[r]] for r in range(world)]).contiguous()
    ys = torch.cat([all_i[r, 1, :counts[r]] for r in range(world)]).contiguous()
    ps = torch.cat([all_f[r, 0, :counts[r]] for r in range(world)]).contiguous()
    ss = torch.cat([all_f[r, 1, :counts[r]] for r in range(world)]).contiguous()
    torch.cuda.synchronize(device)
    m = int(metas[:, 1].sum()) if correct_reliable_only else eng.p * (eng.p - 1) // 2
    eng.pairwise_merge_ptrs(int(sum(counts)), xs.data_ptr(), ys.data_ptr(), ps.data_ptr(), ss.data_ptr(), m, alpha=alpha, FDR=FDR, kind=kind)
    return eng.univar_nbrs() if want_host else None


# ---- row-sharded correlation matrix with a host-side NCCL exchange (superseded by the group API above) -------------------------------------------------------------------------------------
def cor_groups(n_tile_rows, world):
    """Split the tile rows of the upper-triangular cor_mat GEMM into 2*world equal contiguous groups; rank r owns group r
    (long tile rows) and group 2*world-1-r (short ones), so every rank computes the same number of 128x128 tiles.
    Returns (rows_per_group, padded tile-row count)."""
    g = 2 * world
    h = (n_tile_rows + g - 1) // g
    return h, h * g


def sharded_cor(dist, eng, cor_tensor, rev_group=None):
    """cor_mat = Float32.(cor(data)) computed by all ranks together (learning.jl:42-44): every rank computes the upper tiles of
    its two tile-row groups into `cor_tensor` (a [rows_pad, p] float32 CUDA tensor adopted by the engine), the row blocks are
    exchanged by two all-gathers over NVLink (the top half in place, the bottom half into views placed in reversed rank
    order), and the lower triangle is filled from the upper one.  Bit-identical to Engine.cor() on one GPU."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    nb = eng.cor_prepare()
    h, nb_pad = cor_groups(nb, world)
    rows_pad, p = cor_tensor.shape
    assert rows_pad >= nb_pad * 128, (rows_pad, nb_pad)
    eng.cor_rows(rank * h, (rank + 1) * h)
    gb = 2 * world - 1 - rank
    eng.cor_rows(gb * h, (gb + 1) * h)
    eng.synchronize()
    hr = h * 128
    top = cor_tensor[: world * hr]
    bot = cor_tensor[world * hr: 2 * world * hr]
    dist.all_gather_into_tensor(top, top[rank * hr:(rank + 1) * hr])
    # rank i owns bottom group world-1-i: receive its block at that position
    views = [bot[(world - 1 - i) * hr:(world - i) * hr] for i in range(world)]
    dist.all_gather(views, views[rank])
    torch.cuda.synchronize()
    eng.cor_symmetrize()
