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
