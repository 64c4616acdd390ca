#!/usr/bin/env python
"""bench.py — CI-tests/sec of the HITON-PC conditional phase on the BASELINE.json workloads.

Default workload (config C4 of BASELINE.json / SURVEY.md §8d): 50 000 OTUs x 10 000 samples, synthetic "clique-B" table
(B = 24, seed 20190802+3), sensitive=true (Fisher-z), max_k = 3, alpha = 0.01.  `--config C2|C3|C5` selects the other GPU
configurations of BASELINE.json (C2: 10 000 x 2 000 fz, C3: 10 000 x 2 000 mi, C5: 50 000 OTUs + 10 meta variables x 10 000
heterogeneous fz_nz).  One "step" = one pass of the hot path over the whole table: si_HITON_PC (interleaving + elimination, all
conditioning subsets) for every target variable.  The metric is quoted on a fixed table at 1/2/4/8 GPUs (BASELINE.json), so
the total work is fixed and the degree-ordered targets are dealt round-robin to the ranks (target i -> rank i mod N, as
interleaved.jl hands targets to workers): scaling = "strong".  No collective in the timed `value` region.

  value : cond_tests_ref/s, device-resident (cor_mat / bit planes + neighbour lists already in HBM), CUDA events on the
          library's stream, max over ranks.  cond_tests_ref = sum of test_subsets' num_tests exactly as the reference counts
          them (src/tests.jl:322, early exit honoured).
  e2e   : the same count divided by the time of the whole pipeline through the C ABI from HOST buffers: H2D of the table
          from pinned host memory, cor_mat, pairwise stage + BH, HITON-PC of the shard, D2H of the neighbour lists.
          fz, N = 1: column chunks of the upload hidden behind the GEMM (fw_upload_cor_f32).  fz, N > 1: the library's group path (include/fwgpu.h "multi-GPU"): every rank uploads 1/N of
          the columns, the standardising kernel reads the peers' slices over NVLink, cor_mat stays row-sharded and is read
          through peer mappings; no NCCL call in the data path (torch.distributed only sets the group up and reduces the
          timings).  Other kinds, N > 1: every rank needs the whole table - fz_nz: each rank uploads 1/N of the columns and the slices are
          broadcast once with NCCL over NVLink (mi: the 80 MB table is uploaded by every rank) -, pairwise stage split by X with one
          NCCL all-gather of the raw-significant records for the global BH step, targets sharded.
  parity_sample : the oracle re-runs a sample of the targets on the engine's own inputs and must reproduce the engine's PC sets,
          statistics and test counts; a mismatch fails the run.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads; cor(data) through BLAS as the
reference does, src/learning.jl:44) on a bounded sample of the same workload (the reference itself is Julia: cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_RESULT_FD = None

CONFIGS = {
    # name: (p, n, kind, seed index, description)
    "C2": (10000, 2000, "fz", 1, "C2 clique-B synthetic: 10000 OTUs x 2000 samples, fz (sensitive=true)"),
    "C3": (10000, 2000, "mi", 2, "C3 clique-B synthetic, binarised at the median: 10000 OTUs x 2000 samples, mi (sensitive=false)"),
    "C4": (50000, 10000, "fz", 3, "C4 clique-B synthetic: 50000 OTUs x 10000 samples, fz (sensitive=true)"),
    "C5": (50010, 10000, "fz_nz", 4, "C5 heterogeneous synthetic (8 habitats, 10 % dropout): 50000 OTUs + 10 meta variables x 10000 samples, fz_nz (FlashWeaveHE-S)"),
}


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--p", type=int, default=0, help="override the number of variables (experiments)")
    ap.add_argument("--n", type=int, default=0, help="override the number of samples (experiments)")
    ap.add_argument("--B", type=int, default=24)
    ap.add_argument("--max-k", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=0.01)
    ap.add_argument("--cpu-blocks", type=int, default=256, help="blocks of the table in the CPU-baseline sample")
    ap.add_argument("--parity-blocks", type=int, default=64, help="blocks whose targets the oracle re-checks against the engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prefetch", action="store_true", help="collect the pairwise candidates in the cor_mat GEMM epilogue (fw_pairwise_prefetch) instead of one scan of the matrix")
    a = ap.parse_args()
    p, n, kind, si, desc = CONFIGS[a.config]
    a.p, a.n, a.kind, a.seed_idx, a.desc = a.p or p, a.n or n, kind, si, desc
    return a


def workload_name(a):
    return "%s, max_k=%d, B=%d, alpha=%g" % (a.desc if (a.p, a.n) == CONFIGS[a.config][:2] else a.desc + " [shape overridden: %d x %d]" % (a.p, a.n),
                                            a.max_k, a.B, a.alpha)


def make_table(a, synth, p=None):
    """the synthetic table of the configuration ([p, n]); a smaller p gives the first blocks of the same table"""
    p = p or a.p
    seed = synth.BASE_SEED + a.seed_idx
    if a.kind == "fz":
        return synth.clique(p, a.n, B=a.B, seed=seed)
    if a.kind == "mi":
        return synth.binarize(synth.clique(p, a.n, B=a.B, seed=seed))
    if a.kind == "fz_nz":
        return synth.hetero(p, a.n, B=a.B, seed=seed)[0]
    raise ValueError(a.kind)


def n_obs_min_of(a, fw):
    return fw.auto_n_obs_min(a.kind, a.max_k, 5, max_level=2) if a.kind in ("mi", "mi_nz") else 20


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "which": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1600.0, "bf16_sustained": 1400.0, "which": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.proc, self.lines = dev, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def bind_to_gpu_numa(dev):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off, as numactl / NCCL would: the
    H2D copy of a rank's table slice then does not cross the socket interconnect.  Returns (previous affinity, node) or (None, None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(dev)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) > 4:
            bus = bus[-12:]                                          # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus:
            return None, None
        os.sched_setaffinity(0, cpus)
        return prev, node
    except Exception:
        return None, None


def blas_cor_f32(x_pn, threads):
    """cor(data) the way the reference evaluates it (src/learning.jl:44: Statistics.cor of a Float32 matrix = centred columns,
    one BLAS Gram product in Float32, cov2cor!), with all host threads."""
    try:
        from threadpoolctl import threadpool_limits
        ctxm = threadpool_limits(limits=threads)
    except Exception:
        import contextlib
        ctxm = contextlib.nullcontext()
    with ctxm:
        xc = x_pn - x_pn.mean(axis=1, keepdims=True, dtype=np.float32)
        nrm = np.sqrt(np.einsum("ij,ij->i", xc, xc, dtype=np.float32))
        c = xc @ xc.T
        c /= nrm[:, None]
        c /= nrm[None, :]
    np.clip(c, -1.0, 1.0, out=c)
    np.fill_diagonal(c, 1.0)
    return c


def cpu_reference_run(a, steps, warmup, blocks, x=None):
    """The oracle ("port" of the reference) on a bounded sample: the first `blocks` blocks of the same table (+ the meta variables
    for C5).  Blocks are independent in these workloads, so the per-target work is the same as in the full table."""
    import fwload
    from oracle import fwo
    fw = fwload.load()
    synth = fwload.load_sub("synth")
    n_meta = 10 if a.kind == "fz_nz" else 0
    p_s = min(a.p - n_meta, blocks * a.B)
    if x is None:
        x = make_table(a, synth, p_s + n_meta)
    elif n_meta:
        x = np.ascontiguousarray(np.concatenate([x[:p_s], x[a.p - n_meta:]]))
    else:
        x = np.ascontiguousarray(x[:p_s])
    threads = host_threads()          # torchrun exports OMP_NUM_THREADS=1; the oracle sets its own thread count
    nom = n_obs_min_of(a, fw)
    rates, e2e_rates, secs_all, tests = [], [], [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ora = fwo.Oracle(x.T, a.kind, cont32=True)
        t_cor = 0.0
        if a.kind == "fz":
            tc = time.perf_counter()
            ora.set_cor(blas_cor_f32(x, threads).astype(np.float64))
            t_cor = time.perf_counter() - tc
        r = ora.lgl(max_k=a.max_k, alpha=a.alpha, n_obs_min=nom, mode="single", n_threads=threads)
        t1 = time.perf_counter()
        if it >= warmup:
            tests = r["cond_tests"]
            rates.append(r["cond_tests"] / r["secs"]["hiton"])
            e2e_rates.append(r["cond_tests"] / (t1 - t0))
            secs_all.append(dict(r["secs"], cor=t_cor, total=t1 - t0))
        del ora
    return {"cond_rate": float(np.mean(rates)), "e2e_rate": float(np.mean(e2e_rates)), "threads": threads, "p_sample": int(x.shape[0]), "tests": tests,
            "secs": {k: float(np.mean([s[k] for s in secs_all])) for k in secs_all[0]}}


def sample_text(a, r, blocks):
    return ("first %d of %d blocks (%d variables x %d samples) of the same table; cor(data) [BLAS, Float32] + pairwise/BH + HITON-PC, %d conditional tests per step"
            % (min(blocks, (a.p + a.B - 1) // a.B), (a.p + a.B - 1) // a.B, r["p_sample"], a.n, r["tests"]))


def main_reference(a, rank):
    if rank != 0:
        return
    r = cpu_reference_run(a, a.steps, a.warmup, a.cpu_blocks)
    sample = sample_text(a, r, a.cpu_blocks)
    line = {
        "impl": "reference", "metric": "CI-tests/sec (cond_tests_ref/s, HITON-PC conditional phase)", "value": r["e2e_rate"], "unit": "tests/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["secs"]["total"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": r["e2e_rate"], "unit": "tests/s", "cores": r["threads"], "kind": "port", "sample": sample,
                         "cond_phase_value": r["cond_rate"], "secs": r["secs"],
                         "note": "C++/OpenMP restatement of FlashWeave.jl semantics (oracle/), not the Julia package: Julia is not installed; "
                                 "value = whole pipeline (as the GPU arm's e2e), cond_phase_value = conditional phase only (as the GPU arm's value)"},
        "e2e": {"value": r["e2e_rate"], "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main_ours(a, rank, world, local_rank):
    import torch
    import fwload
    fw = fwload.load()
    synth = fwload.load_sub("synth")
    par = fwload.load_sub("parallel")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    prev_affinity, numa_node = bind_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    p, n, kind = a.p, a.n, a.kind
    nom = n_obs_min_of(a, fw)
    np_dtype, t_dtype = (np.int32, torch.int32) if kind == "mi" else (np.float32, torch.float32)
    # ---- setup (untimed): synthetic table in pinned host memory ---------------------------------------------------
    # N = 1: a pinned buffer.  N > 1: one copy in POSIX shared memory (what FlashWeave's SharedArray gives its local workers,
    # src/learning.jl:553-560), page-locked by every rank: each rank uploads its own columns over its own PCIe link.
    gen_s = 0.0
    shm_path = None
    if world > 1:
        shm_path = "/dev/shm/fw_bench_table_%s.bin" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            t0 = time.time()
            mm = np.memmap(shm_path, mode="w+", shape=(p, n), dtype=np_dtype)
            mm[:] = make_table(a, synth)
            mm.flush(); del mm
            gen_s = time.time() - t0
        dist.barrier()
        mm = np.memmap(shm_path, mode="r+", shape=(p, n), dtype=np_dtype)
        host_x = torch.from_numpy(mm)
        rc = torch.cuda.cudart().cudaHostRegister(host_x.data_ptr(), host_x.numel() * 4, 0)
        if int(rc) != 0:
            raise SystemExit("cudaHostRegister of the shared table failed: %s" % rc)
    else:
        t0 = time.time()
        host_x = torch.empty((p, n), dtype=t_dtype, pin_memory=True)
        host_x.numpy()[:] = make_table(a, synth)
        gen_s = time.time() - t0
    host_np = host_x.numpy()
    eng = fw.Engine(local_rank)
    group = dist is not None and kind == "fz"
    slice_host = None
    if group:
        par.attach_group(dist, eng, n, p)            # handle exchange, once (setup)
        # this rank's columns in its own pinned buffer (first-touched on the GPU's NUMA node): what a worker that owns 1/N of the table holds
        c0, c1 = par.table_slice(p, rank, world)
        slice_host = torch.empty((c1 - c0, n), dtype=t_dtype, pin_memory=True)
        slice_host.copy_(host_x[c0:c1])
    gather_tab = dist is not None and kind == "fz_nz"        # float table needed whole on every rank: upload 1/N each, NCCL broadcasts over NVLink
    dev_table = None
    if gather_tab:
        c0, c1 = par.table_slice(p, rank, world)
        slice_host = torch.empty((c1 - c0, n), dtype=t_dtype, pin_memory=True)
        slice_host.copy_(host_x[c0:c1])
        dev_table = torch.empty((p, n), dtype=t_dtype, device=torch.device("cuda", local_rank))
    ext = torch.cuda.ExternalStream(eng.stream)

    phase_wall = {"table_and_cor_ms": [], "pairwise_ms": [], "hiton_ms": [], "upload_ms": []}

    def pipeline():
        """e2e: host table -> neighbour lists of this rank's target shard, through the C ABI."""
        t0 = time.perf_counter()
        if kind == "fz":
            if a.prefetch:
                eng.pairwise_prefetch(a.alpha, nom)
            if not group:
                eng.upload_and_cor(host_x.data_ptr(), n=n, p=p)       # upload chunked and hidden behind the cor_mat GEMM
            else:
                eng.multi_set_data_ptr(slice_host.data_ptr(), n, p)
                phase_wall["upload_ms"].append((time.perf_counter() - t0) * 1e3)
                eng.multi_cor()
            eng.synchronize()
        elif kind == "mi":
            eng._ck(eng.L.fw_set_data_i32(eng.h, fw.C.c_void_p(host_x.data_ptr()), n, p, n))
            eng.kind, eng.n, eng.p = kind, n, p
        elif gather_tab:
            par.upload_and_gather_table(dist, dev_table, slice_host, rank, world)
            eng.adopt_data_device(dev_table.data_ptr(), n, p, kind)
        else:
            eng.set_data_ptr(host_x.data_ptr(), n, p, kind)
        t1 = time.perf_counter()
        if dist is not None and kind != "fz":
            # table-based kinds: every rank holds the table, the X variables of the pairwise stage are dealt to the ranks, one all-gather
            # of the raw-significant records (NCCL) for the global Benjamini-Hochberg step (fw_pairwise_partial / fw_pairwise_merge)
            par.sharded_pairwise(dist, eng, kind, alpha=a.alpha, n_obs_min=nom, device=torch.device("cuda", local_rank))
        else:
            eng.pw_univar_neighbors(alpha=a.alpha, n_obs_min=nom, want_host=False, kind=kind)
        off = np.zeros(p + 1, np.int64)
        eng._ck(eng.L.fw_pairwise_copy(eng.h, off.ctypes.data_as(fw.C.c_void_p), None, None, None))
        order = np.argsort(np.diff(off), kind="stable").astype(np.int64)      # learning.jl:97-98
        shard = par.shard_targets(order, rank, world)                          # target i -> rank i mod N
        t2 = time.perf_counter()
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=nom, want_tpc=False, reuse_buffers=True, kind=kind)
        t3 = time.perf_counter()
        phase_wall["table_and_cor_ms"].append((t1 - t0) * 1e3); phase_wall["pairwise_ms"].append((t2 - t1) * 1e3); phase_wall["hiton_ms"].append((t3 - t2) * 1e3)
        return shard, res

    # ---- e2e region ----------------------------------------------------------------------------------
    for _ in range(a.warmup):
        shard, res = pipeline()
    barrier()
    launches0 = eng.launch_count()
    e2e_t = []
    phase = {"cor_ms": [], "pairwise_ms": [], "hiton_ms": [], "cor_standardise_ms": [], "cor_barrier_wait_ms": []}
    for k in phase_wall:
        phase_wall[k].clear()
    for _ in range(a.steps):
        barrier()
        t0 = time.perf_counter()
        shard, res = pipeline()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_t.append(t1 - t0)
        lt = eng.last_timing()
        for k in phase:
            phase[k].append(lt[k])
    e2e_launches = (eng.launch_count() - launches0) / max(a.steps, 1)
    if world > 1:                                    # every rank's phases on stderr (skew between ranks)
        sys.stderr.write("[rank %d] e2e %.2f ms; device %s; wall %s\n" % (rank, float(np.mean(e2e_t)) * 1e3, {k: round(float(np.mean(v)), 2) for k, v in phase.items()},
                                                                          {k: round(float(np.mean(v)), 2) for k, v in phase_wall.items() if v}))
    tests_rank = int(res.num_tests.sum())
    if kind == "fz":
        h2d = p * n * 4 // world if group else p * n * 4
    elif gather_tab:
        h2d = slice_host.numel() * 4
    else:
        h2d = p * n * 4
    h2d += len(shard) * 8
    d2h = int(res.off[-1]) * 24 + len(shard) * 24 + (p + 1) * 8
    # the cor_mat GEMM alone (table already resident), for its tensor-core roofline
    gemm_ms = []
    if kind == "fz" and dist is None:
        eng.pairwise_prefetch(0.0, 0)
        for _ in range(3):
            eng.cor(want_host=False); eng.synchronize()
            gemm_ms.append(eng.last_timing()["cor_ms"])
        eng.pw_univar_neighbors(alpha=a.alpha, n_obs_min=nom, want_host=False)

    # ---- device-resident region (the contract's K timed steps) -------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()                                  # runs over the warm-up and the timed steps (and, when the timed region is shorter
                                                     # than a few sampling periods, over extra untimed steps of the same kernel, below)
    for _ in range(a.warmup):
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=nom, want_tpc=False, reuse_buffers=True, kind=kind)
    barrier()
    launches1 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(ext)
    kern_ms = []
    for _ in range(a.steps):
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=nom, want_tpc=False, reuse_buffers=True, kind=kind)
        kern_ms.append(eng.last_timing()["hiton_ms"])
    ev1.record(ext)
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / a.steps
    launches = eng.launch_count() - launches1
    t_extra = time.perf_counter()
    n_extra = 0
    while len(sampler.lines) < 4 and time.perf_counter() - t_extra < 1.0:        # untimed: keep the GPU under the same load until nvidia-smi has sampled it
        eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=nom, want_tpc=False, reuse_buffers=True, kind=kind)
        n_extra += 1
    clocks = sampler.stop()
    clocks["sampled_over"] = "warm-up + timed steps of the device-resident region" + (" + %d untimed extra steps of the same kernel" % n_extra if n_extra else "")
    exec_k = eng.hiton_exec_by_k()
    tests_exec = res.tests_executed

    # ---- parity sample: the oracle re-runs a sample of this rank's targets on the engine's own inputs --------------------
    parity = None
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)       # the oracle legs (parity sample, CPU baseline) use every host core
    if rank == 0 and a.parity_blocks > 0:
        from oracle import parity as opar
        uni = eng.univar_nbrs()
        lim = a.parity_blocks * a.B
        pos = np.nonzero(np.asarray(shard) < lim)[0]
        parity = opar.sampled_hiton_parity(eng, kind, host_np, pos, res, uni, a.max_k, a.alpha, nom, n_threads=host_threads())
        parity["sample"] = "this rank's targets among the first %d blocks (variables < %d)" % (a.parity_blocks, lim)
        if kind == "fz":
            # and against Float32(cor in fp64), i.e. without the tensor-core rounding of cor_mat: edge-set difference
            U = np.arange(min(lim, p))
            c64 = np.corrcoef(host_np[U].astype(np.float64)).astype(np.float32)
            parity["cor_mat_max_abs_err_vs_fp64"] = float(np.abs(eng.cor_gather(U) - c64).max())

    # ---- reduce over ranks: max time, sum of tests ----------------------------------------------------------
    stats = torch.tensor([dev_ms, float(np.mean(e2e_t)) * 1e3, float(tests_rank), float(tests_exec), float(launches), float(h2d), float(d2h)],
                         dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = stats, stats
    if world > 1:
        if group:
            eng.comm_detach()
        torch.cuda.cudart().cudaHostUnregister(host_x.data_ptr())
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    dev_ms_max, e2e_ms_max = mx[0].item(), mx[1].item()
    tests_total, exec_total = sm[2].item(), sm[3].item()
    value = tests_total / (dev_ms_max * 1e-3)
    e2e_value = tests_total / (e2e_ms_max * 1e-3)

    pk = peaks()
    k_ms = float(np.mean(kern_ms))
    if kind == "fz":
        alg_bytes = float(12 * exec_k[0] + 24 * exec_k[1] + 40 * exec_k[2])      # 4*C(k+2,2) B of correlations per test
        kname = "hiton_fz_kernel"
        note = "40 B of correlations per k=3 test: instruction-issue bound, not HBM bound (DESIGN.md 4.1; ncu: profiles/)"
    elif kind == "mi":
        alg_bytes = float(n) * float(3 * exec_k[0] + 4 * exec_k[1] + 5 * exec_k[2])   # (2+k)*n B of level codes per test
        kname = "hiton_mi_kernel"
        note = "(2+k)*n B of uint8 level codes per test in the reference's layout; the engine reads 1-bit planes (8x denser, L2-resident at this size), so achieved/peak may exceed 1"
    else:
        alg_bytes = None                                                         # SURVEY §8d a9 is per job, not per test: reported as null
        kname = "hiton_fz_kernel<NZ>"
        note = "fz_nz: (m+2)*n*4 B per (X,Y) job (SURVEY §8d a9); jobs are not counted by the kernel, achieved left null"
    tfile, traffic = os.path.join("profiles", "r02_traffic.json"), None
    if os.path.exists(os.path.join(ROOT, tfile)) and world == 1:
        try:
            traffic = json.load(open(os.path.join(ROOT, tfile))).get("%s_%s_dram_bytes_per_launch" % (kname.split("<")[0], a.config))
        except Exception:
            traffic = None
    ach = alg_bytes / (k_ms * 1e-3) / 1e9 if alg_bytes is not None else None
    roofline = {"kernel": "%s (si_HITON_PC conditional phase; dominant kernel of the timed `value` region)" % kname,
                "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"] if ach is not None else None, "traffic": traffic,
                "traffic_source": (tfile + " (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch)") if traffic else None,
                "peak_source": pk["which"], "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "note": note}
    line = {
        "metric": "CI-tests/sec (cond_tests_ref/s, HITON-PC conditional phase)", "value": value, "unit": "tests/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "targets_per_gpu": int(len(shard)), "sharding": "target i (ascending univariate degree) -> rank i mod N",
                   "l2": ("inputs larger than L2 (cor_mat %.1f GB, read by gather)" % (p * p * 4 / 1e9)) if kind == "fz" else
                         ("inputs larger than L2 (table %.1f GB)" % (p * n * 4 / 1e9) if kind == "fz_nz" else
                          "bit planes %.1f MB: L2-resident at this configuration (stated, see roofline.note)" % (p * ((n + 31) // 32) * 4 / 1e6)),
                   "cond_tests_ref_per_step": tests_total, "cond_tests_executed_per_step": exec_total,
                   "pairwise_tests": p * (p - 1) // 2, "parity_semantics": "parallel=\"single\" (SURVEY.md §3.6)"},
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": sm[5].item(), "d2h_bytes_per_step": sm[6].item(),
                "ms_per_step": e2e_ms_max, "gpu_launches_per_step": e2e_launches,
                "phases_device_ms_rank0": {k: float(np.mean(v)) for k, v in phase.items()},
                "phases_wall_ms_rank0": {k: float(np.mean(v)) for k, v in phase_wall.items() if v},
                "pairwise_tests_per_s": p * (p - 1) / 2 / (float(np.mean(phase_wall["pairwise_ms"])) * 1e-3),
                "multi_gpu_path": ("library group (CUDA IPC peer mappings over NVLink, row-sharded cor_mat, no NCCL in the data path)" if group else
                                   (("table uploaded 1/N per rank + NCCL broadcasts over NVLink, " if gather_tab else "replicated table upload, ") + "pairwise stage split by X (one NCCL all-gather of the raw-significant records), sharded targets" if world > 1 else "single GPU"))},
        "gpu_launches": int(sm[4].item()),
        "roofline": roofline,
        "clocks": clocks,
        "setup": {"table_gen_s": gen_s, "numa_node_of_gpu": numa_node, "host_cores_bound": host_threads()},
    }
    if kind == "fz":
        if dist is None:
            cor_ms, peak, pname = float(min(gemm_ms)), pk["bf16_burst"], "burst (kernel timed alone)"
            cname = "cor_mat GEMM (fw_cor_matrix on the resident table, timed alone; includes the standardise+split kernel)"
        else:
            cor_ms, peak, pname = float(np.mean(phase["cor_ms"])), pk["bf16_sustained"], "sustained (kernel timed inside the step)"
            cname = "cor_mat GEMM, this rank's 1/N of the tiles (fw_multi_cor: standardise from the peers' slices + row-sharded GEMM, device time)"
        line["roofline_cor_gemm"] = {"kernel": cname, "bound": "tensor", "achieved": 2.0 * n * p * p / world / (cor_ms * 1e-3) / 1e12, "peak": peak, "peak_kind": pname,
                                     "unit": "TFLOP/s per GPU", "frac": 2.0 * n * p * p / world / (cor_ms * 1e-3) / 1e12 / peak, "peak_source": pk["which"], "kernel_ms": cor_ms,
                                     "note": "useful flop = 2*n*p^2 (SURVEY §8d a1: the 3-term bf16 split and the symmetric half do not change the numerator); "
                                             "the MMAs issued are 1.5x that on half the matrix, so 0.66 is the ceiling of this fraction at 100 % tensor-pipe issue"}
    if parity is not None:
        line["parity_sample"] = parity
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(a, 1, 0, a.cpu_blocks, x=host_np)
        line["cpu_baseline"] = {"value": r["cond_rate"], "unit": "tests/s", "cores": r["threads"], "kind": "port",
                                "sample": sample_text(a, r, a.cpu_blocks) + "; value = conditional phase only (as `value`), e2e_value = whole pipeline (as `e2e`)",
                                "e2e_value": r["e2e_rate"], "secs": r["secs"],
                                "note": "C++/OpenMP restatement of FlashWeave.jl (oracle/), not the Julia package; omits Julia's per-test String/Vector allocations"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and parity["mismatches"]:
        raise SystemExit("parity_sample: %d of %d sampled targets differ from the oracle: %s" % (parity["mismatches"], parity["targets"], parity["first_mismatch"]))


def main():
    a = parse()
    # stdout carries exactly ONE line (the JSON result): libraries that print to fd 1 (NCCL's version banner at communicator
    # creation, torchrun notices) are sent to stderr for the whole run, the result line goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        main_reference(a, rank)
    else:
        main_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
