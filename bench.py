#!/usr/bin/env python
"""bench.py — CI-tests/sec of the HITON-PC conditional phase on the BASELINE.json workload.

Workload (config C4 of BASELINE.json / SURVEY.md §8d): 50 000 OTUs x 10 000 samples, synthetic
"clique-B" table (B = 24, seed 20190802+3), sensitive=true (Fisher-z), max_k = 3, alpha = 0.01.
One "step" = one pass of the hot path over the whole table: si_HITON_PC (interleaving +
elimination, all conditioning subsets) for every target variable.  The metric is quoted on this
fixed table at 1/2/4/8 GPUs (BASELINE.json), so the total work is fixed and the degree-ordered
targets are dealt round-robin to the ranks (target i -> rank i mod N, as interleaved.jl hands
targets to workers): scaling = "strong".  There is no data-path collective in the timed `value`
region (the table is broadcast once, inside the e2e region).

  value : cond_tests_ref/s, device-resident (cor_mat + neighbour lists already in HBM), CUDA events
          on the library's stream, max over ranks.  cond_tests_ref = sum of test_subsets' num_tests
          exactly as the reference counts them (src/tests.jl:322, early exit honoured).
  e2e   : the same count divided by the time of the whole pipeline through the C ABI from HOST
          buffers: H2D of the table from pinned host memory (N = 1: column chunks hidden behind the
          cor_mat GEMM, fw_upload_cor_f32; N > 1: each rank uploads 1/N of the shared table, NCCL
          all-gather, row-sharded GEMM + all-gather), pairwise stage + BH, HITON-PC of the shard,
          D2H of the neighbour lists.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a
bounded sample of the same workload (the reference itself is Julia and cannot run here).
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
    ap.add_argument("--p", type=int, default=50000)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--B", type=int, default=24)
    ap.add_argument("--max-k", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=0.01)
    ap.add_argument("--cpu-blocks", type=int, default=64, help="blocks of the table in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return "C4 clique-B synthetic: %d OTUs x %d samples, fz (sensitive=true), max_k=%d, B=%d, alpha=%g" % (a.p, a.n, a.max_k, a.B, a.alpha)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "which": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "which": "fallback (B200_PROFILING.md)"}


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


def cpu_reference_run(a, steps, warmup, blocks):
    """The oracle ("port" of the reference) on a bounded sample: the first `blocks` blocks of the same table
    (blocks are independent in the clique workload, so per-target work is the same as in the full table)."""
    import fwload
    from oracle import fwo
    synth = fwload.load_sub("synth")
    p_s = min(a.p, blocks * a.B)
    x = synth.clique(p_s, a.n, B=a.B, seed=synth.BASE_SEED + 3)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle sets its own thread count)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rates, e2e_rates, secs_all, tests = [], [], [], 0
    for it in range(warmup + steps):
        ora = fwo.Oracle(x.T, "fz", cont32=True)
        t0 = time.perf_counter()
        r = ora.lgl(max_k=a.max_k, alpha=a.alpha, mode="single", n_threads=threads)
        t1 = time.perf_counter()
        if it >= warmup:
            tests = r["cond_tests"]
            rates.append(r["cond_tests"] / r["secs"]["hiton"])
            e2e_rates.append(r["cond_tests"] / (t1 - t0))
            secs_all.append(dict(r["secs"], total=t1 - t0))
        del ora
    return {"cond_rate": float(np.mean(rates)), "e2e_rate": float(np.mean(e2e_rates)), "threads": threads, "p_sample": p_s, "tests": tests,
            "secs": {k: float(np.mean([s[k] for s in secs_all])) for k in secs_all[0]}}


def main_reference(a, rank):
    if rank != 0:
        return
    r = cpu_reference_run(a, a.steps, a.warmup, a.cpu_blocks)
    sample = ("first %d of %d blocks (%d OTUs x %d samples) of the same table; cor + pairwise + HITON-PC, %d conditional tests per step"
              % (a.cpu_blocks, (a.p + a.B - 1) // a.B, r["p_sample"], a.n, r["tests"]))
    line = {
        "impl": "reference", "metric": "CI-tests/sec (cond_tests_ref/s, HITON-PC conditional phase)", "value": r["e2e_rate"], "unit": "tests/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["secs"]["total"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": r["e2e_rate"], "unit": "tests/s", "cores": r["threads"], "kind": "port", "sample": sample,
                         "cond_phase_value": r["cond_rate"], "secs": r["secs"],
                         "note": "C++/OpenMP restatement of FlashWeave.jl semantics (oracle/), not the Julia package: Julia is not installed"},
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    p, n = a.p, a.n
    # ---- setup (untimed): synthetic table in pinned host memory ---------------------------------------------------
    # N = 1: rank 0's pinned buffer.  N > 1: one copy in POSIX shared memory (what FlashWeave's SharedArray gives its local
    # workers, src/learning.jl:553-560), page-locked by every rank, so each rank uploads its 1/N slice over its own PCIe link.
    host_x = None
    gen_s = 0.0
    shm_path = None
    split_h2d = world > 1 and p % world == 0
    if split_h2d:
        shm_path = "/dev/shm/fw_bench_table_%s.bin" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            t0 = time.time()
            mm = np.memmap(shm_path, mode="w+", shape=(p, n), dtype=np.float32)
            mm[:] = synth.clique(p, n, B=a.B, seed=synth.BASE_SEED + 3)
            mm.flush(); del mm
            gen_s = time.time() - t0
        dist.barrier()
        mm = np.memmap(shm_path, mode="r+", shape=(p, n), dtype=np.float32)
        host_x = torch.from_numpy(mm)
        rc = torch.cuda.cudart().cudaHostRegister(host_x.data_ptr(), host_x.numel() * 4, 0)
        if int(rc) != 0:
            raise SystemExit("cudaHostRegister of the shared table failed: %s" % rc)
    elif rank == 0:
        t0 = time.time()
        host_x = torch.empty((p, n), dtype=torch.float32, pin_memory=True)
        host_x.numpy()[:] = synth.clique(p, n, B=a.B, seed=synth.BASE_SEED + 3)
        gen_s = time.time() - t0
    d_x = torch.empty((p, n), dtype=torch.float32, device="cuda")
    eng = fw.Engine(local_rank)
    d_cor = None
    if dist is not None:
        # row-sharded cor_mat: every rank computes 1/N of the tiles, two in-place all-gathers, symmetrise (parallel.sharded_cor)
        h, nb_pad = par.cor_groups((p + 127) // 128, world)
        d_cor = torch.empty((nb_pad * 128, p), dtype=torch.float32, device="cuda")
    ext = torch.cuda.ExternalStream(eng.stream)

    h2d_ms = []
    cor_wall_ms = []

    def pipeline():
        """e2e: host table -> neighbour lists of this rank's target shard, through the C ABI."""
        th = time.perf_counter()
        if dist is None:
            # one GPU: the upload is chunked and hidden behind the cor_mat GEMM inside one C-ABI call
            eng.upload_and_cor(host_x.data_ptr(), n=n, p=p)
            eng.synchronize()
            h2d_ms.append(0.0)
            cor_wall_ms.append((time.perf_counter() - th) * 1e3)
        elif split_h2d:
            r0, r1 = rank * (p // world), (rank + 1) * (p // world)
            d_x[r0:r1].copy_(host_x[r0:r1], non_blocking=True)       # H2D of this rank's slice from pinned (shared) host memory
            dist.all_gather_into_tensor(d_x, d_x[r0:r1])             # the table over NVLink, in place
        else:
            if rank == 0:
                d_x.copy_(host_x, non_blocking=True)                 # H2D from pinned host memory
            par.broadcast_table(dist, d_x, src=0)                    # the one collective: table over NVLink
        if dist is not None:
            torch.cuda.synchronize()
            h2d_ms.append((time.perf_counter() - th) * 1e3)
            eng.adopt_data_device(d_x.data_ptr(), n, p, "fz")
            tc = time.perf_counter()
            eng.adopt_cor_device_rows(d_cor.data_ptr(), p, d_cor.shape[0])
            par.sharded_cor(dist, eng, d_cor)                        # cor_mat = Float32.(cor(data)), 1/N of the tiles per rank
            eng.synchronize()
            cor_wall_ms.append((time.perf_counter() - tc) * 1e3)
        eng.pw_univar_neighbors(alpha=a.alpha, n_obs_min=20, want_host=False)
        off = np.zeros(p + 1, np.int64)
        eng._ck(eng.L.fw_pairwise_copy(eng.h, off.ctypes.data_as(fw.C.c_void_p), None, None, None))
        order = np.argsort(np.diff(off), kind="stable").astype(np.int64)      # learning.jl:97-98
        shard = par.shard_targets(order, rank, world)                          # target i -> rank i mod N
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=20, want_tpc=False, reuse_buffers=True)
        return shard, res

    # ---- e2e region ----------------------------------------------------------------------------------
    for _ in range(a.warmup):
        shard, res = pipeline()
    barrier()
    launches0 = eng.launch_count()
    e2e_t = []
    phase = {"cor_ms": [], "pairwise_ms": [], "hiton_ms": []}
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
    # the cor_mat GEMM alone (table already resident), for its tensor-core roofline
    gemm_ms = []
    if dist is None:
        for _ in range(3):
            eng.cor(want_host=False); eng.synchronize()
            gemm_ms.append(eng.last_timing()["cor_ms"])
        eng.pw_univar_neighbors(alpha=a.alpha, n_obs_min=20, want_host=False)
    tests_rank = int(res.num_tests.sum())
    h2d = (p * n * 4 // world if split_h2d else (p * n * 4 if rank == 0 else 0)) + len(shard) * 8
    d2h = int(res.off[-1]) * 24 + len(shard) * 24 + (p + 1) * 8

    # ---- device-resident region (the contract's K timed steps) -------------------------------------------
    for _ in range(a.warmup):
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=20, want_tpc=False, reuse_buffers=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches1 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(ext)
    kern_ms = []
    for _ in range(a.steps):
        res = eng.si_HITON_PC(shard, max_k=a.max_k, alpha=a.alpha, n_obs_min=20, want_tpc=False, reuse_buffers=True)
        kern_ms.append(eng.last_timing()["hiton_ms"])
    ev1.record(ext)
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / a.steps
    clocks = sampler.stop()
    launches = eng.launch_count() - launches1
    exec_k = eng.hiton_exec_by_k()
    tests_exec = res.tests_executed

    # ---- reduce over ranks: max time, sum of tests ----------------------------------------------------------
    stats = torch.tensor([dev_ms, float(np.mean(e2e_t)) * 1e3, float(tests_rank), float(tests_exec), float(launches), float(h2d), float(d2h)],
                         dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = stats, stats
    if split_h2d:
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
    alg_bytes = float(12 * exec_k[0] + 24 * exec_k[1] + 40 * exec_k[2])      # 4*C(k+2,2) B per test
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload (profiles/)
    traffic = traffic_cor = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and p == 50000 and n == 10000 and world == 1:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("hiton_fz_kernel_C4_dram_bytes_per_launch")
            traffic_cor = tj.get("cor_tc_kernel_C4_dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"kernel": "hiton_fz_kernel (si_HITON_PC conditional phase; dominant kernel of the timed `value` region)",
                "bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["which"],
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "note": "40 B of correlations per k=3 test: this kernel is instruction-issue bound, not HBM bound (see DESIGN.md 4.1)",
                "ncu": {"issue_slots_busy": 0.525, "fp64_pipe": 0.25, "alu_pipe": 0.30, "warp_instr_per_test": 10.4,
                        "source": "profiles/r01s2_hiton_fz_C4_raw.csv (one ncu --set full capture of this launch, not live)"}}
    cor_ms = float(np.mean(cor_wall_ms[-a.steps:])) if dist is not None else float(min(gemm_ms))
    roofline_cor = {"kernel": "cor_mat GEMM (N=1: fw_cor_matrix on the resident table, timed alone; N>1: row-sharded + NCCL all-gather + symmetrise, wall time of the whole step)", "bound": "tensor",
                    "achieved": 2.0 * n * p * p / world / (cor_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s per GPU",
                    "frac": 2.0 * n * p * p / world / (cor_ms * 1e-3) / 1e12 / pk["bf16_tflops"], "traffic": traffic_cor, "peak_source": pk["which"], "kernel_ms": cor_ms,
                    "note": "useful flop = 2*n*p^2 (the 3-term bf16 split and the symmetric half do not change the numerator); includes the standardise+split kernel"}

    line = {
        "metric": "CI-tests/sec (cond_tests_ref/s, HITON-PC conditional phase)", "value": value, "unit": "tests/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "targets_per_gpu": int(len(shard)), "sharding": "target i (ascending univariate degree) -> rank i mod N",
                   "l2": "inputs larger than L2 (cor_mat %.1f GB, read by gather)" % (p * p * 4 / 1e9),
                   "cond_tests_ref_per_step": tests_total, "cond_tests_executed_per_step": exec_total,
                   "pairwise_tests": p * (p - 1) // 2, "parity_semantics": "parallel=\"single\" (SURVEY.md §3.6)"},
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": sm[5].item(), "d2h_bytes_per_step": sm[6].item(),
                "ms_per_step": e2e_ms_max, "gpu_launches_per_step": e2e_launches,
                "phases_ms_rank0": dict({k: float(np.mean(v)) for k, v in phase.items()}, h2d_and_broadcast_ms=float(np.mean(h2d_ms[-a.steps:])),
                                        cor_wall_ms=float(np.mean(cor_wall_ms[-a.steps:]))),
                "pairwise_tests_per_s": p * (p - 1) / 2 / (float(np.mean(phase["pairwise_ms"])) * 1e-3)},
        "gpu_launches": int(sm[4].item()),
        "roofline": roofline, "roofline_cor_gemm": roofline_cor,
        "clocks": clocks,
        "setup": {"table_gen_s": gen_s},
    }
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(a, 1, 0, a.cpu_blocks)
        line["cpu_baseline"] = {"value": r["cond_rate"], "unit": "tests/s", "cores": r["threads"], "kind": "port",
                                "sample": "first %d blocks (%d OTUs x %d samples) of the same table, %d conditional tests; conditional phase only (as `value`)"
                                          % (a.cpu_blocks, r["p_sample"], a.n, r["tests"]),
                                "e2e_value": r["e2e_rate"], "secs": r["secs"],
                                "note": "C++/OpenMP restatement of FlashWeave.jl (oracle/), not the Julia package; omits Julia's per-test String/Vector allocations"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


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
