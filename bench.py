#!/usr/bin/env python3
"""Benchmark of the X-Mapper aligner stage on B200 (BASELINE.json: aligned reads/sec + DP GCUPS; CPU path beside it).

A step = one pass of the hot path (seed lookup -> candidate windows -> scoring/traceback) over one batch of synthetic
reads.  Workload at N=1: BASELINE.json configs[1] — 5 Mbp random reference, 1 M simulated 150 bp single-end reads with
1% substitutions + small indels.  Multi-GPU: reads are sharded, index replicated, no data-path collective; the
per-position depth planes are all-reduced with NCCL after the timed steps when --counts is given (weak scaling).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference algorithm's CPU restatement on the host cores

Prints ONE JSON line on rank 0.
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
sys.path.insert(0, ROOT)

METRIC = "aligned reads/sec"
UNIT = "reads/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1000000, help="reads per step per GPU")
    ap.add_argument("--ref-bases", type=int, default=5000000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--counts", action="store_true", help="accumulate depth planes and NCCL all-reduce them at the end")
    ap.add_argument("--cpu-sample", type=int, default=1000000, help="reads in the cpu_baseline sample (about 30 core-seconds of the oracle)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return "synthetic %.3g Mbp reference, %s simulated %d bp %s reads/step/GPU, 1%% substitutions + 0.1%%/base indels (BASELINE.json configs[%d])" % (
        a.ref_bases / 1e6, "{:,}".format(a.reads), a.read_len, "paired-end 2x" if a.paired else "single-end", 2 if a.paired else 1)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=reasons, samples=len(self.samples))


def make_inputs(a, seed_offset):
    from mapper_b200 import synth
    ref = synth.random_reference(a.ref_bases, seed=1)
    batch = synth.simulate_reads_fast(ref, a.reads, a.read_len, seed=2 + seed_offset, paired=a.paired, inner_mean=300.0, inner_sd=30.0, per_penalty=50.0)
    return ref, batch


def cpu_arm(a, ref, batch, n_reads, threads):
    """The reference's algorithm on the host cores: the oracle port (oracle/, C++ restatement of the Java aligner).
    bench.py is allowed to execute oracle/ here and only here (cpu_baseline / --impl reference)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import xm_oracle as xo
    from mapper_b200 import synth
    t0 = time.time()
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, threads=threads, dup=dict(min_copies=2, window=1000))
    db.build_through(a.read_len)
    db.detect_duplications()
    t_index = time.time() - t0
    n_mates = 2 if a.paired else 1
    ns = n_reads * n_mates
    sub = dict(packed=batch["packed"][:batch["seq_word_off"][ns]], seq_word_off=batch["seq_word_off"][:ns + 1], seq_len=batch["seq_len"][:ns],
               n_seqs=batch["n_seqs"][:n_reads], expected_inner=batch["expected_inner"][:n_reads], per_penalty=batch["per_penalty"][:n_reads])
    sub = {k: np.ascontiguousarray(v) for k, v in sub.items()}

    def run():
        t = time.time()
        r = db.align_batch(synth.DEFAULT_PARAMS, sub, threads=threads)
        return time.time() - t, r
    return run, t_index


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        # the reference arm: rank 0 alone runs and prints; other ranks exit 0 without work
        if rank != 0:
            return 0
        ref, batch = make_inputs(a, 0)
        n_sample = min(a.reads, a.cpu_sample)
        run, t_index = cpu_arm(a, ref, batch, n_sample, cores)
        for _ in range(max(0, min(a.warmup, 1))):
            run()
        times = []
        for _ in range(a.steps):
            dt, r = run()
            times.append(dt)
        total = sum(times)
        value = n_sample * a.steps / total
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=1000.0 * total / a.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=dict(workload=workload_name(a), sample="%d reads per step (bounded sample of the workload)" % n_sample),
                    cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                      sample="%d reads/step x %d steps; C++ restatement of mathjeff/Mapper @ ae7f346a (the Java reference cannot be built here: no JVM), %d threads" % (n_sample, a.steps, cores)),
                    e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), index_build_s=t_index)
        _emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from mapper_b200 import capi, synth
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the product has no CPU path", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ref, batch = make_inputs(a, rank)
    g = capi.XMapper(synth.DEFAULT_PARAMS, device=local_rank)
    t0 = time.time()
    g.set_reference([synth.pack_contig(s) for _, s in ref], [len(s) for _, s in ref])
    g.build_index(a.read_len)  # on the device
    g.build_duplications(-1, -1, 2, 1000)
    t_index = time.time() - t0
    if a.counts:
        g.counts_enable(0.1)
        if world > 1:
            # xm_comm_init / xm_counts_reduce: torch.distributed only carries the 128-byte NCCL id from rank 0 to the others.
            # The first reduce runs here, on planes that are still all zero, so that NCCL's channel setup is not in the timed one.
            uid = torch.tensor(list(g.comm_unique_id() if rank == 0 else bytes(128)), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, 0)
            g.comm_init(world, rank, bytes(uid.cpu().tolist()))
            g.counts_reduce()
    mi, mb = g.index_info()
    index_bytes = 0
    for n in range(mb + 1):
        t = g.get_index_length(n)
        index_bytes += 8 * t["capacity"] + 4 * len(t["positions"])

    # pinned host copies (e2e leg) and device-resident copies (kernel leg)
    keys = ["packed", "seq_word_off", "seq_len", "n_seqs", "expected_inner", "per_penalty"]
    pinned, dev = {}, {}
    for k in keys:
        arr = batch[k]
        view = arr.view(np.int16) if arr.dtype == np.uint16 else arr
        t = torch.from_numpy(view.copy()).pin_memory()
        pinned[k] = t.numpy().view(arr.dtype)
        dev[k] = t.cuda(non_blocking=False)
    host_batch = {k: pinned[k] for k in keys}
    nq = len(batch["n_seqs"])
    n_words = int(batch["seq_word_off"][-1])

    def step_device():
        return g.align_batch_device(nq, dev["packed"].data_ptr(), n_words, dev["seq_word_off"].data_ptr(), dev["seq_len"].data_ptr(), dev["n_seqs"].data_ptr(),
                                    dev["expected_inner"].data_ptr(), dev["per_penalty"].data_ptr(), a.read_len)

    def step_host():
        return g.align_batch(host_batch, copy=False)  # results stay in the library's pinned slab (zero-copy views)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        r = step_device()
    bad = int((r["q_status"] != 0).sum())
    # ---- timed: kernel leg (inputs resident in HBM), device time from CUDA events inside the library ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    dev_ns = 0
    align_ns = 0
    easy_ns_sum = full_ns_sum = 0
    launches = 0
    stats = None
    for _ in range(a.steps):
        r = step_device()
        dev_ns += int(r["stats"][capi.STAT["kernel_ns"]])
        align_ns += int(r["stats"][capi.STAT["align_kernel_ns"]])
        easy_ns_sum += int(r["stats"][capi.STAT["easy_ns"]])
        full_ns_sum += int(r["stats"][capi.STAT["tier0_ns"]])
        launches += int(r["stats"][capi.STAT["launches"]])
        stats = r["stats"]
    barrier()
    wall_dev = time.time() - t0
    # ---- timed: end to end through the C ABI with host buffers (H2D + kernels + D2H + result assembly) ----
    keep = [step_host() for _ in range(2)]  # untimed warm-up of the host path: staging buffers and BOTH pinned result slabs (a caller holds one result while the next batch runs)
    del keep
    barrier()
    t0 = time.time()
    h2d = d2h = 0
    for _ in range(a.steps):
        r2 = step_host()
        bad_e2e = int(np.count_nonzero(r2["q_status"]))  # read the step's result on the host
        h2d = int(r2["stats"][capi.STAT["h2d_bytes"]])
        d2h = int(r2["stats"][capi.STAT["d2h_bytes"]])
    barrier()
    wall_e2e = time.time() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    allreduce_ms = None
    if a.counts:
        ptr, n_int = g.counts_device_ptr()
        if world > 1:
            barrier()
            t0 = time.time()
            g.counts_reduce()  # the library's own NCCL all-reduce of the planes (int32 sum: exact, order-free), once: the planes stay the sum over ranks
            torch.cuda.synchronize()
            allreduce_ms = 1000.0 * (time.time() - t0)

    # max over ranks
    tvals = torch.tensor([dev_ns / 1e9, wall_dev, wall_e2e, align_ns / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tvals, op=dist.ReduceOp.MAX)
    t_dev, t_wall_dev, t_e2e, t_align = [float(x) for x in tvals.tolist()]
    total_reads = nq * a.steps * world
    value = total_reads / t_dev
    e2e_value = total_reads / t_e2e

    if rank == 0:
        aligned = int((np.diff(r["comp_choice_off"])[r["q_comp_off"][:-1]] > 0).sum())
        S = capi.STAT
        # Two align kernels per step: the first pass over every read (xm_align_kernel<true>) and the full aligner over the reads
        # it handed on (xm_align_kernel<false>).  Algorithmic bytes per launch: 8 B per bucket probe, 4 B position +
        # 19 B flank window per hit, ceil(L/2) B of packed read per query, ceil(L/2) B of reference per ungapped score.
        half = (a.read_len + 1) // 2
        n_mates = 2 if a.paired else 1
        probes, hits, straight = int(stats[S["probes"]]), int(stats[S["hits"]]), int(stats[S["straight"]])
        e_probes, e_hits, e_straight = int(stats[S["easy_probes"]]), int(stats[S["easy_hits"]]), int(stats[S["easy_straight"]])
        n_easy_in, n_easy_done = int(stats[S["easy"]]), int(stats[S["easy_done"]])
        n_full_in = int(stats[S["tier0"]])
        bytes_easy = 8 * e_probes + 23 * e_hits + half * n_mates * n_easy_in + half * e_straight
        bytes_full = 8 * (probes - e_probes) + 23 * (hits - e_hits) + half * n_mates * n_full_in + half * (straight - e_straight)
        ms_easy = easy_ns_sum / 1e6 / a.steps  # average launch duration over the timed steps (CUDA events on the library stream)
        ms_full = full_ns_sum / 1e6 / a.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        # measured DRAM traffic of the two kernels (one ncu capture of this command, profiles/): only quoted for the workload it was taken on
        traffic = {}
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r1o_traffic.json")))
            if t["workload"] == dict(reads=a.reads, read_len=a.read_len, paired=bool(a.paired)) and a.ref_bases == 5000000:
                traffic = t
        except Exception:
            pass
        dom_full = ms_full >= ms_easy
        dom_bytes, dom_ms = (bytes_full, ms_full) if dom_full else (bytes_easy, ms_easy)
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else None
        cells = int(stats[S["path_cells"]])
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=1000.0 * t_dev / a.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=workload_name(a), reads_per_step_per_gpu=nq, parallelism="reads sharded x%d, index replicated" % world,
                                l2="inputs larger than L2: index %.0f MB + packed reads %.0f MB + per-warp workspaces (GBs) per step vs 126 MB L2" % (index_bytes / 1e6, batch["packed"].nbytes / 1e6),
                                timing="value: CUDA-event device time of all kernels of a step (library stream), max over ranks; e2e: wall clock around xm_align_batch with pinned host buffers"),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=1000.0 * t_e2e / a.steps),
                    gpu_launches=launches,
                    roofline=dict(bound="hbm", kernel="xm_align_kernel<false> (full aligner over the reads the first pass handed on)" if dom_full else "xm_align_kernel<true> (first pass)",
                                  achieved=achieved, peak=peak, unit="GB/s", frac=(achieved / peak) if achieved else None,
                                  traffic=(lambda k: (k["dram_bytes_read"] + k["dram_bytes_write"]) if k else None)(traffic.get("full" if dom_full else "first_pass")),
                                  traffic_source=traffic.get("source"),
                                  algorithmic_bytes_per_launch=dom_bytes, kernel_ms_per_launch=dom_ms, peak_source=peak_src,
                                  other_kernel=dict(name="xm_align_kernel<true> (first pass)" if dom_full else "xm_align_kernel<false>",
                                                    algorithmic_bytes_per_launch=bytes_easy if dom_full else bytes_full, kernel_ms_per_launch=ms_easy if dom_full else ms_full),
                                  note="not bandwidth-bound: warp-uniform scalar code bound by instruction fetch (ncu: stall_no_instruction 74% of stall cycles, issue slots 19% busy, 0.75 warp-instructions per SM-cycle; profiles/r1k_*). The DRAM traffic is per-warp workspace (lattices, pyramids, stack frames), ~2.5% of HBM bandwidth"),
                    gcups=dict(value=(cells * 1.0 / (t_dev / a.steps) / 1e9), unit="GCUPS", cells_per_step=cells, path_aligner_calls=int(stats[S["path_calls"]]),
                               cells_explored_per_step=int(stats[S["path_steps"]]),
                               note="cells = A x B of every PathAligner lattice of a step (SURVEY.md §8d) / device time of the whole step"),
                    clocks=sampler.summary(),
                    aligned_fraction=aligned / nq, failed_queries=bad, index_build_s=t_index,
                    first_pass=dict(queries=n_easy_in, completed=n_easy_done, ms=ms_easy),
                    full_pass=dict(tiers=[int(stats[S["tier0"]]), int(stats[S["tier1"]]), int(stats[S["tier2"]])],
                                   ms=[float(stats[S["tier%d_ns" % t]]) / 1e6 for t in range(3)]),
                    wall_ms_per_step_device_api=1000.0 * t_wall_dev / a.steps, allreduce_ms=allreduce_ms)
        if not a.no_cpu_baseline:
            n_sample = min(a.reads, a.cpu_sample)
            run, _ = cpu_arm(a, ref, batch, n_sample, cores)
            dt, rc = run()
            # the oracle's results for the sample double as a parity check of this run's device results (bit for bit, doubles included)
            line["parity"] = dict(reads=n_sample, identical_to_oracle=bool(same_prefix(rc, r, n_sample)),
                                  note="all eleven result arrays of the first `reads` queries of rank 0 compared with the CPU oracle")
            line["cpu_baseline"] = dict(value=n_sample / dt, unit=UNIT, cores=cores, kind="port",
                                        sample="%d reads of the same workload, %.1f s on %d threads; C++ restatement of mathjeff/Mapper @ ae7f346a (no JVM available)" % (n_sample, dt, cores))
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    g.close()
    return 0


def same_prefix(want, got, n):
    """True if the result arrays of the first n queries of `got` equal `want` (a result for exactly those n queries) bit for bit."""
    c = int(got["q_comp_off"][n]); ch = int(got["comp_choice_off"][c]); sa = int(got["choice_sa_off"][ch]); bl = int(got["sa_block_off"][sa])
    sizes = dict(q_comp_off=n + 1, comp_choice_off=c + 1, choice_sa_off=ch + 1, sa_block_off=sa + 1, choice_f64=4 * ch, sa_f64=2 * sa,
                 choice_inner=ch, sa_contig=sa, blocks=4 * bl, q_status=n, sa_reversed=sa)
    for k, m in sizes.items():
        x, y = np.ascontiguousarray(want[k]), np.ascontiguousarray(got[k][:m])
        if x.dtype != y.dtype:
            x = x.astype(y.dtype)
        if len(x) != m or x.tobytes() != y.tobytes():
            return False
    return True


def _emit(line):
    """The contract is ONE JSON line on stdout: libraries (NCCL's version banner, torchrun) write to fd 1 as well, so fd 1 is
    pointed at stderr for the whole run and the line goes to the saved descriptor."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

if __name__ == "__main__":
    sys.exit(main())
