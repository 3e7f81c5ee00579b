#!/usr/bin/env python
"""bench.py — batched 64-bit range-proof verifications/sec (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE config 2, a batch of 1024 independent 64-bit RangeProofs (m = 1,
672 B proof + 32 B commitment each).  One *step* = one pass of the verification hot path over one batch:
transcript replay + verification scalars + 17 408 Ristretto decompressions + one 17 538-term
random-linear-combination MSM + identity check, all on the GPU.

  value  : whole-job proofs/s with the input batches already resident in HBM (device-pointer entry point)
  e2e    : the same metric through the host-buffer C-ABI call (pinned host buffers; the H2D copy of the
           proofs and the D2H copy of the verdicts are inside the timed region)
  --impl reference : the reference's CPU path (the oracle restatement; the Rust crate cannot be built
           in this image) on all host cores, same metric and config.

Steps are pipelined over `--streams` CUDA streams (one bp_ctx each); timing is CUDA events on the launching
streams bracketed by device synchronisation (and a barrier under torchrun), max over ranks.  Inputs rotate
through a pool of distinct batches larger than L2 (config.l2 says so).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

# one hardware work queue per stream (the default of 8 makes >8 streams share queues and serialise)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_BITS, M_PARTIES, BATCH = 64, 1, 1024
LABEL = b"AggregateRangeProofBenchmark"            # benches/range_proof.rs:34
L2_BYTES = 126 * 1024 * 1024


def effective_cores():
    """host cores this process may actually use: min(affinity, cgroup cpu.max quota)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` capture
# profiles/r1_final_ncu_full.md; the scratch arrays (sorted ids, buckets, contrib) make it larger than the algorithmic bytes
NCU_TRAFFIC_BYTES = {"k_msm_accumulate": 3.91e6, "k_msm_reduce": 3.37e6, "k_msm_combine": 0.025e6, "k_rp_transcript": 0.77e6, "k_rp_head": 1.29e6,
                     "k_rp_scalars": 2.93e6, "k_rp_decompress": 0.83e6, "k_rp_static_reduce": 4.27e6}
# sm__pipe_fmaheavy_cycles_active x elapsed cycles summed over the kernels of one config-2 batch, per SM, same capture
NCU_FMAHEAVY_BUSY_CYCLES_PER_SM = 151400


def make_workload(count, rank):
    """Synthetic input: `count` valid (64,1) proofs over uniform 64-bit values and uniform blindings.
    The proofs are produced by the CPU oracle's prover (test infrastructure used as a data generator only;
    nothing on the measured path touches it)."""
    import random
    from oracle_binding import Oracle, L_ORDER
    orc = Oracle()
    og = orc.gens(N_BITS, M_PARTIES)
    rnd = random.Random(1000 + rank)
    values = [rnd.randrange(1 << N_BITS) for _ in range(count * M_PARTIES)]
    blind = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(count * M_PARTIES))
    seeds = b"".join((rank * count + i).to_bytes(8, "little") + bytes(24) for i in range(count))
    proofs, Vs = orc.prove_many(og, orc.transcript(LABEL), values, blind, N_BITS, M_PARTIES, seeds, nthreads=effective_cores())
    return orc, og, proofs, Vs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe): one `nvidia-smi -lms 50`
    process, every line stamped on arrival; stop() summarises the samples that fell inside mark_begin()..mark_end() windows."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.windows, self.proc = index, [], [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 8:
                    self.samples.append((time.perf_counter(), f))
        except Exception:
            pass

    def mark_begin(self):
        self.windows.append([time.perf_counter(), None])

    def mark_end(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=3)
        inside = [f for t, f in self.samples if any(a <= t <= (b or 1e30) for a, b in self.windows)] or [f for _, f in self.samples]
        sm = sorted(int(float(f[1])) for f in inside)
        reasons = set()
        for f in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(inside[0][2])) if inside else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max((float(f[3]) for f in inside), default=None)}


def run_reference(args, rank, world):
    """The reference's own CPU path for the metric: per-proof RangeProof::verify_multiple on every host core
    (oracle restatement, kind "port").  Rank 0 only."""
    if rank != 0:
        return
    cores = effective_cores()
    orc, og, proofs, Vs = make_workload(BATCH, 0)
    plen = len(proofs) // BATCH
    t = orc.transcript(LABEL)
    probe = min(BATCH, 8 * cores)
    t0 = time.perf_counter(); st = orc.verify_many(og, t, proofs[:probe * plen], plen, Vs[:probe * 32 * M_PARTIES], N_BITS, M_PARTIES, probe, nthreads=cores); dt = time.perf_counter() - t0
    assert not any(st)
    rate = probe / dt
    budget_s = 90.0
    sample = int(max(cores, min(BATCH, rate * budget_s / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        orc.verify_many(og, t, proofs[:sample * plen], plen, Vs[:sample * 32 * M_PARTIES], N_BITS, M_PARTIES, sample, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = orc.verify_many(og, t, proofs[:sample * plen], plen, Vs[:sample * 32 * M_PARTIES], N_BITS, M_PARTIES, sample, nthreads=cores)
    dt = time.perf_counter() - t0
    assert not any(st)
    value = sample * args.steps / dt
    cpu = {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
           "sample": f"{sample} of the {BATCH} (64,{M_PARTIES}) proofs per step, per-proof verify_multiple ({2 * N_BITS * M_PARTIES + 2 * (N_BITS * M_PARTIES).bit_length() - 2 + M_PARTIES + 6} terms), one proof per task on {cores} threads"}
    print(json.dumps({"impl": "reference", "metric": "64-bit rangeproof verifications/sec (batched)", "value": value, "unit": "proofs/s", "n_gpus": 0,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u64 (51-bit limbs, CPU)", "data": "synthetic",
                      "config": {"workload": f"batched verify of {BATCH}x 64-bit RangeProofs (m={M_PARTIES}) per GPU", "n": N_BITS, "m": M_PARTIES, "batch": BATCH,
                                 "reference_sample": f"{sample} proofs per step on {cores} host threads"},
                      "cpu_baseline": cpu, "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    global M_PARTIES, BATCH
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--streams", type=int, default=24)
    ap.add_argument("--threads", type=int, default=0, help="host threads issuing steps (each drives streams/threads contexts); 0 = min(3, host cores per rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--m", type=int, default=M_PARTIES, help="parties per proof (BASELINE config 3: --m 16 --batch 256); the bench line is the default")
    ap.add_argument("--batch", type=int, default=BATCH, help="proofs per verified batch")
    args = ap.parse_args()
    M_PARTIES, BATCH = args.m, args.batch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import bulletproofs_b200 as bp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback on the MSM path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    orc, og, proofs, Vs = make_workload(BATCH, rank)
    plen = len(proofs) // BATCH
    S = max(1, args.streams)
    streams = [torch.cuda.Stream(device=local) for _ in range(S)]
    ctxs = [bp.Context(local, stream=s.cuda_stream) for s in streams]

    # generator table: derived once on rank 0, one NCCL broadcast over NVLink, imported by every context's table
    gens0 = bp.Gens(ctxs[0], N_BITS, M_PARTIES, empty=(rank != 0))
    _, table_bytes = gens0.device_table()
    table = torch.empty(table_bytes, dtype=torch.uint8, device="cuda")
    if rank == 0:
        gens0.table_export(table.data_ptr())
    from bulletproofs_b200.dist import broadcast_table
    broadcast_table(table, src=0)                  # the path's only collective: one NCCL broadcast over NVLink
    torch.cuda.synchronize()
    gens = [gens0]
    if rank != 0:
        gens0.table_import(table.data_ptr())
    for c in ctxs[1:]:
        g = bp.Gens(c, N_BITS, M_PARTIES, empty=True); g.table_import(table.data_ptr()); gens.append(g)
    assert gens[-1].G(0, 5) == orc.gens_get(og, 0, 0, 5)

    transcript = bp.Transcript(LABEL)
    ver = [bp.BatchVerifier(ctxs[i], gens[i], transcript, N_BITS, M_PARTIES, BATCH) for i in range(S)]

    # input pool larger than L2: rotations of the proof order (distinct memory, same proofs)
    batch_bytes = BATCH * (plen + 32 * M_PARTIES)
    P = L2_BYTES // batch_bytes + 8
    pr = np.frombuffer(proofs, dtype=np.uint8).reshape(BATCH, plen); vs = np.frombuffer(Vs, dtype=np.uint8).reshape(BATCH, 32 * M_PARTIES)
    h_proofs = torch.empty((P, BATCH, plen), dtype=torch.uint8).pin_memory(); h_vs = torch.empty((P, BATCH, 32 * M_PARTIES), dtype=torch.uint8).pin_memory()
    for i in range(P):
        h_proofs[i] = torch.from_numpy(np.roll(pr, i * 5, axis=0).copy()); h_vs[i] = torch.from_numpy(np.roll(vs, i * 5, axis=0).copy())
    d_proofs = h_proofs.cuda(); d_vs = h_vs.cuda()
    d_verdicts = torch.zeros((S, BATCH), dtype=torch.int32, device="cuda")
    h_ok = torch.ones(S, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    # correctness gate before timing: GPU verdicts == oracle verdicts on one good and one damaged batch
    bad = bytearray(proofs); bad[7 * plen + 300] ^= 1
    got = bp.verify_batch(ctxs[0], gens[0], transcript, bytes(bad), Vs, N_BITS, M_PARTIES, BATCH)
    assert [i for i, v in enumerate(got) if v] == [7], "GPU verdicts differ from the expected ones"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    from concurrent.futures import ThreadPoolExecutor
    NT = args.threads if args.threads > 0 else min(3, max(1, effective_cores() // world))     # 3 issuing threads: e2e +5 % over one (profiles/r1_timeline.md)
    NT = max(1, min(NT, S))
    pool = ThreadPoolExecutor(NT) if NT > 1 else None

    def run_steps(step_fn, first, n):
        """issue steps first..first+n-1; with several host threads, thread t issues the steps whose context index k = i % S has k % NT == t"""
        if pool is None:
            for i in range(first, first + n):
                step_fn(i)
            return
        def worker(t):
            for i in range(first, first + n):
                if (i % S) % NT == t:
                    step_fn(i)
        list(pool.map(worker, range(NT)))

    def timed(step_fn, drain_fn, steps, warmup):
        run_steps(step_fn, 0, warmup)
        drain_fn()
        barrier()
        l0 = sum(c.launches for c in ctxs)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event() for _ in range(S)]
        start.record(streams[0])
        for s in streams[1:]:
            s.wait_event(start)
        t0 = time.perf_counter()
        run_steps(step_fn, warmup, steps)
        host_issue = time.perf_counter() - t0
        drain_fn()
        for s, e in zip(streams, ends):
            e.record(s); streams[0].wait_event(e)
        end.record(streams[0])
        barrier()
        wall = time.perf_counter() - t0
        ms = start.elapsed_time(end)
        if world > 1:
            tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, wall, sum(c.launches for c in ctxs) - l0, host_issue

    # ---- value: inputs resident in HBM
    def step_dev(i):
        k, j = i % S, i % P
        ver[k].run_device(d_proofs[j].data_ptr(), d_vs[j].data_ptr(), d_verdicts[k].data_ptr(), h_ok[k:].data_ptr())

    clk = ClockSampler(local); clk.start()
    time.sleep(0.3)                                   # let the sampler come up before the first timed region
    clk.mark_begin()
    ms_dev, _, launches, host_issue_dev = timed(step_dev, lambda: None, args.steps, args.warmup)
    clk.mark_end()
    assert int(h_ok.min()) == 1 and int(d_verdicts.abs().max()) == 0, "a timed batch did not verify"
    value = world * BATCH * args.steps / (ms_dev * 1e-3)

    # ---- e2e: host buffers through the public C-ABI call (H2D + kernels + D2H per step)
    def step_e2e(i):
        k, j = i % S, i % P
        if ver[k].busy:
            assert not any(ver[k].finish())
        ver[k].begin(h_proofs[j].data_ptr(), h_vs[j].data_ptr())

    def drain_e2e():
        for v in ver:
            if v.busy:
                assert not any(v.finish())

    clk.mark_begin()
    ms_e2e, wall_e2e, _, _ = timed(step_e2e, drain_e2e, args.steps, args.warmup)
    clk.mark_end()
    clocks = clk.stop()
    e2e_value = world * BATCH * args.steps / (max(ms_e2e * 1e-3, wall_e2e))
    h2d = BATCH * (plen + 32 * M_PARTIES) + bp.TRANSCRIPT_BYTES + 32 + 8 + 4
    d2h = 4 * BATCH + 4

    # ---- per-kernel durations (CUDA events around every launch, single stream) -> roofline block
    ctxs[0].prof_enable(True)
    psteps = 10
    for i in range(psteps):
        ver[0].run_device(d_proofs[i % P].data_ptr(), d_vs[i % P].data_ptr(), d_verdicts[0].data_ptr(), None)
    prof = ctxs[0].prof_report(); ctxs[0].prof_enable(False)
    total_ms = sum(v[0] for v in prof.values())
    # dominant kernel = the Pippenger bucket accumulation: the largest share of executed warp instructions of a step
    # (27 % in the committed ncu capture profiles/r1_final_ncu_full.md, next to the per-proof decompressions' 34 %; it is the MSM kernel north_star names).  The single-warp k_msm_combine and the 32-warp
    # k_rp_transcript have longer durations when a batch runs alone, but they are latency chains that overlap with the
    # other batches in flight and use <1 % of the issue slots.
    dom = "k_msm_accumulate" if "k_msm_accumulate" in prof else max(prof, key=lambda k: prof[k][0])
    dom_ms = prof[dom][0] / prof[dom][1]
    k_lg = (N_BITS * M_PARTIES).bit_length() - 1
    T_terms = 2 + 2 * N_BITS * M_PARTIES + BATCH * (4 + 2 * k_lg + M_PARTIES)
    alg_bytes_step = BATCH * (32 * (9 + 2 * k_lg) + 32 * M_PARTIES + 1) + 32 * (2 * N_BITS * M_PARTIES + 2)     # SURVEY.md §8(d), per verified batch
    alg_bytes = 64 * T_terms + 32 if dom.startswith("k_msm") else alg_bytes_step                                     # SURVEY.md §8(d), MSM of T terms
    # the roofline that actually binds: IMAD.WIDE.U32 issue (1 per 4 cycles per SM sub-partition, measured 9.25e12 thread-level
    # wide multiply-adds/s on this pool's B200, profiles/r1_imad_peak.md).  Wide multiplies per verified batch, counted from the
    # kernels' formulas: fe_mul 72, fe_sq 44, Montgomery product 192.
    W_win = (255 + 10) // 11; pts = BATCH * (4 + 2 * k_lg + M_PARTIES); Nv = N_BITS * M_PARTIES
    wide = (pts * (257 * 44 + 29 * 72)                      # decompress: one 2^252-3 exponentiation + decode + Niels form per point
            + T_terms * W_win * 7 * 72                       # bucket accumulation: one mixed addition per term and window
            + W_win * (59 * 64 + 1) * 9 * 72                 # bucket reduction: 59 additions per thread, 64 threads per window
            + BATCH * (Nv * 3 + (4 + 2 * k_lg + M_PARTIES) * 3 + 170) * 192)   # scalar assembly (3 products per generator index) + per-proof head
    INT_PEAK = 9.25e12
    int_pipe = {"unit": "wide multiply-adds/s (IMAD.WIDE.U32, thread level)", "per_step": wide, "achieved": wide * (value / world / BATCH), "peak": INT_PEAK,
                "frac": wide * (value / world / BATCH) / INT_PEAK, "peak_source": "measured, benchmarks/imad_microbench.cu (profiles/r1_imad_peak.md)"}
    if (M_PARTIES, BATCH) == (1, 1024) and clocks.get("sm_mhz"):
        # the same fraction from the hardware counter of the committed ncu capture: busy cycles of the FMA-heavy pipe per batch / cycles per batch
        int_pipe["ncu_fmaheavy"] = {"busy_cycles_per_sm_per_step": NCU_FMAHEAVY_BUSY_CYCLES_PER_SM,
                                    "frac": NCU_FMAHEAVY_BUSY_CYCLES_PER_SM / (ms_dev / args.steps * 1e-3 * clocks["sm_mhz"] * 1e6), "source": "profiles/r1_final_ncu_full.md"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES.get(dom),
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes, "units_per_launch": f"{T_terms} MSM terms x 64 B + 32 B" if dom.startswith("k_msm") else f"{BATCH} proofs",
                "kernel_ms": dom_ms, "kernel_share_of_step": prof[dom][0] / total_ms,
                "whole_step": {"algorithmic_bytes": alg_bytes_step, "achieved_GBps_at_value": alg_bytes_step * (value / world / BATCH) / 1e9},
                "int_pipe": int_pipe,
                "note": "integer-pipe bound path (IMAD.WIDE field multiplies): the HBM fraction is reported as BASELINE.json asks; issue-slot / FMA-pipe utilisation per kernel is in profiles/",
                "per_kernel_ms_per_step": {k: round(v[0] / psteps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}

    out = {"metric": "64-bit rangeproof verifications/sec (batched)", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (mod 2^255-19) / u32x8 (mod l)",
           "data": "synthetic (oracle-proved valid proofs over uniform 64-bit values)",
           "config": {"workload": f"batched verify of {BATCH}x 64-bit RangeProofs (m={M_PARTIES}) per GPU", "n": N_BITS, "m": M_PARTIES, "batch": BATCH, "streams": S, "host_threads": NT, "host_issue_ms_per_step": round(1e3 * host_issue_dev / args.steps, 4),
                      "l2": f"inputs larger than L2: pool of {P} distinct input batches ({P * batch_bytes >> 20} MiB) cycled", "parallelism": f"independent batches per GPU x{world}"},
           "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": max(ms_e2e, wall_e2e * 1e3) / args.steps},
           "gpu_launches": launches, "clocks": clocks, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = effective_cores()
        t = orc.transcript(LABEL)
        done, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            st = orc.verify_many(og, t, proofs, plen, Vs, N_BITS, M_PARTIES, BATCH, nthreads=cores); done += BATCH
            assert not any(st)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": done / dt, "unit": "proofs/s", "cores": cores, "kind": "port",
                               "sample": f"{done} per-proof verify_multiple calls (the {BATCH}-proof batch x{done // BATCH}) on {cores} threads, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
