"""Builds the round-2 profile summaries under profiles/ from the files a GPU run left in gpurun_out/ (scratch):
   python profiles/make_r2.py   ->  r2_final_*.json copies, r2_final_launches.{csv,md}, r2_ncu_full.md, r2_ncu_traffic.json, r2_timeline.md"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
sys.path.insert(0, P)
from summarize_launches import summarize

for name in ("r2_final_bench.json", "r2_final_bench_config3.json", "r2_final_reference.json", "r2_final_timeline.json", "r2_final_reject.json", "r2_final_secondary.json", "r2_final_launches.csv"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))

# ---- launch list
open(os.path.join(P, "r2_final_launches.md"), "w").write(
    "# Round 2, final engine — ncu launch list (one stream, one 8-batch launch group per step)\n\n"
    "Command (gpurun, 1x B200): `ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 --group 8 --no-cpu-baseline`\n"
    "(raw list: `r2_final_launches.csv`).  Cold-cache and serialised: only the shares are comparable with `bench.py`'s `roofline.kernel_share_of_step`.\n"
    "The launches are the nodes of the captured CUDA graph (14 kernels per group of 8 batches = 1.75 launches per 1024-proof batch; round 1: 17 per batch).\n\n"
    + summarize(os.path.join(P, "r2_final_launches.csv")) + "\n")

# ---- ncu --set full
raw = subprocess.run(["ncu", "-i", os.path.join(G, "r2_final_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; col = {n: i for i, n in enumerate(h)}
table = subprocess.run([sys.executable, os.path.join(P, "summarize_ncu.py")], input=raw, capture_output=True, text=True).stdout
seen, busy, traffic = set(), {}, {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    if name in seen:
        continue
    seen.add(name)
    cyc = float(r[col["sm__cycles_elapsed.max"]]); pct = float(r[col["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]])
    busy[name] = cyc * pct / 100.0
    def tobytes(v, u):
        f = float(v.replace(",", "")); return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[name] = tobytes(r[col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_read.sum"]]) + tobytes(r[col["dram__bytes_write.sum"]], rows[1][col["dram__bytes_write.sum"]])
total = sum(busy.values())
bench = json.load(open(os.path.join(P, "r2_final_bench.json")))
mhz = bench["clocks"]["sm_mhz"]; batches_per_s = bench["value"] / 1024.0
cyc_per_batch = mhz * 1e6 / batches_per_s
frac = (total / 8.0) / cyc_per_batch
json.dump({"source": "profiles/r2_ncu_full.md (ncu --set full, one 8-batch group)", "group": 8, "fmaheavy_busy_cycles_per_sm_per_group": round(total),
           "dram_bytes_per_launch": {k: round(v) for k, v in traffic.items()}}, open(os.path.join(P, "r2_ncu_traffic.json"), "w"), indent=1)
md = ["# Round 2, final engine — ncu --set full of one launch group (8 batches of 1024 proofs)\n",
      "Command (gpurun, 1x B200): `ncu --set full --clock-control none --import-source on -k regex:\"k_rp_|k_msm_\" -s 28 -c 14 -o gpurun_out/r2_final_full python bench.py --steps 1 --warmup 0 --streams 1 --group 8 --no-cpu-baseline`;",
      "table made with `ncu -i ... --page raw --csv | python profiles/summarize_ncu.py` (`python profiles/make_r2.py` regenerates this file).  `bench.py` JSON of the same build: `r2_final_bench.json`",
      f"(value {bench['value'] / 1e6:.2f} M proofs/s, e2e {bench['e2e']['value'] / 1e6:.2f} M proofs/s, SM clock {mhz} MHz, throttle reasons {bench['clocks']['reasons']}).\n",
      "No kernel is HBM bound: the largest DRAM traffic per launch is k_msm_accumulate's (below) at 1.4 % of the measured 6.59 TB/s.  `roofline.traffic` in the bench line =",
      f"dram read + write of k_msm_accumulate = {traffic.get('k_msm_accumulate', 0) / 1e6:.1f} MB per launch against {8 * (64 * 17538 + 32) / 1e6:.2f} MB algorithmic (8 MSMs x 17 538 terms x 64 B): the difference is the 96-byte",
      "affine-Niels form of every point (a point sits in 23 windows but stays in L2 after its first read) plus the sorted id lists and the 4-byte point-index map.\n", table,
      "\n## The binding resource: the FMA-heavy (integer multiply) pipe\n",
      "`sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed` x `sm__cycles_elapsed.max`, per SM, one group of 8 batches (this capture):\n",
      "| kernel | heavy-pipe busy cycles / SM / group | pipe busy while it runs alone |", "|---|---|---|"]
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    if name in busy and busy[name] is not None:
        md.append(f"| {name} | {busy[name]:,.0f} | {float(r[col['sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']]):.1f} % |"); busy[name] = None
md += [f"| **total** | **{total:,.0f}** = {total / 8:,.0f} per batch = {total / 8 / mhz:.1f} us at {mhz} MHz | |", "",
       f"`bench.py` runs a batch every {1e6 / batches_per_s:.1f} us = {cyc_per_batch:,.0f} cycles, so the multiply pipe is busy **{100 * frac:.0f} %** of the time over the whole step",
       "(round 1: 55 % at 7.25 M proofs/s claimed, 27 % at the driver-observed rate).  Alone, with the grids of an 8-batch group, the two wide kernels hold it 73-77 % busy; the remainder of a group's",
       "pipe time is the latency-bound stages (bucket reduction, head, scalar sums, transcript) whose warps sit beside the wide kernels' (`r2_timeline.md`)."]
open(os.path.join(P, "r2_ncu_full.md"), "w").write("\n".join(md) + "\n")

# ---- timeline
tl = json.load(open(os.path.join(P, "r2_final_timeline.json")))
md = ["# Round 2 — multi-stream timeline of the final engine (no nsys in the image)\n",
      "`python benchmarks/timeline.py 8 6 8`: CUDA events around every launch of 8 contexts (direct launches, one 8-batch group per call), all timestamps against one reference event,",
      "steady-state window = middle half of the run.  The events cost ~10 % of throughput.\n",
      f"{tl['streams']} groups in flight, {tl['batches_per_group']} batches per group: {tl['us_per_batch']:.1f} us per batch, {tl['avg_kernels_running']} kernels running on average.\n",
      "| kernel | solo us (8-batch group) | mean us with 8 groups in flight | average number running |", "|---|---|---|---|"]
for k, v in tl["kernels"].items():
    md.append(f"| {k} | {v['solo_us']} | {v['mean_us_contended']} | {v['avg_concurrent']} |")
md += ["", "Reading: the two wide kernels (decompression, bucket accumulation) run 1.8-2.7x their solo time because 1.2-1.3 of them are in flight on average and share the multiply pipe; the",
       "latency-bound stages (transcript, head, bucket reduction, window combination, scalar sums: one to three warps per SM) run 3.6-4.5x their solo time — each of their dependent multiplies queues",
       "behind the wide kernels' on the same pipe — and make up three quarters of a group's ~6.9 ms chain.  Throughput = groups in flight / chain length, and adding groups lengthens the chain in",
       "proportion (`r2_experiments.md` §2): the machine is saturated at ~58 % pipe-busy by the mix, not by launches (1.75 per batch) or by the host (`bench.py` `diag.host_issue_ms_per_step`)."]
open(os.path.join(P, "r2_timeline.md"), "w").write("\n".join(md) + "\n")
print("pipe busy fraction over the step:", round(frac, 3), "busy cycles/SM/group:", round(total))
