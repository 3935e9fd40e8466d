"""Turns the ncu artefacts of one GPU visit (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> [batch]   (e.g. r2 256)"""
import csv
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.makedirs(P, exist_ok=True)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__inst_executed.sum"]

lines = ["# ncu summaries (%s)\n" % tag,
         "Captured with `ncu --set full --clock-control none --import-source on` on one B200 while running "
         "`tools/profile_step.py --batch %d --steps 3` (base config, %d sequences x 2096 tokens; `tools/ncu_round.sh`); one launch "
         "per kernel. `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of that launch.\n" % (batch, batch)]
summary = {}
traffic = {}
# kernel-name fragments -> the names bench.py gives the launches (md_* entry + shape detail)
BENCH_KEYS = [("attention_kernel", None, "md_attention_bf16"), ("layernorm_kernel", None, "md_layernorm_bf16"),
              ("posterior_step", None, "md_posterior_step"), ("round_tc_kernel", None, "md_round_argmin_tc"),
              ("gemm_pair_kernel<1", None, "md_linear_bf16:%dx3072x768 epi=1" % (batch * 2096))]
EPI0_ORDER = ["md_linear_bf16:%dx2304x768 epi=0", "md_linear_bf16:%dx768x768 epi=0", "md_linear_bf16:%dx768x3072 epi=0"]
epi0_seen = 0
for f in sorted(os.listdir(G)):
    if not (f.startswith("prof_") and f.endswith(".ncu-rep")):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(G, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {}
        for w in WANT:
            if w in hdr:
                d[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
        stalls = {h.split("stalled_")[1].replace("_per_issue_active.ratio", ""): float(r[i]) for i, h in enumerate(hdr)
                  if "issue_stalled" in h and "per_issue_active" in h and r[i] not in ("", "n/a")}
        d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        summary[f + "::" + name[:80] + ("#%d" % len(summary))] = d
        try:
            tr = 0.0
            for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = r[hdr.index(w)], units[hdr.index(w)]
                tr += float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            key = None
            for frag, _, k in BENCH_KEYS:
                if frag in name:
                    key = k
            if key is None and "gemm_pair_kernel<0" in name and epi0_seen < 3:
                key = EPI0_ORDER[epi0_seen] % (batch * 2096)
                epi0_seen += 1
            if key is not None and key not in traffic:
                traffic[key] = {"dram_bytes_per_launch": tr, "batch": batch, "capture": f, "kernel": name[:100],
                                "gpu_time_ms": d.get("gpu__time_duration.sum")}
        except (ValueError, KeyError):
            pass
        lines.append("\n## %s — `%s`\n" % (f, name[:100]))
        for k, v in d.items():
            lines.append("* %s: %s" % (k, v))
open(os.path.join(P, "%s_ncu_summary.md" % tag), "w").write("\n".join(lines) + "\n")
json.dump(summary, open(os.path.join(P, "%s_ncu_summary.json" % tag), "w"), indent=1)
if traffic:
    json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % tag), "w"), indent=1)

# launch list -> per-kernel shares
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui].startswith("n") else v / 1e3 if r[ui].startswith("u") else v
        a = agg.setdefault(r[ki][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown`\n" % tag,
           "Cold-cache, serialised per-launch times: compare SHARES, not absolutes.  %d launches, %.1f ms total.\n" % (sum(a[0] for a in agg.values()), tot),
           "| share | total ms | launches | avg ms | kernel |", "|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %.1f%% | %.2f | %d | %.3f | `%s` |" % (100 * a[1] / tot, a[1], a[0], a[1] / a[0], k))
    open(os.path.join(P, "%s_launch_list.md" % tag), "w").write("\n".join(out) + "\n")
bj = os.path.join(G, "bench.json")
if os.path.exists(bj):
    txt = open(bj).read().strip().splitlines()
    if txt:
        json.dump(json.loads(txt[-1]), open(os.path.join(P, "%s_bench.json" % tag), "w"), indent=1)
print("wrote", [f for f in os.listdir(P) if f.startswith(tag)])
