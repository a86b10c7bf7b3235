"""Turns gpurun_out/launches.csv (ncu launch list) + gpurun_out/prof_r01.ncu-rep (ncu --set full) into
profiles/rNN_ncu_summary.json.  usage: python tools/summarize_ncu.py r01"""
import collections
import csv
import json
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
L = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[idx["Metric Unit"]], 1)
    L.append((r[idx["Kernel Name"]], ns))
tc = [i for i, (n, _) in enumerate(L) if "search_aug_kernel" in n or "search_tc_kernel" in n]
a, b = tc[3], tc[4]                      # one timed step: from the codebook refresh before one search to the next
first = [i for i, (n, _) in enumerate(L) if "codebook_absmax_kernel" in n]
sa, sb = max(i for i in first if i < a), max(i for i in first if i < b)
step = L[sa:sb]
tot = sum(x[1] for x in step)
agg = collections.OrderedDict()
for n, ns in step:
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    agg[n] = agg.get(n, 0) + ns
kern = [{"name": n, "us": ns / 1e3, "share_pct": 100 * ns / tot} for n, ns in sorted(agg.items(), key=lambda kv: -kv[1])]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h2, units, data = rr[0], rr[1], rr[2:]
ix = {h: i for i, h in enumerate(h2)}
want = {"dur_ms": "gpu__time_duration.sum", "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "sm_ghz": "sm__cycles_elapsed.avg.per_second", "l2_rd_sectors": "lts__t_sectors_srcunit_tex_op_read.sum",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}


def conv(v, u, key):
    v = float(v.replace(",", ""))
    if key.endswith("MB"):
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    if key == "dur_ms":
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[u]
    return v


full = []
for r in data:
    d = {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")}
    for k, m in want.items():
        if m in ix:
            d[k] = round(conv(r[ix[m]], units[ix[m]], k), 4)
    d["dram_GBs"] = round((d["dram_rd_MB"] + d["dram_wr_MB"]) / d["dur_ms"], 1)
    full.append(d)
out = {"launch_list": {"step_ms_under_ncu": tot / 1e6, "launches_in_step": len(step), "kernels": kern}, "ncu_full": full}
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
print(f"one step under ncu: {tot / 1e6:.3f} ms, {len(step)} launches")
for k in kern[:10]:
    print(f"  {k['share_pct']:5.1f}%  {k['us']:9.1f} us  {k['name'][:80]}")
for d in full[:4]:
    print(d)
