"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01

Inputs (produced by tools/gpu_ncu_full.sh on a B200):
  gpurun_out/launches_c2.csv     ncu --metrics gpu__time_duration.sum launch list of one bench step (c2 batch)
  gpurun_out/prof_{gemm,attn,conv0,ln}.ncu-rep   ncu --set full captures
Outputs:
  profiles/launches_<round>.csv          per-kernel launch count / total time / share of the step
  profiles/ncu_<round>_<kernel>.csv      selected raw metrics per captured launch
  (profiles/SUMMARY_<round>.md is written by hand from these + bench.py output)
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
METRICS = [
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def launch_list(tag, src="launches_c2.csv", suffix=""):
    path = os.path.join(SRC, src)
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for row in r:
        if len(row) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", row[idx["Kernel Name"]])
        val = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]]
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, "launches_%s%s.csv" % (tag, suffix)), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_of_step"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], "%.1f" % v[1], "%.4f" % (v[1] / tot)])
        w.writerow(["TOTAL", sum(v[0] for v in agg.values()), "%.1f" % tot, "1.0"])
    print("wrote launches_%s%s.csv (%d kernels, %.2f ms)" % (tag, suffix, len(agg), tot / 1e3))


def ncu_raw(tag, name):
    """prof_<name>.ncu-rep, or its raw-page CSV export prof_<tag>_<name>_raw.csv made on the GPU box (gpurun copies back <= 64 MiB)."""
    rep = os.path.join(SRC, "prof_%s.ncu-rep" % name)
    raw = os.path.join(SRC, "prof_%s_%s_raw.csv" % (tag, name))
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in ["Kernel Name"] + METRICS if m in hdr]
    with open(os.path.join(OUT, "ncu_%s_%s.csv" % (tag, name)), "w") as f:
        w = csv.writer(f)
        w.writerow([m for m, _ in cols])
        w.writerow([units[i] for _, i in cols])
        for d in data:
            w.writerow([re.sub(r"\(CUtensorMap.*", "", d[i]) if m == "Kernel Name" else d[i] for m, i in cols])
    print("wrote ncu_%s_%s.csv (%d launches)" % (tag, name, len(data)))


CLASSES = [("gemm_tc", "gemm_tc_bf16"), ("layernorm_kernel", "layernorm"), ("conv0_tc", "conv0_gn_gelu"), ("conv0_apply", "conv0_gn_gelu"),
           ("attention_tc", "attention_tc_bf16"), ("posconv_tc", "posconv_tc_bf16"), ("memory_attention", "memory_attention")]


def traffic(tag):
    """gpurun_out/traffic_c3.csv (tools/gpu_traffic.sh) -> profiles/traffic_<tag>.json: DRAM bytes per launch per kernel class."""
    import json
    path = os.path.join(SRC, "traffic_%s_c3.csv" % tag)
    if not os.path.exists(path):
        path = os.path.join(SRC, "traffic_c3.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for row in r:
        if len(row) < len(hdr):
            continue
        name = row[idx["Kernel Name"]]
        cls = next((c for key, c in CLASSES if key in name), None)
        if cls is None:
            continue
        metric, unit = row[idx["Metric Name"]], row[idx["Metric Unit"]]
        val = float(row[idx["Metric Value"]].replace(",", ""))
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        a = per.setdefault(cls, {"ids": set(), "read": 0.0, "write": 0.0})
        a["ids"].add(row[idx["ID"]])
        if metric == "dram__bytes_read.sum":
            a["read"] += val
        elif metric == "dram__bytes_write.sum":
            a["write"] += val
    out = {"workload": "c3, 24 utterances (4 batches), bf16, eager launches, ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum "
                       "--clock-control none", "kernels": {}}
    for cls, a in per.items():
        n = len(a["ids"])
        out["kernels"][cls] = {"launches": n, "dram_read_bytes_per_launch": a["read"] / n, "dram_write_bytes_per_launch": a["write"] / n,
                               "traffic_bytes_per_launch": (a["read"] + a["write"]) / n}
    with open(os.path.join(OUT, "traffic_%s.json" % tag), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote traffic_%s.json (%s)" % (tag, ", ".join("%s x%d" % (k, v["launches"]) for k, v in out["kernels"].items())))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launch_list(tag)
    launch_list(tag, "launches_%s_default.csv" % tag, "_default")      # the default bench command (c3, graphs)
    launch_list(tag, "launches_%s_c5.csv" % tag, "_c5")                # the training step
    launch_list(tag, "launches_%s_beam.csv" % tag, "_beam")
    for n in ("gemm", "attn", "conv0", "posconv", "ln", "c5"):
        ncu_raw(tag, n)
    traffic(tag)
