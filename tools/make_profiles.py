"""Turns the artefacts a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/.

    python tools/make_profiles.py <round tag, e.g. r01> <launch csv> <layer log> <bench log> [name=ncu-rep ...]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(tag, path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in data:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("demon::<unnamed>::", "").replace("void ", "")
        v = float(r[iv].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[iu], v)
        tot[name] = tot.get(name, 0) + us
        cnt[name] += 1
    T = sum(tot.values())
    lines = ["# %s -- ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3`)" % tag,
             "", "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
             "%d launches, %.1f ms of kernel time in total." % (sum(cnt.values()), T / 1e3), "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append("| `%s` | %d | %.1f | %.1f %% |" % (k[:70], cnt[k], v, 100 * v / T))
    open(os.path.join(OUT, "%s_launches.md" % tag), "w").write("\n".join(lines) + "\n")


def ncu_rep(tag, name, path):
    if path.endswith(".csv"):   # already exported on the GPU box (`ncu -i rep --page raw --csv`): the reports themselves exceed gpurun's 64 MiB return limit
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for i, h in enumerate(hdr):
        if h in KEY_METRICS or h == "Kernel Name":
            d[h] = (vals[i], units[i])
    return name, d


def main():
    tag, launch_csv, layer_log, bench_log = sys.argv[1:5]
    os.makedirs(OUT, exist_ok=True)
    if os.path.isfile(launch_csv):
        launches(tag, launch_csv)
    if os.path.isfile(layer_log):
        open(os.path.join(OUT, "%s_layers.txt" % tag), "w").write(open(layer_log).read())
    if os.path.isfile(bench_log):
        lines = [l for l in open(bench_log).read().splitlines() if l.startswith("{")]
        open(os.path.join(OUT, "%s_bench.json" % tag), "w").write("\n".join(lines) + "\n")
    reps = []
    for arg in sys.argv[5:]:
        name, path = arg.split("=", 1)
        if os.path.isfile(path):
            reps.append(ncu_rep(tag, name, path))
    if reps:
        lines = ["# %s -- `ncu --set full --clock-control none`, one launch per kernel/shape (batch 64)" % tag, ""]
        traffic = {}
        for name, d in reps:
            lines += ["## %s" % name, "", "kernel: `%s`" % d.get("Kernel Name", ("?", ""))[0][:120], "", "| metric | value | unit |", "|---|---|---|"]
            for k in KEY_METRICS:
                if k in d:
                    lines.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
            lines.append("")
            try:
                def tobytes(x):
                    v, u = float(x[0].replace(",", "")), x[1]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic[name] = tobytes(d["dram__bytes_read.sum"]) + tobytes(d["dram__bytes_write.sum"])
            except Exception:
                pass
        open(os.path.join(OUT, "%s_ncu_full.md" % tag), "w").write("\n".join(lines) + "\n")
        json.dump(traffic, open(os.path.join(OUT, "%s_traffic.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
