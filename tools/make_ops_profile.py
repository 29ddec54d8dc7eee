"""profiles/<tag>_ncu_ops.md from an `ncu --set full` report of tools/run_ops_once.py (the standalone geometry ops at
[8,3,768,1024]): duration, DRAM bytes, algorithmic bytes / duration against the measured copy peak, and the issue-slot
side of the roofline (warp instructions per pixel, issue-active %), which is what bounds the bit-exact arithmetic.

    python tools/make_ops_profile.py r02 gpurun_out/r02_ops.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALG_MB = {"warp2d": 201.3, "depth_to_flow": 75.5, "flow_to_depth": 75.5, "median3x3": 94.4, "sig_": 75.5, "leaky_relu": 151.0}
PIXELS = {"warp2d": 8 * 768 * 1024, "depth_to_flow": 8 * 768 * 1024, "flow_to_depth": 8 * 768 * 1024, "median3x3": 8 * 3 * 384 * 512,
          "sig_": 8 * 768 * 1024, "leaky_relu": 8 * 3 * 768 * 1024}


def main():
    tag, rep = sys.argv[1], sys.argv[2]
    peak = 6590.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            peak = float(json.load(open(p)).get("hbm_gbs", peak))
        except Exception:
            pass
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=None):
        v, u = float(r[ix[name]].replace(",", "")), units[ix[name]]
        if scale == "us":
            return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        if scale == "MB":
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
        return v

    lines = ["# %s -- standalone geometry ops at [8,3,768,1024] (BASELINE.json configs[4]), `ncu --set full --clock-control none`, one launch each" % tag, "",
             "Times under ncu are cold-cache single launches.  `algorithmic GB/s` = algorithmic bytes / duration (what the op must move;",
             "outputs that stay in the 126 MB L2 do not show up as DRAM writes); copy peak of this pool: %.0f GB/s (MEASURED_PEAKS.json)." % peak,
             "`instr / px` = warp instructions x 32 / output pixels: the ops restate x86 float arithmetic operation for operation (IEEE",
             "division, no FMA contraction, cvttss2si emulation), so where `issue active` is high the kernel sits on the issue roofline, not on HBM.", "",
             "| kernel | duration us | dram read MB | dram write MB | algorithmic MB | algorithmic GB/s | frac of copy peak | issue active % | instr / px | regs | warps active % |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = next((k for k in ALG_MB if k in name), None)
        if key is None:
            continue
        dur = val(r, "gpu__time_duration.sum", "us")
        alg = ALG_MB[key]
        inst = val(r, "smsp__inst_executed.sum")
        lines.append("| `%s` | %.1f | %.1f | %.1f | %.1f | %.0f | %.2f | %.0f | %.0f | %d | %.0f |" % (
            name.split("(")[0].replace("void ", "").replace("demon::", "")[:44], dur, val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB"), alg,
            alg / dur * 1e3, alg / dur * 1e3 / peak,
            val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), inst * 32 / PIXELS[key], int(val(r, "launch__registers_per_thread")),
            val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
    open(os.path.join(ROOT, "profiles", "%s_ncu_ops.md" % tag), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
