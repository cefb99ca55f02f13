"""Turns the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python scripts/summarize_profiles.py r01 [--launches gpurun_out/launches2.csv] [--dram gpurun_out/tc_dram.csv] [--rep gpurun_out/prof_tc.ncu-rep]
"""
import argparse
import collections
import csv
import json
import os
import re
import subprocess
import sys


def read_ncu_csv(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"h3d::(<unnamed>::)?|unnamed>::", "", n)
    return re.sub(r"\(.*", "", n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default="gpurun_out/launches.csv")
    ap.add_argument("--dram", default="gpurun_out/tc_dram.csv")
    ap.add_argument("--rep", default="gpurun_out/prof_tc.ncu-rep")
    a = ap.parse_args()
    os.makedirs("profiles", exist_ok=True)
    out = {}
    if os.path.exists(a.launches):
        rows = read_ncu_csv(a.launches)
        names = [(short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"]) for r in rows]
        ends = [i for i, (n, _, _) in enumerate(names) if "CatArray" in n]
        s, e = (ends[2] + 1, ends[3] + 1) if len(ends) > 3 else (0, len(names))
        step = names[s:e]
        agg = collections.OrderedDict()
        for n, v, g in step:
            agg.setdefault(n, [0.0, 0])
            agg[n][0] += v; agg[n][1] += 1
        tot = sum(v for _, v, _ in step)
        with open("profiles/%s_launches.md" % a.tag, "w") as f:
            f.write("# %s: every launch of one bench step (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % a.tag)
            f.write("Command: `ncu ... python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (B = 32, 320x320, bf16x3). "
                    "Per-launch times are cold-cache and serialised: compare SHARES.\n\n")
            f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
            for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
                f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, v, 100 * v / tot))
            f.write("| **sum** | %d | %.1f | 100 %% |\n\n" % (len(step), tot))
            f.write("## launch list\n\n| # | kernel | grid | us |\n|---:|---|---|---:|\n")
            for i, (n, v, g) in enumerate(step):
                f.write("| %d | `%s` | %s | %.1f |\n" % (i, n, g, v))
        out["launches"] = {"step_us": tot, "by_kernel_us": {n: v for n, (v, c) in agg.items()}}
    if os.path.exists(a.dram):
        rows = read_ncu_csv(a.dram)
        by = collections.OrderedDict()
        for r in rows:
            by.setdefault(r["ID"], {"name": short(r["Kernel Name"])})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        rd = sum(v.get("dram__bytes_read.sum", 0) for v in by.values())
        wr = sum(v.get("dram__bytes_write.sum", 0) for v in by.values())
        t = sum(v.get("gpu__time_duration.sum", 0) for v in by.values())
        n = len(by)
        tp = [v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) for v in by.values()]
        tw = sum(p * v.get("gpu__time_duration.sum", 0) for p, v in zip(tp, by.values())) / max(t, 1)
        out["tc_conv"] = {"launches_per_step": n, "dram_bytes_per_step": rd + wr, "dram_bytes_per_launch": (rd + wr) / n,
                          "time_us_per_step": t / 1e3, "tensor_pipe_active_pct_time_weighted": tw}
        with open("profiles/%s_tc_conv_dram.md" % a.tag, "w") as f:
            f.write("# %s: DRAM traffic and tensor-pipe activity of every tcgen05 conv launch of one step\n\n" % a.tag)
            f.write("`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active...`\n\n")
            f.write("| # | kernel | read MB | write MB | us | tensor pipe active %% |\n|---:|---|---:|---:|---:|---:|\n")
            for i, v in enumerate(by.values()):
                f.write("| %d | `%s` | %.1f | %.1f | %.1f | %.1f |\n" % (
                    i, v["name"], v.get("dram__bytes_read.sum", 0) / 1e6, v.get("dram__bytes_write.sum", 0) / 1e6,
                    v.get("gpu__time_duration.sum", 0) / 1e3, v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)))
            f.write("\nTotal: read %.1f MB + write %.1f MB = %.1f MB per step over %d launches (%.1f MB per launch), %.1f us; "
                    "time-weighted tensor-pipe active %.1f %%.\n" % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6, n, (rd + wr) / n / 1e6, t / 1e3, tw))
    if os.path.exists(a.rep):
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) > 2:
            hdr, units = rows[0], rows[1]
            keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
                    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]
            with open("profiles/%s_tc_conv_ncu_full.md" % a.tag, "w") as f:
                f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:conv_tc` (3 launches of the dominant kernel)\n\n" % a.tag)
                for r in rows[2:]:
                    f.write("| metric | value | unit |\n|---|---:|---|\n")
                    for k in keys:
                        if k in hdr:
                            i = hdr.index(k)
                            f.write("| %s | %s | %s |\n" % (k, r[i][:110], units[i]))
                    f.write("\n")
    json.dump(out, open("profiles/%s_summary.json" % a.tag, "w"), indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
