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
        # one row per (launch, metric): duration always, DRAM bytes / tensor-pipe activity when the stage collected them
        by_id = collections.OrderedDict()
        for r in read_ncu_csv(a.launches):
            d = by_id.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"]})
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        names = [(d["name"], d.get("gpu__time_duration.sum", 0.0) / 1e3, d["grid"], d) for d in by_id.values()]
        ends = [i for i, (n, _, _, _) in enumerate(names) if "CatArray" in n or "pack_records_kernel" in n]   # last kernel of a single-GPU step
        s, e = (ends[2] + 1, ends[3] + 1) if len(ends) > 3 else (0, len(names))
        step = names[s:e]
        have_dram = any("dram__bytes_read.sum" in d for _, _, _, d in step)
        agg = collections.OrderedDict()
        for n, v, g, d in step:
            a_ = agg.setdefault(n, [0.0, 0, 0.0, 0.0, 0.0])
            a_[0] += v; a_[1] += 1
            a_[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            a_[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * v
            a_[4] += d.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * v
        tot = sum(v for _, v, _, _ in step)
        with open("profiles/%s_launches.md" % a.tag, "w") as f:
            f.write("# %s: every launch of one bench step (ncu --metrics gpu__time_duration.sum[,dram__bytes_*,sm__pipe_tensor_cycles_active] --clock-control none)\n\n" % a.tag)
            f.write("Command: `ncu ... python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (B = 32, 320x320, bf16x3). "
                    "Per-launch times are cold-cache and serialised: compare SHARES.  HBM GB/s = (DRAM read + write bytes) / duration "
                    "(measured peak copy bandwidth of this pool: 6576 GB/s, MEASURED_PEAKS.json).\n\n")
            f.write("| kernel | launches | total us | share | DRAM MB | HBM GB/s | tensor pipe active % (time-weighted) | tensor unit busy % (`sm__pipe_tc_cycles_active`) |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
            for n, (v, c, byt, tp, tcb) in sorted(agg.items(), key=lambda x: -x[1][0]):
                f.write("| `%s` | %d | %.1f | %.1f %% | %s | %s | %s | %s |\n" % (
                    n, c, v, 100 * v / tot, "%.1f" % (byt / 1e6) if have_dram else "-",
                    "%.0f" % (byt / 1e3 / v) if have_dram and v > 0 else "-", "%.1f" % (tp / v) if have_dram and v > 0 else "-",
                    "%.1f" % (tcb / v) if have_dram and v > 0 else "-"))
            f.write("| **sum** | %d | %.1f | 100 %% | %s | | | |\n\n" % (len(step), tot, "%.1f" % (sum(x[2] for x in agg.values()) / 1e6) if have_dram else "-"))
            f.write("## launch list\n\n| # | kernel | grid | us | DRAM read MB | DRAM write MB | HBM GB/s | tensor pipe % |\n|---:|---|---|---:|---:|---:|---:|---:|\n")
            for i, (n, v, g, d) in enumerate(step):
                rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
                f.write("| %d | `%s` | %s | %.1f | %s | %s | %s | %s |\n" % (
                    i, n, g, v, "%.2f" % (rd / 1e6) if have_dram else "-", "%.2f" % (wr / 1e6) if have_dram else "-",
                    "%.0f" % ((rd + wr) / 1e3 / v) if have_dram and v > 0 else "-",
                    "%.1f" % d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) if have_dram else "-"))
        out["launches"] = {"step_us": tot, "by_kernel_us": {n: x[0] for n, x in agg.items()}}
        if have_dram:
            tc = [(n, v, d) for n, v, _, d in step if re.match(r"conv_(tc|c64|c1f)", n)]
            byt = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for _, _, d in tc)
            t = sum(v for _, v, _ in tc)
            tw = sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * v for _, v, d in tc) / max(t, 1e-9)
            out["tc_conv"] = {"launches_per_step": len(tc), "dram_bytes_per_step": byt, "dram_bytes_per_launch": byt / max(len(tc), 1),
                              "time_us_per_step": t, "tensor_pipe_active_pct_time_weighted": tw}
            out["hbm_kernels"] = {n: {"us": x[0], "dram_mb": x[2] / 1e6, "gb_per_s": x[2] / 1e3 / x[0] if x[0] > 0 else 0.0}
                                  for n, x in agg.items() if not re.match(r"conv_(tc|c64|c1f)", n)}
    if os.path.exists(a.dram) and "tc_conv" not in out:
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
                f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:conv_(tc|c64)` (the first tensor-core launches of one step: conv1_2 .. conv4_4 of HandSegNet)\n\n" % a.tag)
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
