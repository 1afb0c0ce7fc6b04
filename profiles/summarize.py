#!/usr/bin/env python3
"""Turns the raw ncu outputs of a round into the tracked summaries under profiles/.

  python profiles/summarize.py rNN gpurun_out/launches.csv gpurun_out/prof.ncu-rep [workload] [batch]

  launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
                 of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
  prof.ncu-rep : ncu --set full --clock-control none --import-source on of the top kernels

Writes profiles/rNN_launches.csv (our kernels only), profiles/rNN_summary.md and updates
profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py).
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    keep = [hdr]
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if name.startswith("k_") or "cub::" in name:
            keep.append(r)
        a = agg.setdefault(name[:70], [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += float(r[ix["Metric Value"]])
    return agg, keep


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    res = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
        res[name] = {w: (r[ix[w]], units[ix[w]]) for w in want if w in ix}
    return res


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def stalls(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + kernel], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    try:
        hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
    except StopIteration:
        return []
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        for k in hdr:
            if k.startswith("stall_") and "Not" not in k:
                try:
                    tot[k] += int(r[ix[k]])
                except ValueError:
                    pass
    s = sum(tot.values()) or 1
    return [(k, 100.0 * v / s) for k, v in tot.most_common(6)]


def main():
    tag, launches, rep = sys.argv[1:4]
    workload = sys.argv[4] if len(sys.argv) > 4 else "C2"
    batch = sys.argv[5] if len(sys.argv) > 5 else "?"
    agg, keep = launch_table(launches)
    with open(os.path.join(HERE, tag + "_launches.csv"), "w", newline="") as f:
        csv.writer(f).writerows(keep)
    total = sum(a[1] for a in agg.values())
    md = ["# %s — ncu summary (workload %s, %s pictures per launch group)" % (tag, workload, batch), "",
          "## Launch list shares (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
          "cold-cache and serialised: shares, not absolutes)", "",
          "| kernel | launches | total ms | share | grid | block |", "|---|---:|---:|---:|---|---|"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if a[1] / total < 0.0005:
            continue
        md.append("| `%s` | %d | %.3f | %.1f %% | %s | %s |" % (n, a[0], a[1] / 1e6, 100 * a[1] / total, a[2], a[3]))
    met = raw_metrics(rep)
    traffic_path = os.path.join(HERE, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    traffic.setdefault(workload, {})
    md += ["", "## Top kernels (`ncu --set full --clock-control none --import-source on`)", ""]
    for name, m in met.items():
        rd = to_bytes(*m["dram__bytes_read.sum"])
        wr = to_bytes(*m["dram__bytes_write.sum"])
        md += ["### `%s`" % name, "", "| metric | value |", "|---|---|"]
        for k, (v, u) in m.items():
            md.append("| %s | %s %s |" % (k, v, u))
        md.append("| DRAM traffic per launch | %.3f GB |" % ((rd + wr) / 1e9))
        st = stalls(rep, name)
        if st:
            md.append("| warp stall samples | %s |" % ", ".join("%s %.0f %%" % (k.replace("stall_", ""), v) for k, v in st))
        md.append("")
        key = {"k_code_range": "code", "k_decode": "decode", "k_symbolize_planar": "symbolize"}.get(name)
        if key:
            traffic[workload][key] = rd + wr
            traffic[workload][key + "_batch"] = batch
    open(os.path.join(HERE, tag + "_summary.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("\n".join(md))


if __name__ == "__main__":
    main()
