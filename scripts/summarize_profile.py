"""Turns gpurun_out/ ncu artefacts into the committed text summaries under profiles/.

    python scripts/summarize_profile.py launches <launches.csv> <out.md> [first_id last_id]
    python scripts/summarize_profile.py full <report.ncu-rep> <out.md>

`launches` = the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: shares matter,
not absolutes); `full` = selected metrics of an `ncu --set full` capture read with
`ncu -i ... --page raw --csv`.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def launches(path, out, lo=None, hi=None):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    if lo is not None:
        data = [d for d in data if lo <= int(d["ID"]) <= hi]
    total = sum(float(d["Metric Value"]) for d in data)
    agg = OrderedDict()
    for d in data:
        k = d["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"])
    with open(out, "w") as f:
        f.write(f"source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none), launches "
                f"{data[0]['ID']}..{data[-1]['ID']}\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {n} | {t / 1e3:.1f} | {100 * t / total:.1f}% |\n")
        f.write(f"\ntotal {total / 1e3:.1f} us over {len(data)} launches\n")
    print(open(out).read())


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"source: {path} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"- {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
            stalls = [(hdr[i], float(r[i] or 0)) for i in range(len(hdr))
                      if hdr[i].startswith("smsp__average_warps_issue_stalled") and "not_issued" not in hdr[i]]
            top = sorted(stalls, key=lambda t: -t[1])[:5]
            f.write("- top stalls (warps per issue): " + ", ".join(
                f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                for h, v in top) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], *(int(a) for a in sys.argv[4:6]))
    else:
        full(sys.argv[2], sys.argv[3])
