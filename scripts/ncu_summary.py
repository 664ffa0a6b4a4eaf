"""Summarise an ncu launch list (CSV of gpu__time_duration.sum) and a `--set full` report into a markdown table.

usage: python scripts/ncu_summary.py launches.csv prof.ncu-rep > profiles/<round>_ncu_full_summary.md
Reads the report here (no GPU needed) through `ncu -i ... --page raw --csv`."""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.strip()[:90]


def launches(path):
    rows = []
    with open(path) as f:
        text = f.read()
    text = text[text.index('"ID"'):]
    for r in csv.DictReader(io.StringIO(text)):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3))
    return rows


def main():
    lpath, rep = sys.argv[1], sys.argv[2]
    rows = launches(lpath)
    tot = sum(t for _, t in rows)
    agg = collections.OrderedDict()
    for k, t in rows:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += t
        a[1] += 1
    ours = sum(t for k, (t, _) in agg.items() if "gpcsd::" in k)
    print("## Launch list of one step (`%s`): %d launches, %.2f ms serialised/cold; our kernels %.1f %%\n" %
          (lpath.split("/")[-1], len(rows), tot / 1e3, 100 * ours / tot))
    print("| time (us) | share | launches | kernel |\n|---|---|---|---|")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print("| %.1f | %.1f %% | %d | `%s` |" % (t, 100 * t / tot, c, k))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in data:
        name = short(r[col["Kernel Name"]])
        key = (name, r[col["Grid Size"]], r[col["Block Size"]])
        if key in seen:
            continue
        seen.add(key)
        print("\n## %s  (grid %s x block %s)\n\n| metric | value | unit |\n|---|---|---|" % (name, r[col["Grid Size"]], r[col["Block Size"]]))
        for m in KEEP:
            if m in col:
                print("| %s | %s | %s |" % (m, r[col[m]], units[col[m]]))


if __name__ == "__main__":
    main()
