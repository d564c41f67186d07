"""Turn the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv        > profiles/rNN_launches.txt
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep        > profiles/rNN_ncu_full.txt
    python profiles/summarize.py sass     gpurun_out/prof.ncu-rep KERNEL > profiles/rNN_sass_KERNEL.txt
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised:")
    print("# compare SHARES, not absolutes).  source: %s" % path)
    print("%-28s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %6d %12.1f %10.1f %6.1f%%" % (k[-28:], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("%-28s %6d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# selected metrics from `ncu --set full --clock-control none --import-source on` (%s)" % path)
    for r in rows[2:]:
        print("\n== %s" % r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", ""))
        for w in WANT:
            if w in ix:
                print("  %-68s %16s %s" % (w, r[ix[w]], units[ix[w]]))
        if "dram__bytes_read.sum" in ix:
            print("  (traffic = dram read + write, per launch)")


def sass(path, kernel):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    ops, smp = collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[hi + 1:]:
        if len(r) < 10 or not r[0].startswith("0x"):
            continue
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        smp[op] += int(r[ix["# Samples"]])
        tot += n
    print("# warp-level SASS opcode histogram of %s (first profiled launch), from the ncu source page of %s" % (kernel, path))
    for op, n in ops.most_common(20):
        print("%-10s %14d %5.1f%%   stall samples %d" % (op, n, 100.0 * n / tot, smp[op]))
    print("%-10s %14d" % ("TOTAL", tot))


if __name__ == "__main__":
    {"launches": launches, "full": full, "sass": sass}[sys.argv[1]](*sys.argv[2:])
