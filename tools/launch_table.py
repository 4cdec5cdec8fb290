"""Per-launch table from an `ncu --metrics ... --csv --log-file` launch list (time, DRAM bytes, pipe shares)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault(r["ID"], {"name": r["Kernel Name"].replace("void <unnamed>::", "")[:64], "grid": r["Grid Size"]})[r["Metric Name"]] = r["Metric Value"]
    f = lambda d, m: float(d.get(m, "nan").replace(",", "")) if d.get(m, "n/a") != "n/a" else float("nan")
    print(f"{'id':>3} {'kernel':64} {'grid':>16} {'ms':>8} {'rd GB':>7} {'wr GB':>7} {'GB/s':>6} {'dram%':>6} {'fp64%':>6} {'issue%':>6}")
    for i, d in by.items():
        t = f(d, "gpu__time_duration.sum") / 1e6
        rd, wr = f(d, "dram__bytes_read.sum") / 1e9, f(d, "dram__bytes_write.sum") / 1e9
        print(f"{i:>3} {d['name']:64} {d['grid']:>16} {t:8.3f} {rd:7.2f} {wr:7.2f} {(rd + wr) / t * 1e3:6.0f} "
              f"{f(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{f(d, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{f(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
