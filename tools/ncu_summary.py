"""Summarise ncu CSV launch lists / .ncu-rep raw pages into the tables committed under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r01_bench.csv [cells]
  python tools/ncu_summary.py raw gpurun_out/prof_r01_ops.ncu-rep
"""
import collections
import csv
import io
import re
import subprocess
import sys

SZ = {"unsigned char": 1, "signed char": 1, "char": 1, "unsigned short": 2, "short": 2, "unsigned int": 4, "int": 4,
      "unsigned long": 8, "long": 8, "float": 4, "double": 8}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ec::", "")


def cast_bytes(name):
    m = re.search(r"CastF<([^,>]+), ([^,>]+)>", name)
    if not m:
        return None
    return SZ.get(m.group(1).strip()), SZ.get(m.group(2).strip())


def launches(path, cells):
    lines = [l for l in open(path) if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = recs.setdefault(int(row["ID"]), {"name": row["Kernel Name"], "grid": row["Grid Size"], "block": row["Block Size"]})
        d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    rows = list(recs.values())
    print(f"| # | kernel | grid x block | time us | dram read MB | dram write MB | dram total MB | algorithmic MB | traffic/alg | GB/s (alg) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    tot = collections.defaultdict(float)
    for i, r in enumerate(rows):
        t = r.get("gpu__time_duration.sum", 0) / 1e3
        rd, wr = r.get("dram__bytes_read.sum", 0) / 1e6, r.get("dram__bytes_write.sum", 0) / 1e6
        cb = cast_bytes(r["name"])
        small = int(re.sub(r"[^0-9,]", "", r["grid"]).split(",")[0] or 0) < 256  # a parity-check window, not a full buffer
        alg = (cb[0] + cb[1]) * cells / 1e6 if cb and cb[0] and cb[1] and not small else None
        fam = short(r["name"]).split("<")[0]
        tot[fam + "_us"] += t
        tot[fam + "_n"] += 1
        if alg:
            tot["cast_alg"] += alg; tot["cast_dram"] += rd + wr; tot["cast_us"] += t
        print(f"| {i} | `{short(r['name'])[:70]}` | {r['grid']} x {r['block']} | {t:.1f} | {rd:.1f} | {wr:.1f} | {rd + wr:.1f} | "
              f"{alg if alg is None else round(alg, 1)} | {'' if not alg else round((rd + wr) / alg, 3)} | {'' if not alg else round(alg / t * 1e3, 0)} |")
    print()
    print("family totals:", {k: round(v, 1) for k, v in tot.items()})
    if tot["cast_us"]:
        print(f"cast/clone launches: algorithmic {tot['cast_alg']:.0f} MB, dram {tot['cast_dram']:.0f} MB "
              f"(ratio {tot['cast_dram'] / tot['cast_alg']:.3f}), {tot['cast_us']:.0f} us under ncu, {tot["cast_alg"] / tot["cast_us"] * 1e3:.0f} GB/s")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    """`path`: a .ncu-rep, or the CSV that `ncu --csv --page raw --log-file` wrote."""
    if path.endswith(".csv"):
        out = "".join(l for l in open(path) if l.strip() and not l.startswith("=="))
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    cols = [c for c in WANT if c in idx]
    print("| kernel | " + " | ".join(c.replace("__", " ").replace(".sum", "").replace(".avg.pct_of_peak_sustained_elapsed", " %")[:28] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for row in rd:
        if len(row) < len(header):
            continue
        vals = []
        for c in cols:
            v = row[idx[c]]
            u = units[idx[c]]
            try:
                f = float(v.replace(",", ""))
                if "bytes" in c:
                    v = f"{f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1) / 1e6:.1f} MB"
                elif c.startswith("gpu__time"):
                    v = f"{f * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(u, 1):.1f} us"
                else:
                    v = f"{f:g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| `{short(row[idx['Kernel Name']])[:60]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8192 * 8192)
    else:
        raw(sys.argv[2])
