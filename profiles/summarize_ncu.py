"""Turn gpurun_out/*.ncu-rep and the launch-list CSV into the small text summaries kept under profiles/.
    python profiles/summarize_ncu.py <tag> <launches.csv> <rep1.ncu-rep> [<rep2.ncu-rep> ...]
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[mi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.write("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n\n")
    out.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{n}` | {c} | {t / 1e3:.1f} | {t / tot * 100:.1f}% | {t / c / 1e3:.1f} |\n")
    out.write("\n")


TRAFFIC = {}


def _bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out.write(f"## ncu --set full: {path.split('/')[-1]}\n\n")
    for r in rows[2:]:
        try:  # DRAM bytes per launch (first capture of each kernel) for bench.py's roofline.traffic
            name = r[hdr.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").strip()
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            TRAFFIC.setdefault(name, int(_bytes(r[ir], units[ir]) + _bytes(r[iw], units[iw])))
        except (ValueError, IndexError):
            pass
        out.write(f"### `{r[hdr.index('Kernel Name')][:90]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for m in METRICS:
            if m in hdr:
                out.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
        out.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    with open(f"profiles/{tag}_summary.md", "w") as out:
        out.write(f"# ncu evidence, {tag}\n\n")
        if sys.argv[2] != "-":
            launches(sys.argv[2], out)
        for rep in sys.argv[3:]:
            full(rep, out)
    if TRAFFIC:
        import json
        TRAFFIC["_source"] = (f"ncu --set full --clock-control none, profiles/{tag}_summary.md (B200, bench shapes); bytes per launch = "
                              "dram__bytes_read.sum + dram__bytes_write.sum")
        json.dump(TRAFFIC, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print("wrote", f"profiles/{tag}_summary.md")
