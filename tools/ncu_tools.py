"""Read ncu outputs brought back from the GPU box (gpurun_out/) here on the CPU box and print the summaries that are
committed under profiles/.
  python tools/ncu_tools.py list  gpurun_out/<tag>_launches.csv          per-kernel share of one step
  python tools/ncu_tools.py full  gpurun_out/<tag>_<kernel>.ncu-rep      key metrics of an `ncu --set full` capture
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
]


def launch_list(path):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\((?:const |b200::|CUtensorMap|int|__nv).*$", "", name).replace("b200::", "")
        key = f"{name} grid={r['Grid Size']}"
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {total / 1e6:.2f} ms of kernel time (ncu: cold caches, serialised; compare shares)")
    print(f"# {'total us':>10s} {'count':>6s} {'avg us':>9s} {'share':>6s}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:12.1f} {v[0]:6d} {v[1] / v[0] / 1e3:9.1f} {100 * v[1] / total:5.1f}%  {k}")
    fam = collections.Counter()
    for k, v in agg.items():
        fam[k.split("<")[0].split(" ")[0]] += v[1]
    print("# by kernel family")
    for k, v in fam.most_common():
        print(f"{v / 1e3:12.1f} us {100 * v / total:5.1f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")], " block",
              r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]:>18s} {units[i]}")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tmul = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
                "second": 1.0}

        def val(key, table):
            i = hdr.index(key)
            return float(r[i].replace(",", "")) * table[units[i]]
        traffic = val("dram__bytes_read.sum", mult) + val("dram__bytes_write.sum", mult)
        t = val("gpu__time_duration.sum", tmul)
        print(f"  => DRAM traffic {traffic / 1e6:.1f} MB in {t * 1e6:.1f} us = {traffic / t / 1e9:.0f} GB/s "
              f"(under ncu: caches flushed before every replay)")
        print()


if __name__ == "__main__":
    {"list": launch_list, "full": full}[sys.argv[1]](sys.argv[2])
