"""Turn gpurun_out/ artefacts into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>      e.g. r01_cg2
Reads (whatever exists): gpurun_out/launches.csv (ncu launch list), gpurun_out/prof_gemm.ncu-rep
(ncu --set full of the GEMM), gpurun_out/sweep.json, gpurun_out/bench*.log."""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
           "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
           "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed"]

lst = os.path.join(G, "launches.csv")
if os.path.exists(lst):
    lines = [l for l in open(lst) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0][-70:]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "%s_launch_list.md" % tag), "w") as f:
        f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` "
                "over `bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n" % tag)
        f.write("%d launches captured (cold-cache, serialised: compare SHARES, not absolutes).\n\n" % len(rows))
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% | %.1f |\n" % (k, a[0], a[1], 100 * a[1] / tot, 1e3 * a[1] / a[0]))
    with open(os.path.join(P, "%s_launch_list.csv" % tag), "w") as f:
        f.writelines(lines)
    print("launch list:", len(rows), "launches")

sources = {}
for rep in sorted(glob.glob(os.path.join(G, "prof_*.ncu-rep"))):
    sources[os.path.basename(rep)[:-8]] = ("rep", rep)
for rawcsv in sorted(glob.glob(os.path.join(G, "prof_*_raw.csv"))):   # converted on the GPU box
    if os.path.getsize(rawcsv):
        sources[os.path.basename(rawcsv)[:-8]] = ("csv", rawcsv)
for name, (kind, path) in sorted(sources.items()):
    if kind == "rep":
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        raw = open(path).read()
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, "%s_%s_ncu.md" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none: %s (%s), %d launches\n\n" % (name, tag, len(rows) - 2))
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(rows) - 2)) + " |\n")
        f.write("|---|---|" + "---:|" * (len(rows) - 2) + "\n")
        for i, h in enumerate(hdr):
            if h == "Kernel Name" or any(h.endswith(m) or h == m for m in METRICS):
                f.write("| %s | %s | %s |\n" % (h, units[i], " | ".join(r[i][:60] for r in rows[2:])))
    print("ncu summary:", name)

sw = os.path.join(G, "sweep.json")
if os.path.exists(sw) and os.path.getsize(sw):
    d = json.load(open(sw))
    with open(os.path.join(P, "%s_sweep.md" % tag), "w") as f:
        f.write("# Elementwise / reduce sweep (BASELINE config 3), %s\n\nGB/s = algorithmic bytes / best-of-10 "
                "CUDA-event time; peak = measured HBM copy %.0f GB/s.\n\n" % (tag, d["hbm_peak_gbs"]))
        ops = collections.OrderedDict()
        for r in d["rows"]:
            ops.setdefault(r["op"], {})[r["log2_n"]] = r
        sizes = sorted({r["log2_n"] for r in d["rows"]})
        f.write("| op (B/elem) | " + " | ".join("2^%d" % s for s in sizes) + " |\n|---|" + "---:|" * len(sizes) + "\n")
        for op, by in ops.items():
            bpe = next(iter(by.values()))["bytes_per_elem"]
            f.write("| %s (%d) | " % (op, bpe) + " | ".join(
                "%.0f (%.0f%%)" % (by[s]["gbs"], 100 * by[s]["frac_of_measured_hbm"]) for s in sizes) + " |\n")
        if "cpu_numpy_oracle" in d:
            f.write("\nCPU numpy oracle on %d host cores (GB/s): " % d.get("cpu_cores", 0))
            f.write(", ".join("%s@2^%d %.1f" % (c["op"], c["log2_n"], c["gbs"]) for c in d["cpu_numpy_oracle"]) + "\n")
    json.dump(d, open(os.path.join(P, "%s_sweep.json" % tag), "w"))
    print("sweep summary")

with open(os.path.join(P, "%s_bench.jsonl" % tag), "w") as f:
    for b in sorted(glob.glob(os.path.join(G, "bench*.log"))):
        for line in open(b):
            if line.startswith("{"):
                f.write(json.dumps({"source": os.path.basename(b), **json.loads(line)}) + "\n")
print("bench lines")
