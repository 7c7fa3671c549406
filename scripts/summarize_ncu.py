"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.md
  python scripts/summarize_ncu.py full gpurun_out/price_r01.ncu-rep profiles/r01_price_full.md [m n]
"""
import collections, csv, json, os, subprocess, sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(src)})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` (cold-cache, serialised: "
                "compare SHARES, not absolutes).\n\n")
        f.write(f"total kernel time {T/1e3:.2f} ms over {sum(cnt.values())} launches\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {v/T:.4f} | {v/cnt[k]:.1f} |\n")


def full(src, dst, m=None, n=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(src)})\n\n")
        kn = sorted({r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0] for r in data})
        f.write(f"`ncu --set full --clock-control none --import-source on -k regex:{'|'.join(kn)}` over `bench.py`.\n\n")
        for r in data:
            f.write(f"## launch id {r[0]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w, i in idx:
                f.write(f"| {w} | {r[i][:90]} | {units[i]} |\n")
            f.write("\n")
    if m and n:
        big = max(data, key=lambda r: float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")))
        def to_bytes(name):
            i = hdr.index(name)
            v = float(big[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        t = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        json.dump({"m": int(m), "n": int(n), "dram_bytes_per_launch": t, "kernel": big[hdr.index("Kernel Name")].split("(")[0],
                   "source": os.path.basename(dst)}, open(os.path.join(os.path.dirname(dst), "price_traffic.json"), "w"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
