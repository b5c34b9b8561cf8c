"""Summarise ncu --set full reports (read with the local ncu) into a markdown table of the metrics the roofline
discussion uses, and refresh profiles/dram_traffic.json.  Usage: python tools/ncu_summary.py TAG rep1 [rep2 ...]"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    return hdr, units, rd[2:]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = [f"# ncu --set full --clock-control none summaries ({tag})",
          "Times under ncu are cold-cache and serialised; use shares, not absolutes.", ""]
    traffic = {}
    per_kernel = {}
    for rep in reps:
        hdr, units, data = rows(rep)
        ki = hdr.index("Kernel Name")
        count = {}
        for r in data:
            name = r[ki].split("(")[0].split("::")[-1]
            count[name] = count.get(name, 0) + 1
            md += [f"## {name}  launch {count[name]}  ({os.path.basename(rep)})", "", "| metric | value | unit |", "|---|---|---|"]
            tot = 0.0
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    md.append(f"| {m} | {r[i]} | {units[i]} |")
                    if m.startswith("dram__bytes_"):
                        tot += float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
            md.append("")
            per_kernel.setdefault(name.split("<")[0], []).append(tot)
    traffic = {k: int(sum(v) / len(v)) for k, v in per_kernel.items()}  # mean over the captured launches of a kernel
    open(os.path.join(root, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(md))
    tj_path = os.path.join(root, "profiles", "dram_traffic.json")
    tj = json.load(open(tj_path))
    for k, v in traffic.items():
        if k.startswith("loss_"):
            tj[k] = {"config": "N=32768 d=768 n_gpus=1", "bytes_per_launch": v, "source": f"gpurun_out/prof_loss_{tag}.ncu-rep (mean over the captured launches of one step)"}
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
