"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/X_launches.md "title"
  python tools/summarize_ncu.py full gpurun_out/prof_X.ncu-rep profiles/X_full.md "title"
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and "
                f"serialised: read the SHARES)\n\ntotal {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches\n\n"
                "| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {t:.3f} | {100 * t / tot:.1f}% | {n} | `{k}` |\n")


def full(src, dst, title):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nncu --set full --clock-control none --import-source on; source: `{src}` (scratch, not committed)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[idx['Kernel Name']][:100]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            stalls = [(h, r[i]) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
            stalls = sorted(((h, float(v.replace(",", ""))) for h, v in stalls if v), key=lambda kv: -kv[1])[:6]
            if stalls:
                f.write("\ntop stall reasons (warps stalled per issue): " + ", ".join(
                    f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for h, v in stalls) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
