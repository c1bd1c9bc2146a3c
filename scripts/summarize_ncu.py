"""Summarise ncu outputs into small text files for profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rXX_launches.txt
  python scripts/summarize_ncu.py report gpurun_out/prof.ncu-rep > profiles/rXX_prof.txt
"""
import csv
import collections
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_uniform",
        "sm__sass_inst_executed_op_shared", "smsp__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        name = name.replace("void ", "")
        t = float(r[-1])
        a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a[0] += 1; a[1] += t
    total = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, {total/1e6:.3f} ms total (ncu serialised, cold cache: compare SHARES)")
    print(f"{'share':>7} {'ms':>10} {'count':>6}  kernel  [grid block]")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100*a[1]/total:6.2f}% {a[1]/1e6:10.3f} {a[0]:6d}  {name[:110]}  [{a[2]} {a[3]}]")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        print("## kernel:", rec.get("Kernel Name", "?")[:120])
        for h in hdr:
            if any(h.startswith(k) for k in KEYS) and rec[h] not in ("", "n/a"):
                print(f"{h} = {rec[h]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
