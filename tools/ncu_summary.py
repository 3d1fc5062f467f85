"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.txt
  python tools/ncu_summary.py raw gpurun_out/r01_gemm.ncu-rep profiles/r01_gemm_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list: {src}\n")
        f.write(f"# {sum(cnt.values())} launches, {T / 1e3:.3f} ms total device time (cold-cache, serialised: compare SHARES)\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v:12.1f} us {100 * v / T:6.2f}%  n={cnt[k]:5d}  {k}\n")
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h == "Kernel Name" or any(h.startswith(k) for k in KEYS)]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none capture: {src}\n")
        for r in rows[2:]:
            f.write("----\n")
            for i in idx:
                f.write(f"{hdr[i]} [{units[i]}] = {r[i]}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
