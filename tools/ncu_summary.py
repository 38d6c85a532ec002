"""Key metrics (with units) of the kernels in an .ncu-rep, as text for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r01_ncu_x.txt
"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__warps_active.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised; use shares, not absolutes)")
    for v in rows[2:]:
        for n in WANT:
            if n in h:
                i = h.index(n)
                print(f"{n:90s} {v[i][:110]} {u[i]}")
        print()


if __name__ == "__main__":
    main()
