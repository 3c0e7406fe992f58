"""Print a fixed set of raw metrics from an ncu report (dev aid)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'dram__bytes_write.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg',
        'smsp__cycles_active.avg', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'l1tex__t_bytes.sum',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for data in rows[2:]:
        print("==", data[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
        for h, u, v in zip(hdr, units, data):
            if h in KEYS or (len(sys.argv) > 2 and sys.argv[2] in h):
                print(f"  {h:75s} {v:>18s} {u}")
if __name__ == "__main__":
    main()
