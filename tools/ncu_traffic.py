"""Write profiles/tvf_traffic.json (DRAM bytes of one launch of the dominant kernel) from an ncu report (dev aid)."""
import csv, json, subprocess, sys

def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    def num(name):
        u, v = m[name]
        x = float(v.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "": 1, "%": 1}.get(u, 1)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    rec = {"report": rep, "kernel": m["Kernel Name"][1], "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "gpu_time_us_under_ncu": num("gpu__time_duration.sum"),
           "lsu_shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
           "lsu_shared_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
           "shared_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
           "note": "one ncu --set full capture of the named kernel at the bench shape (256 steps x 4096 envs x 128 "
                   "horizons, exponential / advanced 8 samples); writes still resident in L2 at kernel end are not in dram__bytes_write"}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))

if __name__ == "__main__":
    main()
