"""Summarise an ncu report's SASS source page: hottest instructions by samples / executions (dev aid)."""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    body = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    S, E = ci["# Samples"], ci["Instructions Executed"]
    tot_s = sum(int(r[S] or 0) for r in body)
    tot_e = sum(int(r[E] or 0) for r in body)
    print(f"total samples {tot_s}, total warp-instructions {tot_e}, SASS lines {len(body)}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
    # stall columns in recent ncu are named like 'stall_long_sb' ... print names once
    names = [h for h in hdr if h.lower().startswith("stall") or h.lower().startswith("warp stall")]
    idx_lines = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:topn]
    print("---- top by samples (line#, samples, execs, sass, top stall reasons)")
    for i in sorted(idx_lines):
        r = body[i]
        reasons = []
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and not h.endswith("_not_issued"):
                v = int(r[j] or 0)
                if v:
                    reasons.append((v, h[6:]))
        reasons.sort(reverse=True)
        print(f"{i:5d} {int(r[S] or 0):6d} {int(r[E] or 0):9d}  {r[ci['Source']].strip()[:70]:70s} {reasons[:3]}")
    print("---- top by executions")
    for i in sorted(sorted(range(len(body)), key=lambda i: -int(body[i][E] or 0))[:15]):
        r = body[i]
        print(f"{i:5d} {int(r[S] or 0):6d} {int(r[E] or 0):9d}  {r[ci['Source']].strip()[:90]}")

if __name__ == "__main__":
    main()
