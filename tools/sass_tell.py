"""Per-kernel SASS evidence of the sm_100a-specific machinery (TMA loads / stores, mbarrier waits, bulk copies,
register reallocation), written to profiles/: `python tools/sass_tell.py [round tag]`.

cuobjdump -sass of the built library is split per kernel and the mnemonics that prove a feature are counted:
  UTMALDG / UTMASTG   cp.async.bulk.tensor (TMA tile load / store)        UBLKCP    cp.async.bulk (linear bulk copy)
  SYNCS               mbarrier arrive / try_wait / expect_tx              LDGSTS    cp.async (Ampere-style, not used)
  USETMAXREG          setmaxnreg (warp-specialised register budgets)      UTMAPF    TMA prefetch
  LDS.128 / LDS.64    vector shared-memory gathers                        REDG/ATOMG global reductions
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ppo_b200", "csrc", "libppo_b200.so")
TELLS = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "LDS.128", "LDS.64", "LDS", "STS",
         "LDG", "STG", "SHFL", "BAR", "NANOSLEEP", "DFMA", "FFMA"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for t in TELLS:
            if op == t or op.startswith(t + "."):
                counts[cur][t] += 1
    demangle = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(order, demangle if len(demangle) == len(order) else order))
    cols = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "USETMAXREG", "LDGSTS", "LDS.128", "LDS.64", "LDS", "SHFL", "DFMA"]
    out = [f"SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), {tag}",
           "LDS counts every shared-memory load (LDS.128 / LDS.64 included); instr = all SASS instructions of the kernel", "",
           f"{'kernel':<78}" + "".join(f"{c:>11}" for c in cols) + f"{'instr':>9}"]
    for fn in sorted(order, key=lambda f: names[f]):
        c = counts[fn]
        short = re.sub(r"\(.*", "", names[fn]).replace("ppo::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", short)
        out.append(f"{short[:77]:<78}" + "".join(f"{c[k]:>11}" for k in cols) + f"{c['_total']:>9}")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    out.append(f"{'TOTAL':<78}" + "".join(f"{tot[k]:>11}" for k in cols) + f"{tot['_total']:>9}")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_tell.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:3] + [l for l in out[3:] if "tvf_" in l or "TOTAL" in l or l.startswith("kernel")]))
    print("written", path)


if __name__ == "__main__":
    main()
