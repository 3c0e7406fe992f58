"""Offline estimate of shared-memory wavefronts per consumer instruction for a sample table (dev aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from ppo_b200 import _tvf_tables

def v_mult(n, lo, live):
    """wavefronts of one LDS.32 by 32 lanes: max over banks of distinct (row, col) addresses"""
    banks = {}
    for i in range(32):
        if live[i]:
            banks.setdefault(lo[i] % 32, set()).add((n[i], lo[i]))
    return max((len(v) for v in banks.values()), default=0)

def s_mult(n, live, tcp=68):
    """wavefronts of one LDS.128: per quarter-warp max over 8 bank-quads of distinct rows"""
    tot = 0
    for qd in range(4):
        quads = {}
        for i in range(qd * 8, qd * 8 + 8):
            if live[i]:
                quads.setdefault((n[i] * tcp // 4) % 8, set()).add(n[i])
        tot += max((len(v) for v in quads.values()), default=0)
    return tot

def analyse(tabs, label):
    E, K = tabs.ent_n.shape
    tv = ts = iv = is_ = 0
    for kb in range(K // 32):
        sl = slice(kb * 32, kb * 32 + 32)
        for e in range(E):
            n = tabs.ent_n[e, sl]; cnt = tabs.ent_cnt[e, sl]
            lo = tabs.ent_idx[e, sl] & 0xffff
            boot = (cnt > 0) & ((tabs.ent_wlo[e, sl] != 0) | (tabs.ent_whi[e, sl] != 0))
            if (cnt > 0).any():
                ts += s_mult(n, cnt > 0); is_ += 4
            if boot.any():
                tv += v_mult(n, lo, boot); iv += 1
    print(f"{label}: V wavefronts/LDS {tv / max(iv,1):.2f}  S wavefronts/LDS.128 {ts / max(is_,1) * 4:.2f} (ideal 4)")

if __name__ == "__main__":
    hz = oracle.get_value_head_horizons(128, 30000)
    for seed in (1, 2):
        table = oracle.draw_n_step_samples("exponential", "advanced", hz, 256, 20, 8, seed=seed)
        analyse(_tvf_tables.build(0.99997, 256, hz, hz, table), f"advanced8 seed {seed}")
    table = oracle.draw_n_step_samples("exponential", "standard", hz, 256, 20, 8, seed=1)
    analyse(_tvf_tables.build(0.99997, 256, hz, hz, table), "standard8")


def reassign(tabs, tcp=68, passes=3):
    """Greedy + local search slot permutation per lane; returns a new (ent_*) set."""
    from scipy.optimize import linear_sum_assignment
    E, K = tabs.ent_n.shape
    names = ["ent_n", "ent_cnt", "ent_idx", "ent_wlo", "ent_whi", "ent_g"]
    arrs = {a: getattr(tabs, a).copy() for a in names}
    for kb in range(K // 32):
        ks = list(range(kb * 32, kb * 32 + 32))
        # entries per lane
        ent = {k: [e for e in range(E) if tabs.ent_cnt[e, k] > 0] for k in ks}
        slot_of = {k: {} for k in ks}  # entry -> slot
        def key(k, e):
            n = int(tabs.ent_n[e, k]); lo = int(tabs.ent_idx[e, k] & 0xffff)
            boot = (tabs.ent_wlo[e, k] != 0) or (tabs.ent_whi[e, k] != 0)
            return n, lo, boot
        def cost_matrix(k):
            C = np.zeros((E, E))
            for ei, e in enumerate(ent[k]):
                n, lo, boot = key(k, e)
                for s in range(E):
                    cv = cs = 0
                    for k2 in ks:
                        if k2 == k:
                            continue
                        for e2, s2 in slot_of[k2].items():
                            if s2 != s:
                                continue
                            n2, lo2, boot2 = key(k2, e2)
                            if boot and boot2 and lo2 % 32 == lo % 32 and (n2, lo2) != (n, lo):
                                cv += 1
                            if (k2 - kb * 32) // 8 == (k - kb * 32) // 8 and n2 != n and (n2 * tcp // 4) % 8 == (n * tcp // 4) % 8:
                                cs += 1
                    C[ei, s] = 8 * cv * cv + cs * cs + 8 * cv + cs
            return C
        for it in range(passes):
            for k in ks:
                if not ent[k]:
                    continue
                slot_of[k] = {}
                C = cost_matrix(k)
                r, c = linear_sum_assignment(C[:len(ent[k])])
                slot_of[k] = {ent[k][ri]: int(ci) for ri, ci in zip(r, c)}
        for k in ks:
            for a in names:
                col = np.zeros(E, dtype=arrs[a].dtype)
                for e, s in slot_of[k].items():
                    col[s] = getattr(tabs, a)[e, k]
                arrs[a][:, k] = col
    class T: pass
    t = T()
    for a in names:
        setattr(t, a, arrs[a])
    return t

if __name__ == "__main__":
    hz = oracle.get_value_head_horizons(128, 30000)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, 256, 20, 8, seed=1)
    tabs = _tvf_tables.build(0.99997, 256, hz, hz, table)
    for passes in (1, 3):
        analyse(reassign(tabs, passes=passes), f"advanced8 seed 1 reassigned passes={passes}")
