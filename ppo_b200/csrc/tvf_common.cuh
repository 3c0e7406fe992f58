// Pieces shared by the TVF target kernels (generic gather kernel and TMA ring kernel).
#pragma once

#include "common.cuh"

namespace ppo {

constexpr int kNoDone = 1 << 30;  // "no terminal before the end of the window"

struct TvfParams {
    const float* rewards;
    int64_t r_rs;
    const void* dones;
    int64_t d_rs;
    int done_kind;
    const float* vs;  // [N+1, A, V]
    int64_t vs_rs, vs_es;
    int steps, n_req, n_val;
    int64_t envs;
    ppo_tvf_tables tab;
    float* out;  // [N, A, K]
};

// Launch constants of the transposed slab kernel (tvf_tslab.cu), resolved on the host by ppo_tvf_tables_build.
constexpr int kTsMaxK = 128;
constexpr int kTsMaxEntries = 1024;
constexpr int kTsConstMagic = 0x54534c42;  // header word of the constants section: magic, n_ent (-1 = none), pitch, V
struct __align__(16) TsEntry {
    uint32_t off;    // byte offset of W^T[col][n - 1] in the slab
    float wlo, whi;  // multiplicity * gamma^n * lerp weights / samples per horizon
    int32_t n;
};
struct __align__(16) TsHorizon {
    int32_t beg, cnt;  // entries [beg, beg + cnt)
    float rscale;      // 1 for a live horizon, 0 for h = 0 (all-zero targets)
    int32_t maxn;      // largest n-step of the horizon (tail rows: N - t < maxn)
};

// Optional GAE + TD(lambda) stage fused into the transposed slab launch (host-side description).
struct TsGae {
    double gamma, lambda_adv, lambda_ret;
    float* adv;            // [N, A] contiguous
    float* ret;
    double* stats;         // PPO_STATS_LEN doubles
    double* partials;      // [max_partials][PPO_STATS_LEN]
    unsigned int* ticket;  // zero on entry, zero again on exit
    int64_t max_partials;
};

// Launch constants of the column slab kernel (tvf_cslab.cu), resolved on the host by ppo_tvf_tables_build.
// A consumer thread of that kernel owns aligned groups of four consecutive time rows and gathers with 16-byte loads,
// so an entry (horizon, n-step) is described by the aligned row group it reads ("bin" m = (n - 1) / 4: slab rows
// 4 (i + m) .. 4 (i + m) + 3 for the thread that owns rows 4 i .. 4 i + 3) and by how far the group is shifted against
// the thread's own rows (shift s = (n - 1) % 4: the values belong to rows 4 i - s .. 4 i - s + 3).  Entries of one
// horizon are sorted by shift (then n): the kernel runs one loop per shift, each with its accumulators fixed at
// compile time.
constexpr int kCsMaxK = 128;
constexpr int kCsMaxEntries = 1024;
constexpr int kCsGroups = 5;                // column groups of 32 value columns (at most four) + the -R column
constexpr int kCsSpecial = kCsGroups - 1;   // group index of the -R column
constexpr int kCsConstMagic = 0x43534c33;   // header: magic, n_ent (-1 = none), steps, V, jobs, arrivals[5], reserved
constexpr int kCsHeaderWords = 16;
constexpr uint32_t kCsOffMask = 0x3fff0u;   // byte offset (multiple of 16) of the first gathered group in the slab
struct __align__(16) CsEntry {
    uint32_t w;      // off | n << 18
    float wlo, whi;  // multiplicity * gamma^n * lerp weights / samples per horizon
    uint32_t pad;
};
struct __align__(16) CsHorizon {
    uint32_t w0;   // first entry | largest n-step << 11 | tail mode << 21 | live << 23
    uint32_t w1;   // entries with shift 0 | shift 1 << 8 | shift 2 << 16 | shift 3 << 24
    uint32_t w2;   // tail: first nibble of the horizon in the multiplicity table
    int32_t w3;    // tail mode 1: h_k - N (table index of row t is w3 + t); mode 2: value column s | (h_k - vh[s-1]) << 8
    float w4;      // tail mode 2: 1 / (vh[s] - vh[s-1])
    uint32_t w5, w6, w7;
};
constexpr int kCsTailNone = 0, kCsTailTable = 1, kCsTailLinear = 2, kCsTailGeneric = 3;

__device__ __forceinline__ bool done_at(const void* dones, int kind, int64_t idx) {
    if (kind == PPO_DONE_U8) return static_cast<const unsigned char*>(dones)[idx] != 0;
    if (kind == PPO_DONE_F32) return static_cast<const float*>(dones)[idx] != 0.f;
    return false;
}

// lerp of rl/returns_truncated.py:174 in float32: v_lo * (1 - f) + v_hi * f with both weights
// resolved on the host (zero weights encode "horizon <= 0 -> 0", :151-153).
__device__ __forceinline__ float lerp2(float v_lo, float v_hi, float w_lo, float w_hi) {
    return fmaf(v_hi, w_hi, v_lo * w_lo);
}

}  // namespace ppo
