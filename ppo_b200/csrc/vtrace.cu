// V-trace off-policy targets and the trust-region weighting.
//
// Replaces rl/vtrace.py:44-114 `importance_sampling_v_trace` and rl/vtrace.py:3-39
// `v_trace_trust_region` of the reference (rho-bar = c-bar = 1 hard-coded there, :93-94).
//   rho_t = min(1, exp(log pi(a_t) - log mu(a_t))),  c_t = lamb * rho_t
//   acc_t = rho_t (r_t + g_t V_{t+1} - V_t) + g_t c_t acc_{t+1},   g_t = float32(gamma) * (1 - done_t),  acc_N = 0
//   vs_t  = acc_t + V_t
//   adv_t = rho_t (r_t + g_t vs_{t+1} - V_t),   vs_N = V_final
// Parallel structure: a CTA of 16 warps owns EW adjacent envs (EW = 32, 16 or 8) x 512 / EW time chunks; each thread
// gathers its chunk once (one 4-byte read per log-policy row: the gather is the only uncoalesced access, and all
// gathers of a chunk are issued before the first use), composes it into an affine map, folds the later chunks through
// shared memory and replays from registers.  Narrow env tiles exist for occupancy: at 4096 envs a 32-env tile gives 128
// CTAs for 148 SMs and 16 warps per SM in flight on a latency-bound gather; 8-env tiles give 512 CTAs, four resident
// per SM, every SM busy.  All arithmetic is float32 like the reference (:78).
#include <stdlib.h>

#include "common.cuh"

namespace ppo {

constexpr int kVtWarps = 16;
constexpr int kVtThreads = kVtWarps * 32;

struct VtraceParams {
    const float* beh;  // [N, B, ACT]
    const float* tgt;
    const void* actions;  // [N, B] int64 or int32
    int action_bytes;
    const float* rewards;
    const void* dones;
    int done_kind;
    const float* values;
    const float* final_value;
    int64_t steps, batch;
    int n_act;
    float gamma32, lamb32;
    float* vs;
    float* adv;
    float* cs;
    int chunk_len;
    int* bad_action;  // optional: set to 1 when an action lies outside [-n_act, n_act)
};

// NumPy index semantics (negative indices wrap once); anything else is clamped -- the launch never reads outside the
// log-policy rows -- and reported through the flag word (rl/vtrace.py:90 raises IndexError there).
__device__ __forceinline__ int64_t action_at(const void* a, int bytes, int64_t idx, int n_act, int* bad) {
    int64_t v = bytes == 8 ? static_cast<const int64_t*>(a)[idx] : static_cast<const int32_t*>(a)[idx];
    if (v < 0) v += n_act;
    if (v < 0 || v >= n_act) {
        if (bad) *bad = 1;
        v = v < 0 ? 0 : n_act - 1;
    }
    return v;
}

// LEN: steps of a chunk kept in registers (0: chunks of any length, re-read in the replay)
template <int LEN, int EW>
__global__ void __launch_bounds__(kVtThreads, LEN == 4 ? 4 : LEN == 8 ? 2 : 1) vtrace_kernel(const VtraceParams p) {
    constexpr bool CACHED = LEN > 0;
    constexpr int kLen = CACHED ? LEN : 1;
    constexpr int kSub = 32 / EW;              // chunks per warp
    constexpr int kChunks = kVtWarps * kSub;
    __shared__ float sB[kChunks * EW];
    __shared__ float sC[kChunks * EW];
    __shared__ float sV[(kChunks + 1) * EW];  // vs at the first step of each chunk

    const int lane = threadIdx.x & 31;
    const int e = lane % EW, c = (threadIdx.x >> 5) * kSub + lane / EW;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * EW + e;
    const bool live = env < p.batch;
    const int64_t lo = static_cast<int64_t>(c) * p.chunk_len;
    const int64_t hi = min(p.steps, lo + p.chunk_len);
    const int len = hi > lo ? static_cast<int>(hi - lo) : 0;

    // per-step quantities of this chunk
    float rho[kLen], gnt[kLen], r[kLen], v[kLen];
    float vtop = 0.f;
    if (live && len > 0) vtop = (hi == p.steps) ? p.final_value[env] : p.values[hi * p.batch + env];

    auto load_step = [&](int i, float& o_rho, float& o_gnt, float& o_r, float& o_v) {
        const int64_t idx = (lo + i) * p.batch + env;
        const int64_t a = action_at(p.actions, p.action_bytes, idx, p.n_act, p.bad_action);
        const float lt = __ldg(p.tgt + idx * p.n_act + a);
        const float lb = __ldg(p.beh + idx * p.n_act + a);
        o_rho = fminf(1.f, expf(__fsub_rn(lt, lb)));                       // rl/vtrace.py:93
        const float alive = alive_f32(p.dones, p.done_kind, idx);            // :78
        o_gnt = __fmul_rn(p.gamma32, alive);                              // (gamma * not_terminal[i]), :100
        o_r = __ldg(p.rewards + idx);
        o_v = __ldg(p.values + idx);
    };

    if (CACHED) {
#pragma unroll
        for (int i = 0; i < kLen; ++i) {
            rho[i] = 0.f, gnt[i] = 0.f, r[i] = 0.f, v[i] = 0.f;
            if (live && i < len) load_step(i, rho[i], gnt[i], r[i], v[i]);
        }
    }

    // tdv and the multiplier of the carried value, reference order (rl/vtrace.py:100-102)
    auto td_of = [&](float o_rho, float o_gnt, float o_r, float o_v, float vn) {
        return __fmul_rn(o_rho, __fsub_rn(__fadd_rn(o_r, __fmul_rn(o_gnt, vn)), o_v));
    };

    // ---- phase 1: compose
    float Bc = 0.f, Cc = 1.f;
    if (live) {
        if (CACHED) {
#pragma unroll
            for (int i = kLen - 1; i >= 0; --i) {
                if (i < len) {
                    const float vn = (i == len - 1) ? vtop : v[(i + 1) % kLen];
                    const float cc = __fmul_rn(gnt[i], __fmul_rn(p.lamb32, rho[i]));
                    Bc = fmaf(cc, Bc, td_of(rho[i], gnt[i], r[i], v[i], vn));
                    Cc *= cc;
                }
            }
        } else {
            float vn = vtop;
            for (int i = len - 1; i >= 0; --i) {
                float a, g, rr, vv;
                load_step(i, a, g, rr, vv);
                const float cc = __fmul_rn(g, __fmul_rn(p.lamb32, a));
                Bc = fmaf(cc, Bc, td_of(a, g, rr, vv, vn));
                Cc *= cc;
                vn = vv;
            }
        }
    }
    sB[c * EW + e] = Bc, sC[c * EW + e] = Cc;
    __syncthreads();

    // ---- phase 2: the later chunks folded into the value carried into this one
    float x = 0.f;
    for (int j = kChunks - 1; j > c; --j) x = fmaf(sC[j * EW + e], x, sB[j * EW + e]);

    // ---- phase 3: replay; vs of this chunk stays in v[] (CACHED) for the advantage pass
    float vs_first = 0.f;   // vs at step lo
    float top_rho = 0.f, top_gnt = 0.f, top_r = 0.f, top_v = 0.f;  // the last step of the chunk waits for vs of the next chunk
    if (live) {
        float vs_next = 0.f;
        if (CACHED) {
#pragma unroll
            for (int i = kLen - 1; i >= 0; --i) {
                if (i < len) {
                    const float vn = (i == len - 1) ? vtop : v[(i + 1) % kLen];
                    const float cs_i = __fmul_rn(p.lamb32, rho[i]);
                    const float td = td_of(rho[i], gnt[i], r[i], v[i], vn);
                    x = __fadd_rn(td, __fmul_rn(__fmul_rn(gnt[i], cs_i), x));          // :102
                    const float vs_i = __fadd_rn(x, v[i]);                              // :105
                    const int64_t idx = (lo + i) * p.batch + env;
                    store_stream(p.vs + idx, vs_i);
                    if (p.cs) store_stream(p.cs + idx, cs_i);
                    if (i == len - 1) {
                        top_rho = rho[i], top_gnt = gnt[i], top_r = r[i], top_v = v[i];
                    } else {
                        store_stream(p.adv + idx, td_of(rho[i], gnt[i], r[i], v[i], vs_next));  // :109
                    }
                    vs_next = vs_i;
                }
            }
        } else {
            float vn = vtop;
            for (int i = len - 1; i >= 0; --i) {
                float a, g, rr, vv;
                load_step(i, a, g, rr, vv);
                const float cs_i = __fmul_rn(p.lamb32, a);
                const float td = td_of(a, g, rr, vv, vn);
                x = __fadd_rn(td, __fmul_rn(__fmul_rn(g, cs_i), x));
                const float vs_i = __fadd_rn(x, vv);
                const int64_t idx = (lo + i) * p.batch + env;
                store_stream(p.vs + idx, vs_i);
                if (p.cs) store_stream(p.cs + idx, cs_i);
                if (i == len - 1) {
                    top_rho = a, top_gnt = g, top_r = rr, top_v = vv;
                } else {
                    store_stream(p.adv + idx, td_of(a, g, rr, vv, vs_next));
                }
                vs_next = vs_i;
                vn = vv;
            }
        }
        vs_first = vs_next;
    }
    sV[c * EW + e] = vs_first;
    __syncthreads();
    if (live && len > 0) {
        // vs following the last step of this chunk: first vs of the next non-empty chunk, or V_final (:106)
        const float vs_after = (hi == p.steps) ? p.final_value[env] : sV[(c + 1) * EW + e];
        store_stream(p.adv + (hi - 1) * p.batch + env, td_of(top_rho, top_gnt, top_r, top_v, vs_after));
    }
}

template <int EW>
void launch_vtrace(VtraceParams& p, cudaStream_t s) {
    constexpr int kChunks = kVtWarps * (32 / EW);
    p.chunk_len = static_cast<int>((p.steps + kChunks - 1) / kChunks);
    dim3 grid(static_cast<unsigned>((p.batch + EW - 1) / EW)), block(kVtThreads);
    if (p.chunk_len <= 4) vtrace_kernel<4, EW><<<grid, block, 0, s>>>(p);
    else if (p.chunk_len <= 8) vtrace_kernel<8, EW><<<grid, block, 0, s>>>(p);
    else if (p.chunk_len <= 16) vtrace_kernel<16, EW><<<grid, block, 0, s>>>(p);
    else vtrace_kernel<0, EW><<<grid, block, 0, s>>>(p);
}

// ------------------------------------------------------------------------------------------
// Trust region weight 1 / (1 + KL(pi || pi_hat)), pi_hat = min(mu, pi) renormalised (rl/vtrace.py:27-39).
struct TrustParams {
    const float* beh;
    const float* tgt;
    int64_t rows;
    int n_act;
    float* out;
};

__global__ void __launch_bounds__(256) trust_region_kernel(const TrustParams p) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= p.rows) return;
    const float* b = p.beh + row * p.n_act;
    const float* t = p.tgt + row * p.n_act;
    float z = 0.f;
    for (int a = 0; a < p.n_act; ++a) z += fminf(expf(__ldg(b + a)), expf(__ldg(t + a)));  // :31,:34
    float kl = 0.f;
    for (int a = 0; a < p.n_act; ++a) {
        const float pi = expf(__ldg(t + a));
        const float implied = __fdiv_rn(fminf(expf(__ldg(b + a)), pi), z);
        kl += logf(__fdiv_rn(pi, implied)) * pi;  // :37
    }
    p.out[row] = __fdiv_rn(1.f, 1.f + kl);  // :39
}

}  // namespace ppo

using namespace ppo;

extern "C" int ppo_vtrace(const float* behaviour_log_policy, const float* target_log_policy, const void* actions,
                          int action_bytes, const float* rewards, const void* dones, int done_kind,
                          const float* values, const float* final_value, int64_t steps, int64_t batch,
                          int64_t n_actions, double gamma, double lamb, float* vs_out, float* adv_out,
                          float* cs_out, int* bad_action, void* stream) {
    PPO_REQUIRE(steps >= 0 && batch >= 0 && n_actions >= 1 && n_actions < (1 << 24), "ppo_vtrace: bad shape");
    PPO_REQUIRE(action_bytes == 4 || action_bytes == 8, "ppo_vtrace: actions must be int32 or int64");
    PPO_REQUIRE(done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32 || done_kind == PPO_DONE_NONE,
                "ppo_vtrace: unknown done_kind %d", done_kind);
    if (steps == 0 || batch == 0) return PPO_OK;
    PPO_REQUIRE(behaviour_log_policy && target_log_policy && actions && rewards && values && final_value &&
                    vs_out && adv_out,
                "ppo_vtrace: null pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones, "ppo_vtrace: dones is null");
    VtraceParams p{};
    p.beh = behaviour_log_policy, p.tgt = target_log_policy;
    p.actions = actions, p.action_bytes = action_bytes;
    p.rewards = rewards, p.dones = dones, p.done_kind = done_kind;
    p.values = values, p.final_value = final_value;
    p.steps = steps, p.batch = batch, p.n_act = static_cast<int>(n_actions);
    p.gamma32 = static_cast<float>(gamma), p.lamb32 = static_cast<float>(lamb);
    p.vs = vs_out, p.adv = adv_out, p.cs = cs_out;
    p.bad_action = bad_action;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // env tile: the widest one that still gives every SM about two CTAs (narrow tiles only buy occupancy)
    static const int sms = sm_count();
    const char* ew_env = getenv("PPO_VTRACE_EW");  // development / tests: force the tile width
    const int forced = ew_env ? atoi(ew_env) : 0;
    (void)sms;
    int ew = 16;  // measured on B200 at 256 steps: 37 / 43 / 47 us (16 / 32 / 8 envs) at 4096 envs, 246 / 282 / 298 us at 32768
    if (forced == 8 || forced == 16 || forced == 32) ew = forced;
    if (ew == 32) launch_vtrace<32>(p, s);
    else if (ew == 16) launch_vtrace<16>(p, s);
    else launch_vtrace<8>(p, s);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_vtrace_trust_region(const float* behaviour_log_policy, const float* target_log_policy,
                                       int64_t rows, int64_t n_actions, float* out, void* stream) {
    PPO_REQUIRE(rows >= 0 && n_actions >= 1 && n_actions < (1 << 24), "ppo_vtrace_trust_region: bad shape");
    if (rows == 0) return PPO_OK;
    PPO_REQUIRE(behaviour_log_policy && target_log_policy && out, "ppo_vtrace_trust_region: null pointer");
    TrustParams p{};
    p.beh = behaviour_log_policy, p.tgt = target_log_policy, p.rows = rows, p.n_act = static_cast<int>(n_actions);
    p.out = out;
    trust_region_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
