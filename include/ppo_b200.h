/*
 * ppo_b200 -- C ABI of the B200-native return / advantage hot path.
 *
 * Drop-in boundary for the reference trainer dremovd/PPO.  The reference has no FFI
 * layer: the path is a set of plain Python functions called by rl/rollout.py
 * (SURVEY.md section 8b).  Each entry point below replaces the NumPy body of one of
 * those functions; the Python package `ppo_b200` keeps the reference's names and
 * signatures and hands raw device pointers to these symbols through ctypes
 * (INTEGRATION.md shows the stub a maintainer adds to the reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all buffers; outputs are pre-allocated; nothing is freed or
 *     retained by the library (no global state besides a per-thread error string and
 *     the cached TMA driver entry point);
 *   - sizes and strides are int64 and counted in ELEMENTS;
 *   - scalar hyper-parameters cross the ABI as double (gamma must never be demoted to
 *     float before the reference does it, SURVEY.md section 7 "hard parts");
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     and return 0 or a negative ppo_status; nothing throws across the boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns
 *     PPO_ERR_CUDA.
 *
 * Layouts (row-major, env axis contiguous unless a stride argument says otherwise)
 *   rewards, dones, advantages, returns : [N, A]       N = steps, A = envs
 *   values                              : [N(+1), A]
 *   value_samples                       : [N+1, A, V]  V = value-sample horizons
 *   tvf targets                         : [N, A, K]    K = required horizons
 *   log-policies                        : [N, B, ACT]
 */
#ifndef PPO_B200_H
#define PPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PPO_OK = 0,
    PPO_ERR_BAD_ARG = -1,     /* shape / pointer / value rejected (Python: ValueError)      */
    PPO_ERR_UNSUPPORTED = -2, /* valid request this build cannot serve (Python: ValueError) */
    PPO_ERR_CUDA = -3,        /* CUDA runtime / driver error (Python: RuntimeError)         */
    PPO_ERR_WORKSPACE = -4    /* workspace missing or too small (Python: ValueError)        */
} ppo_status;

/* how the `dones` / terminal array is stored */
typedef enum {
    PPO_DONE_NONE = 0, /* no terminal array: rl/returns.py:23-24 `batch_terminal is None` */
    PPO_DONE_U8 = 1,   /* numpy bool / uint8, non-zero = terminal                         */
    PPO_DONE_F32 = 2   /* float32 0/1 mask (rl/unit_tests.py:205)                         */
} ppo_done_kind;

/* Layout of one statistics record (doubles).  Sums are over the float32-rounded outputs; the VAL_* slots are the
 * moments of the value estimates the scan read, all N + 1 rows (rl/rollout.py:1205 `*ext_value_estimates`). */
enum {
    PPO_STAT_COUNT = 0,
    PPO_STAT_ADV_SUM = 1,
    PPO_STAT_ADV_SUMSQ = 2,
    PPO_STAT_ADV_MIN = 3,
    PPO_STAT_ADV_MAX = 4,
    PPO_STAT_RET_SUM = 5,
    PPO_STAT_RET_SUMSQ = 6,
    PPO_STAT_RET_MIN = 7,
    PPO_STAT_RET_MAX = 8,
    PPO_STAT_VAL_SUM = 9,
    PPO_STAT_VAL_SUMSQ = 10,
    PPO_STAT_VAL_COUNT = 11,
    PPO_STAT_ADV_ZEROS = 12, /* number of exact zeros: the "sparsity" of rl/logger.py:201-229 is zeros / count */
    PPO_STAT_RET_ZEROS = 13,
    PPO_STATS_LEN = 16
};

/* Library version (major*10000 + minor*100 + patch) and the message of the last failing call
 * made by the calling thread. */
int ppo_version(void);
const char* ppo_last_error(void);

/* Number of visible CUDA devices, or a negative ppo_status. */
int ppo_device_count(void);

/* Bytes of zero-initialised device scratch the GAE entry points need for `envs` envs.
 * The buffer must be zeroed ONCE when allocated; the kernels leave it zeroed again. */
int64_t ppo_gae_workspace_bytes(int64_t envs);

/*
 * GAE advantages and TD(lambda) returns in one reverse-time scan.
 * Replaces rl/returns.py:7-29 `gae` and rl/returns.py:58-67 `td_lambda`, which the trainer calls
 * back to back on identical inputs with two lambdas (rl/rollout.py:1207-1223):
 *   delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t
 *   A_t     = delta_t + gamma * lamb * (1 - done_t) * A_{t+1},   A_N = 0,  V_N = final_value
 *   adv_out = A(lamb_adv),  ret_out = A(lamb_ret) + V
 * adv_out / ret_out may each be NULL.  `stats` (PPO_STATS_LEN doubles) may be NULL; when given it
 * receives count / sum / sum of squares / min / max of adv_out and ret_out over this call
 * (the "return-scale statistics" of rl/rollout.py:1252-1256 and the moments needed by the
 * normalisation of rl/rollout.py:1891).  Deterministic: partial sums are combined in a fixed order.
 * values: element (t, a) at values[t * value_row_stride + a * value_env_stride];
 * final_value: element a at final_value[a * final_env_stride].
 */
int ppo_gae(const float* rewards, int64_t reward_row_stride,
            const float* values, int64_t value_row_stride, int64_t value_env_stride,
            const float* final_value, int64_t final_env_stride,
            const void* dones, int done_kind, int64_t done_row_stride,
            int64_t steps, int64_t envs,
            double gamma, double lamb_adv, double lamb_ret,
            float* adv_out, float* ret_out, double* stats,
            void* workspace, int64_t workspace_bytes, void* stream);

#define PPO_MAX_PEERS 8

/*
 * ppo_gae followed by ppo_adv_normalize on the raw advantages (single GPU: no moment exchange in between), as ONE
 * cooperative launch when the whole grid of the tile kernel is resident: every CTA keeps its advantages in shared
 * memory, waits for the folded statistics and writes the normalised values itself.  Returns 1 when fused, 0 when it
 * fell back to the two launches, < 0 on error.  stats, adv_out and norm_out are required; workspace as for ppo_gae
 * (zero-initialised, left zero).
 */
int ppo_gae_normalized(const float* rewards, int64_t reward_row_stride,
                       const float* values, int64_t value_row_stride, int64_t value_env_stride,
                       const float* final_value, int64_t final_env_stride,
                       const void* dones, int done_kind, int64_t done_row_stride,
                       int64_t steps, int64_t envs,
                       double gamma, double lamb_adv, double lamb_ret, double eps, double clip,
                       float* adv_out, float* ret_out, float* norm_out, double* stats,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* 1 when ppo_gae_normalized / ppo_gae_normalized_peers would take the single cooperative launch for this shard shape
 * on the current device (nothing is launched): ranks agree on it (all-reduce MIN) before relying on the peer exchange. */
int ppo_gae_normalized_fits(int64_t steps, int64_t envs, int done_kind, int64_t value_env_stride);

/*
 * Env-sharded form of ppo_gae_normalized (one process per GPU, envs split over `world` GPUs of one NVLink domain): the
 * exchange of the advantage moments -- the only cross-GPU step of the path, SURVEY 8e -- happens INSIDE the launch
 * over peer memory instead of an NCCL all-reduce between two launches.  Every rank owns a mailbox
 * [2][world][PPO_STATS_LEN] doubles and 2 x `world` 64-bit flag words (two halves, used alternately by the parity of
 * the epoch so that a rank one launch ahead never overwrites a record a slower rank still reads), mapped into all
 * peers (CUDA IPC); peer_mail[r] / peer_flags[r] are rank r's buffers as seen from this process (host arrays of `world`
 * device pointers).  The CTA that folds the local statistics stores the record into slot `rank` of every mailbox and
 * release-stores `epoch` into flag `rank` there; all CTAs then wait until their own `world` flags have reached `epoch`,
 * add the moments in rank order (identical bits on every rank) and normalise their tile.  epoch must increase by one
 * with every launch on every rank (flags are never reset; zero them once at start-up).  stats receives the local
 * record, global_stats (optional) the merged one.
 * Returns 1 when launched, 0 when the shard shape does not fit the resident tile kernel (nothing launched: every rank
 * decides alike; use ppo_gae + all-reduce + ppo_adv_normalize), < 0 on error.
 * A peer that does not arrive within timeout_ms (<= 0: 4000) makes the kernel give up instead of hanging: it stores
 * `epoch` into *status (optional; device memory or mapped pinned host memory, zero it at start-up and read it after
 * the stream has drained -- non-zero means the normalised advantages of that launch are NaN and the pass must be
 * repeated, e.g. with the all-reduce path, after the ranks have met at a host barrier).
 */
int ppo_gae_normalized_peers(const float* rewards, int64_t reward_row_stride,
                             const float* values, int64_t value_row_stride, int64_t value_env_stride,
                             const float* final_value, int64_t final_env_stride,
                             const void* dones, int done_kind, int64_t done_row_stride,
                             int64_t steps, int64_t envs,
                             double gamma, double lamb_adv, double lamb_ret, double eps, double clip,
                             float* adv_out, float* ret_out, float* norm_out, double* stats, double* global_stats,
                             void* workspace, int64_t workspace_bytes, int rank, int world,
                             double* const* peer_mail, unsigned long long* const* peer_flags,
                             unsigned long long epoch, unsigned long long* status, double timeout_ms, void* stream);

/*
 * Batch advantage normalisation: out = clip((x - mean) / (std + eps), -clip, +clip), population std.
 * Replaces the inline NumPy at rl/rollout.py:1887-1897.  `stats` is a device record as written by
 * ppo_gae (after an optional cross-GPU sum of COUNT / ADV_SUM / ADV_SUMSQ).  clip <= 0 disables
 * clipping (rl/rollout.py:1895 `advantage_clipping is None`).  In-place (out == adv) is allowed.
 */
int ppo_adv_normalize(const float* adv, int64_t count, const double* stats,
                      double eps, double clip, float* out, void* stream);

/*
 * Extrinsic + intrinsic advantage, rl/rollout.py:1225-1231: out = ext_adv + float32(scale) * int_adv (out may alias
 * ext_adv), and a statistics record of the result: the ADV_* slots hold count / sum / sumsq / min / max / zeros of the
 * sum (feed it to ppo_adv_normalize, after the cross-GPU exchange when sharded), the RET_* slots the same of the
 * scaled intrinsic part (`adv_int` and, with the extrinsic record of ppo_gae, `adv_ratio` of the reference's log).
 * workspace: ppo_gae_workspace_bytes(0) bytes, zeroed once (left reusable).
 */
int ppo_adv_combine(const float* ext_adv, const float* int_adv, double scale, int64_t count, float* out,
                    double* stats, void* workspace, int64_t workspace_bytes, void* stream);

/* Statistics records [n_records][PPO_STATS_LEN] of disjoint env blocks or ranks -> one record: sums are added in
 * record order, min / max are reduced (out may alias records[0]).  Used for the env-blocked host pipeline and, after an
 * all-gather of the per-rank records, for the cross-GPU return statistics. */
int ppo_merge_stats(const double* records, int64_t n_records, double* out, void* stream);

/* count / sum / sumsq / min / max of a float32 vector into the ADV_* slots of `stats`
 * (for advantages that were combined on the device, e.g. ext + int, rl/rollout.py:1225-1228). */
int ppo_moments(const float* x, int64_t count, double* stats,
                void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Discounted returns bootstrapped from a final value.  Replaces rl/returns.py:32-55
 * `calculate_bootstrapped_returns`:  R_t = r_t + R_{t+1} * gamma_t * (1 - done_t),  R_N = final_value.
 * gamma_array == NULL: scalar gamma, applied as float32(gamma) like the reference (rl/returns.py:49-50);
 * otherwise gamma_array is [N, A] float32 with the given row stride and `gamma` is ignored.
 */
int ppo_discounted_returns(const float* rewards, int64_t reward_row_stride,
                           const void* dones, int done_kind, int64_t done_row_stride,
                           const float* final_value, int64_t final_env_stride,
                           double gamma, const float* gamma_array, int64_t gamma_row_stride,
                           int64_t steps, int64_t envs, float* out, void* stream);

/*
 * Index tables of one TVF target pass, resolved on the host from the [K, C] n-step sample table
 * (ppo_b200/_tvf_tables.py documents the construction; it touches no rollout data).  All pointers
 * are device pointers; entry arrays are [n_entries, K], tail arrays are [nmax + 1, K].
 *   ent_n    clamped n-step n' = min(sample, h_k) of the entry, 0 = padding
 *   ent_cnt  multiplicity of n' among the C samples of horizon k (0 = padding)
 *   ent_idx  interpolation slots at horizon h_k - n':  lo | hi << 16
 *   ent_wlo / ent_whi   lerp weights (both 0 when h_k - n' <= 0: pure n-step return)
 *   ent_g    gamma^n' as the reference's float32 running product
 *   tail_*   the same slots / weights at horizon h_k - m for rows with m = N - t < n'
 *   tail_cwlo / tail_cwhi   tail weights folded with gamma^m and the summed multiplicity of the
 *            entries with n' > m (so the run-off-the-end bootstrap is one fused term per target)
 *   gpow     gamma^n, n = 0..nmax
 *   inv_c    float32(1 / C)
 *   tailT_*  (optional) horizon-major copies [K, tail_pitch] of tail_idx / tail_cwlo / tail_cwhi
 *   host_blob / dev_blob  (optional) when every table above lives in ONE packed device allocation starting at
 *            dev_blob and host_blob is a HOST copy of the same bytes (both as written by
 *            ppo_tvf_tables_build), ppo_tvf_targets may resolve per-horizon launch constants on the host
 *            without a device round trip; required by the transposed slab kernel (PPO_TVF_TSLAB).  host_blob
 *            only has to stay valid for the duration of the call.
 *   gamma    (optional, 0 = absent) the float64 discount the tables were built for; the transposed slab kernel
 *            carries the discounted reward-to-go in float64 with it
 *   tslab_offset_words  (optional, 0 = absent) offset inside the blob of the per-(horizon, n-step) launch
 *            constants of the transposed slab kernel (ppo_tvf_table_layout.offset_words[15]); they are copied from
 *            host_blob into the kernel parameters, so a launch does no table arithmetic
 */
typedef struct {
    int32_t n_entries;
    int32_t nmax;
    const int32_t* ent_n;
    const float* ent_cnt;
    const int32_t* ent_idx;
    const float* ent_wlo;
    const float* ent_whi;
    const float* ent_g;
    const int32_t* tail_idx;
    const float* tail_wlo;
    const float* tail_whi;
    const float* tail_cwlo;
    const float* tail_cwhi;
    const float* gpow;
    float inv_c;
    int32_t tail_pitch;
    const int32_t* tailT_idx;
    const float* tailT_cwlo;
    const float* tailT_cwhi;
    const void* host_blob;
    const void* dev_blob;
    double gamma;
    int64_t tslab_offset_words;
    int64_t cslab_offset_words; /* (optional, 0 = absent) the same for the column slab kernel
                                   (ppo_tvf_table_layout.cslab_offset_words) */
} ppo_tvf_tables;

/* Offsets (in 32-bit words) of the tables inside the packed blob written by ppo_tvf_tables_build: [0..11] in the
 * order of the first twelve pointer members of ppo_tvf_tables (ent_n ... gpow), [12..14] tailT_idx / tailT_cwlo /
 * tailT_cwhi, [15] launch constants of the transposed slab kernel (header, horizon records, entries, then the tail
 * tables once more as interleaved 16-byte records); every table starts 16-byte aligned. */
typedef struct {
    int32_t n_entries;
    int32_t nmax;
    int64_t total_words;
    int64_t offset_words[16];
    float inv_c;
    int32_t tail_pitch;
    double gamma;
    int64_t cslab_offset_words; /* launch constants of the column slab kernel (header, horizon words, job words,
                                   entries in ascending n-step order with their aligned row groups) */
    int32_t sections;           /* PPO_TVF_TABLES_* bits of what was actually written (see ppo_tvf_tables_build_ex) */
    int32_t reserved;
} ppo_tvf_table_layout;

/* Sections of the blob.  GENERIC: tables [0..14] (entry / tail tables of the generic and ring kernels); TSLAB /
 * CSLAB: launch constants of the transposed / column slab kernels. */
enum {
    PPO_TVF_TABLES_GENERIC = 1,
    PPO_TVF_TABLES_TSLAB = 2,
    PPO_TVF_TABLES_CSLAB = 4,
    PPO_TVF_TABLES_ALL = 7
};

/*
 * HOST-side construction of the index tables (no device work, no rollout data): resolves a [K, C] n-step sample
 * table exactly like rl/returns_truncated.py:590-591 (clamp to the horizon), :142-174 (searchsorted slot and lerp
 * weights, float64 arithmetic), :565,:573-574 (log-interpolation axis) and :669 (float32 discount powers).
 * All pointers are HOST pointers; `blob` receives the packed tables (copy it to the device and add its device
 * address to the offsets to fill ppo_tvf_tables).  ppo_tvf_tables_words gives a sufficient blob size.
 * Returns PPO_ERR_BAD_ARG when a clamped n-step is outside
 * 1..N (the reference's assert, :597) and PPO_ERR_UNSUPPORTED when a bootstrap horizon lies beyond the last value
 * horizon (the reference's IndexError, :157).
 */
int64_t ppo_tvf_tables_words(int64_t n_req, int64_t n_samples, int64_t steps);
int ppo_tvf_tables_build(double gamma, int64_t steps, const int64_t* required_horizons, int64_t n_req,
                         const int64_t* value_horizons, int64_t n_val, const int64_t* samples, int64_t n_samples,
                         int use_log_interpolation, int32_t* blob, int64_t blob_words,
                         ppo_tvf_table_layout* layout);
/* The same with a choice of sections, for a caller that rebuilds the tables every rollout and knows which kernel
 * runs (the build time is proportional to what is written).  A request without PPO_TVF_TABLES_GENERIC is honoured only
 * if the transposed slab kernel can serve the table (>= 3 distinct n-steps per horizon, <= 1024 entries, adjacent
 * interpolation slots, 32 / 64 / 128 value horizons, K % 32 == 0, N <= 512): otherwise the generic tables are written
 * as well, so that some kernel can always run.  layout->sections reports what was written; leave the ppo_tvf_tables
 * pointers of an unwritten section NULL (ppo_tvf_targets then refuses the kernels that need them instead of reading
 * stale memory).  Blob offsets and size do not depend on the sections. */
int ppo_tvf_tables_build_ex(double gamma, int64_t steps, const int64_t* required_horizons, int64_t n_req,
                            const int64_t* value_horizons, int64_t n_val, const int64_t* samples, int64_t n_samples,
                            int use_log_interpolation, int sections, int32_t* blob, int64_t blob_words,
                            ppo_tvf_table_layout* layout);

/* which TVF kernel ppo_tvf_targets may use */
typedef enum {
    PPO_TVF_AUTO = 0,    /* >= 3 n-steps per horizon: the quad slab kernel at 256 steps x 128 horizons and up to 4096
                            envs per launch (PPO_TVF_QSLAB=0 / 1 in the environment forces the choice), else the
                            transposed slab kernel; fewer: the TMA ring kernel; the generic kernel when the shape
                            allows none of them */
    PPO_TVF_GENERIC = 1, /* global-memory gather kernel, any shape                            */
    PPO_TVF_RING = 2,    /* TMA ring kernel or PPO_ERR_UNSUPPORTED                            */
    PPO_TVF_TSLAB = 3,   /* transposed slab kernel or PPO_ERR_UNSUPPORTED                     */
    PPO_TVF_CSLAB = 4,   /* column slab kernel (128 or 256 steps, <= 128 value horizons in multiples of 32, horizon
                            count a multiple of 4) or PPO_ERR_UNSUPPORTED */
    PPO_TVF_QSLAB = 5    /* quad slab kernel (128 or 256 steps, 128 required horizons on 128 value horizons) or
                            PPO_ERR_UNSUPPORTED */
} ppo_tvf_path;

/*
 * Multi-horizon truncated-return targets [N, A, K].  Replaces rl/returns_truncated.py:623-693
 * `_calculate_sampled_return_multi_fast` with its helpers `_n_step_estimate` (:558-620) and
 * `_interpolate` (:142-174); the dispatcher `get_return_estimate` (:10-139) stays in Python and
 * draws the sample table with the reference's own NumPy RNG calls.
 *   out[t,a,k] = 1/C * sum_c ( S_n(t,a) + D_n(t,a) * I(V[t+n,a,:], h_k - n) ),  n = min(sample[k,c], h_k, N - t)
 * value_samples is [N+1, A, V] with a unit-stride V axis; element (t, a, j) at
 * value_samples[t * vs_row_stride + a * vs_env_stride + j].  dones must be a 0/1 mask.
 * `stats` (PPO_STATS_LEN doubles, may be NULL) receives count / sum / sumsq / min / max of the
 * longest horizon out[:, :, K-1] in the ADV_* slots (rl/rollout.py:1270 `tvf_return_ext`); it needs
 * `workspace` as sized by ppo_tvf_workspace_bytes.
 */
int64_t ppo_tvf_workspace_bytes(int64_t steps, int64_t envs);
int ppo_tvf_targets(const float* rewards, int64_t reward_row_stride,
                    const void* dones, int done_kind, int64_t done_row_stride,
                    const float* value_samples, int64_t vs_row_stride, int64_t vs_env_stride,
                    int64_t steps, int64_t envs, int64_t n_req, int64_t n_val,
                    const ppo_tvf_tables* tables, int path,
                    float* out, double* stats, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * ppo_tvf_targets plus the value estimate the GAE stage reads off the same value curves
 * (rl/rollout.py:1196 -> rl/tvf.py:303-360 `get_tvf_ext_value_estimate`):
 *     value_out[r, a] = sum_j value_weights[j] * value_samples[r, a, j],   r = 0..N
 * value_weights is a HOST array of n_val floats: w_0 = -rho_1, w_j = rho_j - rho_{j+1}, w_{V-1} = rho_{V-1} with the
 * clipped rediscount ratios rho of rl/tvf.py:424-431 gives `get_rediscounted_value_estimate` (rl/tvf.py:388-433);
 * NULL selects the last value horizon (`tvf_value[:, :, -1]`, rl/tvf.py:325-340).  On the transposed slab kernel the
 * estimate is produced by the same launch from the value rows it streams anyway; otherwise one more kernel runs.
 * value_out is [N+1, A] with row stride value_out_row_stride (0 = envs).  extras == NULL: plain ppo_tvf_targets.
 */
typedef struct {
    const float* value_weights;
    float* value_out;
    int64_t value_out_row_stride;
    /* Optional: GAE + TD(lambda) over that value estimate in the SAME launch (rl/rollout.py:1207-1223 ->
     * rl/returns.py:7-29 `gae`, :58-67 `td_lambda`), transposed slab kernel only.  gae_adv_out / gae_ret_out are
     * contiguous [N, A]; gae_stats receives the PPO_STATS_LEN record of ppo_gae; gae_workspace is sized by
     * ppo_gae_workspace_bytes(envs).  On return gae_done is 1 when the launch produced them and 0 when the caller has
     * to run ppo_gae itself (another kernel was taken, or value_out is missing).  gae_enable == 0: ignored. */
    int32_t gae_enable;
    int32_t gae_done;
    double gae_gamma, gae_lambda_adv, gae_lambda_ret;
    float* gae_adv_out;
    float* gae_ret_out;
    double* gae_stats;
    void* gae_workspace;
    int64_t gae_workspace_bytes;
} ppo_tvf_extras;

int ppo_tvf_targets_ex(const float* rewards, int64_t reward_row_stride,
                       const void* dones, int done_kind, int64_t done_row_stride,
                       const float* value_samples, int64_t vs_row_stride, int64_t vs_env_stride,
                       int64_t steps, int64_t envs, int64_t n_req, int64_t n_val,
                       const ppo_tvf_tables* tables, int path,
                       float* out, double* stats, void* workspace, int64_t workspace_bytes,
                       ppo_tvf_extras* extras, void* stream);

/* Standalone forms.  ppo_value_estimate: out[b] = sum_j host_weights[j] * values[b * row_stride + j] (replaces
 * rl/tvf.py:388-433 with the weights above).  ppo_rediscount_horizons: out[b, k] = sum_{i=1..k} (V[b,i] - V[b,i-1]) *
 * host_ratios[i] (replaces rl/tvf.py:362-380 for old_gamma != new_gamma; host_ratios[0] is ignored). */
int ppo_value_estimate(const float* values, int64_t rows, int64_t row_stride, int64_t n_val,
                       const float* host_weights, float* out, void* stream);
int ppo_rediscount_horizons(const float* values, int64_t rows, int64_t row_stride, int64_t n_val,
                            const float* host_ratios, float* out, int64_t out_row_stride, void* stream);

/* Per-row interpolation of value curves at per-row target horizons.  Replaces rl/tvf.py:527-573 `horizon_interpolate`
 * (float64 result like the reference; targets are clipped to the axis, targets at or below the first horizon give 0).
 * host_horizons: HOST [n_val] strictly ascending axis (any scale, e.g. the log scale of rl/tvf.py:143-152);
 * values [rows, n_val] float32 and targets / out [rows] float64 are device arrays. */
int ppo_horizon_interpolate(const double* host_horizons, int64_t n_val, const float* values, int64_t rows,
                            int64_t row_stride, const double* targets, double* out, void* stream);

/*
 * Horizon trimming of one rollout step.  Replaces the arithmetic of rl/tvf.py:91-208 `TVFRunnerModule.trim_horizons`
 * (modes 0 interpolate, 1 average, 2 substitute, 3 random).  values [A, K, VH] float32 (device, contiguous).  The host
 * resolves every index with the reference's own NumPy calls: trim_index [A] = searchsorted(horizons, time left)
 * (:137); for mode 0 the slot (pre, post, frac) of `horizon_interpolate` on the log-scaled axis (:143-152); for mode 3
 * picks [A, K], the source horizon drawn for every horizon (:179-189).  out is [A, K, VH] ([A, K, 1] for mode 2,
 * :172-178); final_value [A] is the mean of the valid untrimmed ext estimates (:193-196); clip < 0 disables the
 * clipping of the change (:199-201).
 */
int ppo_tvf_trim_horizons(const float* values, int64_t envs, int64_t n_val, int64_t n_heads, int mode,
                          const int32_t* trim_index, const int32_t* interp_pre, const int32_t* interp_post,
                          const double* interp_frac, const int32_t* picks, double clip, float* out,
                          float* final_value, void* stream);

/* Development aid: with the environment variable PPO_TVF_DEBUG bit 64 set, the transposed slab kernel records
 * globaltimer stamps (ns) of its first CTA at ten hand-off points (start, shared memory initialised, first R / lim
 * tables, first chunk / first env in the slab, first step starts, first env done, last step starts, consumers done,
 * last store done); this call synchronises the device and copies the 16 slots of the last such launch. */
int ppo_tvf_debug_times(unsigned long long* out16);

/* The kernel the last successful ppo_tvf_targets call of this thread launched: a ppo_tvf_path value (never AUTO),
 * plus 0x100 when the statistics of the longest horizon were fused into that launch (no second kernel) and 0x200
 * when the value estimate of ppo_tvf_targets_ex was. */
int ppo_tvf_last_path(void);

/* count / sum / sumsq / min / max of the longest horizon targets[:, K-1] of a [rows, K] target buffer into the
 * ADV_* slots of `stats` (rl/rollout.py:1270 `tvf_return_ext`); what ppo_tvf_targets runs when `stats` is given. */
int ppo_tvf_last_horizon_stats(const float* targets, int64_t rows, int64_t n_req, double* stats,
                               void* workspace, int64_t workspace_bytes, void* stream);

/*
 * V-trace value targets, rho-weighted advantages and c weights.  Replaces rl/vtrace.py:44-114
 * `importance_sampling_v_trace` (rho-bar = c-bar = 1 as hard-coded there):
 *   rho_t = min(1, exp(log pi(a_t) - log mu(a_t))),  c_t = lamb * rho_t,  g_t = gamma * (1 - done_t)
 *   acc_t = rho_t (r_t + g_t V_{t+1} - V_t) + g_t c_t acc_{t+1};  vs = acc + V;  adv_t = rho_t (r_t + g_t vs_{t+1} - V_t)
 * log-policies are contiguous [N, B, ACT]; actions [N, B] are int32 (action_bytes 4) or int64 (8) and may be
 * negative with NumPy semantics; every other array is contiguous [N, B] / [B].  cs_out may be NULL.
 * bad_action (optional device word, zero it before the launch): set to 1 when some action lies outside
 * [-ACT, ACT) -- the reference raises IndexError there (rl/vtrace.py:90).  Such an index is clamped to the nearest
 * valid action, so the launch never reads outside the log-policy rows whether or not the word is given.
 */
int ppo_vtrace(const float* behaviour_log_policy, const float* target_log_policy,
               const void* actions, int action_bytes, const float* rewards,
               const void* dones, int done_kind, const float* values, const float* final_value,
               int64_t steps, int64_t batch, int64_t n_actions, double gamma, double lamb,
               float* vs_out, float* adv_out, float* cs_out, int* bad_action, void* stream);

/* Trust-region weight 1 / (1 + KL(pi || pi_hat)) per row, pi_hat = min(mu, pi) renormalised over the
 * actions.  Replaces rl/vtrace.py:3-39 `v_trace_trust_region`; rows = N * B, contiguous [rows, ACT]. */
int ppo_vtrace_trust_region(const float* behaviour_log_policy, const float* target_log_policy,
                            int64_t rows, int64_t n_actions, float* out, void* stream);

/*
 * Forward return trackers and running moments: intrinsic-return normalisation (rl/rollout.py:1147-1165,
 * Runner.calculate_intrinsic_returns) and reward normalisation (rl/wrappers.py:795-885, VecNormalizeRewardWrapper.step)
 * over a whole [N, A] rollout.  Both run  x_t = r_t + gamma * (1 - done_t) * x_{t-1}  forwards in time with a float64
 * carry per env that survives between rollouts (`ems_norm`, rl/rollout.py:280; `current_returns`, rl/wrappers.py:826),
 * and after every step merge the batch mean / variance of x_t over the envs into a running (mean, var, count)
 * (rl/utils.py:379-443 RunningMeanStd.update).
 *
 * ppo_return_tracker_scan: carry [A] float64 in/out; tracker_out [N, A] float64 (row t = x_t).  done_kind NONE is
 *   `ir.propagation` (rl/rollout.py:1151: terminals are ignored).
 * ppo_row_sums: out[t] = sum_a x[t, a]  (row_sums == NULL)  or  sum_a (x[t, a] - row_sums[t] / total_envs)^2, the two
 *   passes of np.mean / np.var; env-sharded callers all-reduce `out` (sum) between the two calls.
 * ppo_rms_fold: the N merges in step order.  rms = {mean, var, count} in/out (device); row_var [N] (optional) receives
 *   the running variance after each step.
 * ppo_scale_rewards: out[t, a] = clip((r[t, a] - shift) / s, +-clip) * scale with
 *   PPO_SCALE_INTRINSIC  s = eps + sqrt(var)   (rl/rollout.py:1164-1165, eps = 1e-5), var = var[0] or var[t] (per_row)
 *   PPO_SCALE_WRAPPER    s = sqrt(var + eps)   (rl/wrappers.py:893-895, eps = 1e-2)
 *   clip < 0 disables clipping (rl/wrappers.py:873); center_stats != NULL: shift = mean of the raw rewards from a
 *   ppo_moments record (`ir.center`, rl/rollout.py:1167-1168), else 0; clips (optional) is incremented by the number of
 *   clipped elements (rl/wrappers.py:876).
 */
typedef enum { PPO_SCALE_INTRINSIC = 0, PPO_SCALE_WRAPPER = 1 } ppo_scale_mode;

int ppo_return_tracker_scan(const float* rewards, int64_t reward_row_stride, const void* dones, int done_kind,
                            int64_t done_row_stride, int64_t steps, int64_t envs, double gamma, double* carry,
                            double* tracker_out, void* stream);
int ppo_row_sums(const double* x, int64_t steps, int64_t envs, const double* row_sums, double total_envs, double* out,
                 void* stream);
int ppo_rms_fold(const double* row_sums, const double* row_sq, int64_t steps, double total_envs, double* rms,
                 double* row_var, void* stream);
/* The same merges plus the wrapper's mode "ema" (rl/wrappers.py:878-881): after every step
 * alpha = 1 - total_envs / min(count, ema_horizon), *ema_var = alpha * *ema_var + (1 - alpha) * var_over_envs(x_t);
 * row_var (optional) then records the EMA variance after each step. */
int ppo_rms_fold_ema(const double* row_sums, const double* row_sq, int64_t steps, double total_envs, double* rms,
                     double* ema_var, double ema_horizon, double* row_var, void* stream);
/* One GPU, no exchange: ppo_row_sums twice + ppo_rms_fold(_ema) in ONE launch (a CTA per row makes both passes, the
 * last CTA to finish runs the merges); bit-identical to the three launches.  row_sums / row_sq [steps] are scratch
 * outputs, ema_var may be NULL, ticket is a zeroed device word (left zero). */
int ppo_row_moments_fold(const double* x, int64_t steps, int64_t envs, double* row_sums, double* row_sq, double* rms,
                         double* ema_var, double ema_horizon, double* row_var, unsigned int* ticket, void* stream);
/* ReturnTracker.update of rl/wrappers.py:737-755 (the multi-env reward normaliser, :946-971) over the N steps of a
 * rollout: only the envs with mask[t, a] != 0 move (float32 state `carry` [A], rounded after every step like the
 * reference's in-place NumPy assignment), and the running moments rms = {mean, var, count} take the float32 batch mean /
 * variance of the masked envs of every step in order (a step without masked envs changes nothing, :750-751).
 * tracker_out [N, A] f64 (scratch: the state after each step), row_records [3 N] f64 (scratch: count, mean, variance per
 * step), row_var [N] f64 (optional: the running variance after each step -- what step t's masked rewards are divided
 * by), ticket: a zeroed device word (left zero).  Two launches. */
int ppo_return_tracker_masked(const float* rewards, int64_t reward_row_stride, const void* dones, int done_kind,
                              int64_t done_row_stride, const unsigned char* mask, int64_t mask_row_stride, int64_t steps,
                              int64_t envs, double gamma, float* carry, double* tracker_out, double* row_records,
                              double* rms, double* row_var, unsigned int* ticket, void* stream);
int ppo_scale_rewards(const float* rewards, int64_t reward_row_stride, int64_t steps, int64_t envs, const double* var,
                      int per_row, int mode, double eps, double clip, double scale, const double* center_stats,
                      float* out, unsigned long long* clips, void* stream);

/*
 * Consumers of device-resident targets.
 *
 * ppo_gather_rows: out[i, :] = src[index[i], :] for byte rows (the minibatch selection `var_value[sample]` of
 *   rl/rollout.py:2349-2372 on buffers that never left the device).  index is int64 [n_index], negative values count
 *   from the end like NumPy; an index outside [-n_src_rows, n_src_rows) zero-fills the row and sets *out_of_range = 1
 *   (the caller raises IndexError).  Pitches are in bytes.
 * ppo_tvf_value_loss: TVFRunnerModule.on_train_value_minibatch (rl/tvf.py:32-77), forward and gradient in one pass:
 *     loss[b]    = sqrt(K) * mean_k( coef * 0.5 * (targets[b, head_index[k]] - pred[b, k])^2 * head_weights[k] * mask[b, k] )
 *     grad[b, k] = d loss[b] / d pred[b, k]
 *   head_index (optional, int32 [K]) is `required_tvf_heads` (:45-46); head_weights [K] carries the duplicate-head
 *   weights and the h-weighting (:52-62) folded on the host; mask (optional, [rows, K], contiguous) is the horizon
 *   dropout mask (:64-69, drawn by the caller).  predictions are strided: [rows, K, VH] with VH heads -> pred_head_stride.
 */
int ppo_gather_rows(const void* src, int64_t src_pitch_bytes, int64_t row_bytes, int64_t n_src_rows,
                    const int64_t* index, int64_t n_index, void* out, int64_t out_pitch_bytes, int* out_of_range,
                    void* stream);
int ppo_tvf_value_loss(const float* targets, int64_t target_row_stride, const int32_t* head_index,
                       const float* predictions, int64_t pred_row_stride, int64_t pred_head_stride,
                       const float* head_weights, const float* mask, int64_t rows, int64_t heads, double coef,
                       float* loss_out, float* grad_out, void* stream);
/* out[b, k] = grad[b, k] * row_scale[b] (contiguous [rows, width]): the backward step of the loss above. */
int ppo_scale_rows(const float* grad, const float* row_scale, int64_t rows, int64_t width, float* out, void* stream);

/*
 * Stream-ordered strided copy between host (pinned) and device memory: `rows` rows of `width_bytes`, any direction.
 * Plumbing for calls whose rollout buffers live in host memory (the reference keeps them in NumPy arrays,
 * rl/rollout.py:646-680): an env block of a [N, A, K] host buffer is `N` rows of `A_block * K * 4` bytes with the
 * full row as pitch, so blocks can be uploaded / downloaded while other blocks are being computed.
 */
int ppo_memcpy_2d_async(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                        int64_t width_bytes, int64_t rows, void* stream);

/* Mailbox memory for ppo_gae_normalized_peers.  ppo_peer_alloc: zeroed cudaMalloc allocation on the current device and
 * its 64-byte CUDA IPC handle (exchange it between the processes of the node any way you like); ppo_peer_open: map a
 * peer's allocation for access from the CURRENT device (over NVLink); ppo_peer_close / ppo_peer_free undo them. */
int ppo_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int ppo_peer_open(const unsigned char* handle64, void** ptr);
int ppo_peer_close(void* ptr);
int ppo_peer_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* PPO_B200_H */
