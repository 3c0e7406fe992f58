// TMA ring kernel for the TVF targets (placeholder until the kernel lands).
#include "tvf_common.cuh"

namespace ppo {

int launch_tvf_ring(const TvfParams& p, cudaStream_t stream, bool probe_only) {
    (void)p;
    (void)stream;
    (void)probe_only;
    set_error("ppo_tvf_targets: the TMA ring kernel does not support this shape");
    return PPO_ERR_UNSUPPORTED;
}

}  // namespace ppo
