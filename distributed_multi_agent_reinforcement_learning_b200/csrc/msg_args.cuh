// msg_args.cuh — argument block shared by the DHGN message kernels (policy_kernels.cu, msg_grouped.cu).
#pragma once
#include "common.cuh"

namespace marl {

struct MsgArgs {
    int S, N, O, NW, OW;
    const float *p;            // [S,N,4]
    const float *e;            // [S,4]
    const float *oxy;          // [Bo,O,2] obstacle-cell coordinates (vx=vy=0 implied, pursuit_env.py:22-26)
    const int32_t *o_index;    // [S] row of oxy for this sample
    const int32_t *o_count;    // [Bo] slots that exist for all-ones adjacency (critic): O_b in rollout, O in training
    const uint32_t *p_adj;     // [S,N,NW] or null when all_ones
    const uint8_t *e_adj;      // [S,N]
    const uint32_t *o_adj;     // [S,N,OW]
    int all_ones;              // critic: AttributeDataset(is_critic=True) (mappo_parallel.py:64-65)
};

// sample-grouped kernels (msg_grouped.cu): E = 128, N <= 16, O <= 256; return MARL_EUNSUPPORTED otherwise
int launch_msg_fwd_grouped(const MsgArgs &a, const float *W0, const float *b0, const float *W1, const float *b1, const float *W2,
                           const float *b2, float *agg, cudaStream_t stream);
int launch_msg_bwd_grouped(const MsgArgs &a, const float *W0, const float *b0, const float *W1, const float *b1, const float *W2,
                           const float *b2, const float *d_agg, float *gW0, float *gb0, float *gW1, float *gb1, float *gW2, float *gb2,
                           cudaStream_t stream);
bool msg_grouped_supported(const MsgArgs &a, int E);

}  // namespace marl
