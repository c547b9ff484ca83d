// gnn_kernels.cu — the non-GEMM part of the second network family (obstacle_differ_3hop/mappo_parallel.py:34-70,
// GnnExtractor.forward): L1-normalised adjacency-weighted means over the N+O entities of every agent, forward and backward.
//   h0_agg[s,i,:]   = sum_j w_ij h0[s,i,j,:]                 (:57-60, adj.unsqueeze(-2) @ h0)
//   comm_agg[s,i,:] = sum_{j<N} w_ij last_comm[s,j,:]        (:63-66, adj[..., :N] @ last_comm_embedding)
// with w_ij = adj[s,i,j] / max(sum_{j<J} |adj[s,i,j]|, 1e-12) (F.normalize p=1 over ALL J entities, :55), or 1/J for the critic's
// all-ones adjacency (:171).  One warp per (sample, agent) row, lane l owns channels l, l+32, ...: every access is a coalesced
// 128-byte line per 32 channels.  HBM-bound: the [S,N,J,E] tensor is read (forward) / written (backward) exactly once.
#include "common.cuh"
#include <type_traits>

namespace marl {

template <int CPL>
__global__ void __launch_bounds__(128)
entity_agg_fwd_kernel(int64_t rows, int N, int J, int jmax, const float *__restrict__ adj, int all_ones, const float *__restrict__ x,
                      int64_t ss, int64_t as, int64_t js, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31, E = 32 * CPL;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int64_t s = row / N;
    const int i = (int)(row - s * N);
    const float *arow = adj + row * J;
    float l1 = 0.f;
    if (all_ones) l1 = (float)J;
    else {
        for (int j = lane; j < J; j += 32) l1 += fabsf(arow[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    const float inv = 1.f / fmaxf(l1, 1e-12f);
    float acc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[q] = 0.f;
    const float *xb = x + s * ss + (int64_t)i * as;
    for (int j = 0; j < jmax; ++j) {
        const float wj = all_ones ? 1.f : arow[j];
        if (wj == 0.f) continue;                                  // warp-uniform: sparse adjacency rows skip the load
        const float *xr = xb + (int64_t)j * js;
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = fmaf(wj * inv, xr[lane + 32 * q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < CPL; ++q) out[row * E + lane + 32 * q] = acc[q];
}

// dx[s,i,j,:] = w_ij dout[s,i,:] for the dense per-agent entity layout [rows, J, E]
template <int CPL>
__global__ void __launch_bounds__(128)
entity_agg_bwd_kernel(int64_t rows, int J, const float *__restrict__ adj, int all_ones, const float *__restrict__ dout,
                      float *__restrict__ dx)
{
    const int lane = threadIdx.x & 31, E = 32 * CPL;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *arow = adj + row * J;
    float l1 = 0.f;
    if (all_ones) l1 = (float)J;
    else {
        for (int j = lane; j < J; j += 32) l1 += fabsf(arow[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    const float inv = 1.f / fmaxf(l1, 1e-12f);
    float g[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) g[q] = dout[row * E + lane + 32 * q];
    float *xb = dx + row * (int64_t)J * E;
    for (int j = 0; j < J; ++j) {
        const float wj = (all_ones ? 1.f : arow[j]) * inv;
#pragma unroll
        for (int q = 0; q < CPL; ++q) xb[(int64_t)j * E + lane + 32 * q] = wj * g[q];
    }
}

template <typename F>
static int dispatch_cpl8(int E, F &&f)
{
    if (E == 32) return f(std::integral_constant<int, 1>{});
    if (E == 64) return f(std::integral_constant<int, 2>{});
    if (E == 128) return f(std::integral_constant<int, 4>{});
    if (E == 256) return f(std::integral_constant<int, 8>{});
    set_error("marl_entity_agg: channel width %d unsupported (32, 64, 128, 256)", E);
    return MARL_EUNSUPPORTED;
}

}  // namespace marl

using namespace marl;

extern "C" int marl_entity_agg_fwd(int64_t rows, int32_t N, int32_t J, int32_t jmax, int32_t E, const float *d_adj, int32_t all_ones,
                                   const float *d_x, int64_t sample_stride, int64_t agent_stride, int64_t entity_stride, float *d_out,
                                   void *stream)
{
    MARL_REQUIRE(rows > 0 && N >= 1 && J >= 1 && jmax >= 1 && jmax <= J && (rows % N) == 0, "marl_entity_agg_fwd: rows=%lld N=%d J=%d jmax=%d",
                 (long long)rows, N, J, jmax);
    MARL_REQUIRE(d_adj && d_x && d_out, "marl_entity_agg_fwd: null pointer");
    return dispatch_cpl8(E, [&](auto C_) -> int {
        constexpr int CPL = decltype(C_)::value;
        entity_agg_fwd_kernel<CPL><<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(rows, N, J, jmax, d_adj, all_ones, d_x, sample_stride,
                                                                                                   agent_stride, entity_stride, d_out);
        return check_launch("entity_agg_fwd_kernel");
    });
}

extern "C" int marl_entity_agg_bwd(int64_t rows, int32_t J, int32_t E, const float *d_adj, int32_t all_ones, const float *d_dout, float *d_dx,
                                   void *stream)
{
    MARL_REQUIRE(rows > 0 && J >= 1 && d_adj && d_dout && d_dx, "marl_entity_agg_bwd: bad arguments");
    return dispatch_cpl8(E, [&](auto C_) -> int {
        constexpr int CPL = decltype(C_)::value;
        entity_agg_bwd_kernel<CPL><<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(rows, J, d_adj, all_ones, d_dout, d_dx);
        return check_launch("entity_agg_bwd_kernel");
    });
}
