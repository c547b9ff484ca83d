// wgrad_tf32x3.cu — weight gradients of the dense layers on the tcgen05 tensor cores with fp32-level accuracy:
//     dW[n][k] (+)= sum_r dY[r][n] * X[r][k]        (nn.Linear / nn.GRU weight gradients inside loss.backward(),
//                                                    DHGN/mappo_parallel.py:708; r runs over the ~5e5 rows of a minibatch)
// This is a GEMM whose reduction dimension is the (huge) row dimension and whose operands are both stored row-major
// [rows, features], i.e. "MN-major" for the tensor core.  The producers therefore TRANSPOSE on the fly: 32 rows x 128
// features of each operand are loaded with 32-byte-sector-exact global loads and written into the K-major SWIZZLE_128B
// canonical layout ([128 features x 32 rows], hi and lo TF32 planes) with 4-byte stores arranged so that a warp covers an
// 8 (features) x 4 (rows) patch = 32 distinct banks.  (A float4-load + in-register 4x4 transpose variant was measured 4.5x
// slower: the lane-dependent selects of the transpose quadruple the instruction count of the producers, which are the
// bottleneck.)  The column sums of dY (the bias gradient) fall out of the same registers.  The N tiles of one row slice are
// adjacent in the grid (blockIdx.x = tile), so the X block they share is read from HBM once and from L2 afterwards.
// Split-K over row ranges (<= 64 k-blocks per CTA so that no TMEM
// accumulator sees more than 128 truncating accumulate steps; main products alternate between two accumulators, the
// 2^-11-small correction products go to a third, all summed in round-to-nearest fp32 by the epilogue), per-slice partial
// tiles in a workspace, and a deterministic reduction kernel (no atomics: gradients are bit-reproducible run to run).
// CTA = 8 producer warps (they also run the epilogue) + 1 MMA-issuing warp; 3 stages x 64 KB.
#include "tc_common.cuh"

namespace marl {
namespace wg {

using namespace tc;

constexpr int BK = 32, STAGES = 3;
constexpr int PLANE = 128 * 128;                 // [128 features x 32 rows] x 4 B
constexpr int STAGE_BYTES = 4 * PLANE;           // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
constexpr int THREADS = 288;
constexpr int MAX_KB_PER_SLICE = 64;

struct Args {
    const float *dY, *X;
    float *partial;                              // [slices, tiles, 128, 128]
    float *bias_partial;                         // [slices, N_out] or null (written by the k_tile == 0 CTAs)
    int64_t lddy, ldx, R, rows_per_slice;
    int n_tiles_k;                               // K_in / 128 (tile index = n_tile * n_tiles_k + k_tile)
};

// transposed, split store of one operand: rows r0 + 4*rg + (lane >> 3), features c0 + 8*cg + (lane & 7)
__device__ __forceinline__ void load_patch_row(const float *__restrict__ src, int64_t ld, int64_t R, int64_t row, int c0, int lane, float (&v)[16])
{
    const bool ok = row < R;
    const float *p = src + row * ld + c0 + (lane & 7);
#pragma unroll
    for (int cg = 0; cg < 16; ++cg) v[cg] = ok ? __ldg(p + 8 * cg) : 0.f;
}
__device__ __forceinline__ void store_patch_row(unsigned char *hi_plane, int r, int lane, const float (&v)[16])
{
#pragma unroll
    for (int cg = 0; cg < 16; ++cg) {
        const int c = 8 * cg + (lane & 7);
        float hi, lo;
        split_tf32(v[cg], hi, lo);
        const int off = c * 128 + ((((r >> 2) ^ (c & 7))) << 4) + (r & 3) * 4;
        *reinterpret_cast<float *>(hi_plane + off) = hi;
        *reinterpret_cast<float *>(hi_plane + PLANE + off) = lo;
    }
}

__global__ void __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ Args a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);     // full[3], empty[3], done
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, slice = blockIdx.y, n0 = (tile / a.n_tiles_k) * 128, k0 = (tile % a.n_tiles_k) * 128;
    const int64_t r_begin = (int64_t)slice * a.rows_per_slice;
    int64_t r_end = r_begin + a.rows_per_slice;
    if (r_end > a.R) r_end = a.R;
    const int KB = r_end > r_begin ? (int)((r_end - r_begin + BK - 1) / BK) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 8);
            mbar_init(smem_u32(&bars[STAGES + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *slot;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers: warp w owns rows 4w .. 4w+3 of every k-block
        const int r_in = 4 * warp + (lane >> 3);
        // software pipeline: the global loads of k-block kb+1 are in flight while k-block kb is split and stored (the HBM latency
        // of a 32-row block would otherwise be exposed once per k-block: the producers are the same threads that wait for the stage)
        float va[16], vb[16], na[16], nb[16];
        float colsum[16];                                        // bias gradient: column sums of dY over this thread's rows
#pragma unroll
        for (int cg = 0; cg < 16; ++cg) colsum[cg] = 0.f;
        auto fetch = [&](int kb, float (&ua)[16], float (&ub)[16]) {
            const int64_t row = r_begin + (int64_t)kb * BK + r_in;
            load_patch_row(a.dY, a.lddy, r_end, row, n0, lane, ua);
            load_patch_row(a.X, a.ldx, r_end, row, k0, lane, ub);
        };
        auto publish = [&](int kb, const float (&ua)[16], const float (&ub)[16]) {
            const int s = kb % STAGES, round = kb / STAGES;
#pragma unroll
            for (int cg = 0; cg < 16; ++cg) colsum[cg] += ua[cg];
            if (round > 0) mbar_wait(smem_u32(&bars[STAGES + s]), (uint32_t)((round - 1) & 1));
            unsigned char *st = smem + s * STAGE_BYTES;
            store_patch_row(st, r_in, lane, ua);
            store_patch_row(st + 2 * PLANE, r_in, lane, ub);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[s]));       // one arrival per warp (the fence above is per thread)
        };
        if (KB > 0) fetch(0, va, vb);
        for (int kb = 0; kb < KB; kb += 2) {                    // two register sets, no copies: loads stay in flight across a whole iteration
            if (kb + 1 < KB) fetch(kb + 1, na, nb);
            publish(kb, va, vb);
            if (kb + 1 < KB) {
                if (kb + 2 < KB) fetch(kb + 2, va, vb);
                publish(kb + 1, na, nb);
            }
        }
        // ------------------------------------------------------------------ epilogue: partial tile = main0 + main1 + corr
        mbar_wait(smem_u32(&bars[2 * STAGES]), 0);
        fence_after();
        const int n = 32 * (warp & 3) + lane, half = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
        float *out = a.partial + (((size_t)slice * gridDim.x + tile) * 128 + n) * 128;
#pragma unroll 1
        for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 16) {
            uint32_t m0[16], m1[16], cr[16];
            TC_TMEM_LD16(m0, taddr + (uint32_t)c0);
            TC_TMEM_LD16(m1, taddr + (uint32_t)(128 + c0));
            TC_TMEM_LD16(cr, taddr + (uint32_t)(256 + c0));
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 o;
                float *of = reinterpret_cast<float *>(&o);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = KB > 0 ? __uint_as_float(m0[j + q]) : 0.f;
                    if (KB > 1) v += __uint_as_float(m1[j + q]);
                    if (KB > 0) v += __uint_as_float(cr[j + q]);
                    of[q] = v;
                }
                *reinterpret_cast<float4 *>(out + c0 + j) = o;
            }
        }
        if (a.bias_partial && (tile % a.n_tiles_k) == 0) {
            // all MMAs are complete: stage 0 is free; sum the rows of a warp with shuffles, the 8 warps through shared memory
            // (fixed order: deterministic)
            float *sb = reinterpret_cast<float *>(smem);          // [8 warps][128]
#pragma unroll
            for (int cg = 0; cg < 16; ++cg) {
                float v = colsum[cg];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 8) sb[warp * 128 + 8 * cg + lane] = v;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x < 128) {
                float v = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) v += sb[w8 * 128 + threadIdx.x];
                a.bias_partial[(size_t)slice * (gridDim.x / a.n_tiles_k) * 128 + (size_t)(tile / a.n_tiles_k) * 128 + threadIdx.x] = v;
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
        const uint32_t idesc = idesc_tf32(128, 128);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES, round = kb / STAGES;
            mbar_wait(smem_u32(&bars[s]), (uint32_t)(round & 1));
            fence_after();
            const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t a_hi0 = make_desc(base), a_lo0 = make_desc(base + PLANE), b_hi0 = make_desc(base + 2 * PLANE), b_lo0 = make_desc(base + 3 * PLANE);
            const uint32_t main_acc = tmem_base + (uint32_t)(kb & 1) * 128u, corr = tmem_base + 256u;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    umma_tf32(main_acc, a_hi0 + 2 * kk, b_hi0 + 2 * kk, idesc, (kb >= 2 || kk) ? 1u : 0u);
                    umma_tf32(corr, a_lo0 + 2 * kk, b_hi0 + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
                    umma_tf32(corr, a_hi0 + 2 * kk, b_lo0 + 2 * kk, idesc, 1u);
                }
                umma_commit(smem_u32(&bars[STAGES + s]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bars[2 * STAGES]));
        __syncwarp();
    }
    fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// dW[n0+n][k0+k] (+)= sum_s partial[s][tile][n][k]   (fixed summation order: deterministic)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ partial, int slices, int tiles, int n_tiles_k, float *__restrict__ dW, int64_t lddw, int accumulate)
{
    const int tile = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;          // float4 index inside the 128x128 tile
    if (idx >= 128 * 32) return;
    const int n = idx >> 5, k4 = (idx & 31) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < slices; ++s) {
        const float4 v = *reinterpret_cast<const float4 *>(partial + (((size_t)s * tiles + tile) * 128 + n) * 128 + k4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float *dst = dW + (int64_t)((tile / n_tiles_k) * 128 + n) * lddw + (tile % n_tiles_k) * 128 + k4;
    if (accumulate) { dst[0] += acc.x; dst[1] += acc.y; dst[2] += acc.z; dst[3] += acc.w; }
    else { dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w; }
}

// db[n] (+)= sum_s bias_partial[s][n]
__global__ void __launch_bounds__(128)
wgrad_bias_reduce_kernel(const float *__restrict__ bias_partial, int slices, int N_out, float *__restrict__ db, int accumulate)
{
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N_out) return;
    float acc = 0.f;
    for (int s = 0; s < slices; ++s) acc += bias_partial[(size_t)s * N_out + n];
    db[n] = accumulate ? db[n] + acc : acc;
}

static int64_t slices_for(int64_t R)
{
    const int64_t kbs = (R + BK - 1) / BK;
    int64_t s = (kbs + MAX_KB_PER_SLICE - 1) / MAX_KB_PER_SLICE;
    return s < 1 ? 1 : s;
}

}  // namespace wg
}  // namespace marl

using namespace marl;

extern "C" int64_t marl_wgrad_workspace_bytes(int64_t R, int32_t N_out, int32_t K_in)
{
    if (R <= 0 || N_out <= 0 || K_in <= 0) return -1;
    return wg::slices_for(R) * ((int64_t)(N_out / 128) * (K_in / 128) * 128 * 128 + N_out) * 4;
}

// d_dbias (may be NULL): f32 [N_out] receives (or accumulates) sum_r dY[r][:], computed from the same loads
extern "C" int marl_wgrad_tf32x3(int64_t R, int32_t N_out, int32_t K_in, const float *d_dY, int64_t lddy, const float *d_X, int64_t ldx,
                                 float *d_dW, int64_t lddw, float *d_dbias, int32_t accumulate, void *d_workspace, void *stream)
{
    MARL_REQUIRE(R > 0 && N_out > 0 && K_in > 0 && (N_out % 128) == 0 && (K_in % 128) == 0, "marl_wgrad_tf32x3: R=%lld N_out=%d K_in=%d (need multiples of 128)",
                 (long long)R, N_out, K_in);
    MARL_REQUIRE(d_dY && d_X && d_dW && d_workspace && lddy >= N_out && ldx >= K_in && lddw >= K_in, "marl_wgrad_tf32x3: bad pointer / leading dimension");
    const int64_t slices = wg::slices_for(R);
    const int tiles = (N_out / 128) * (K_in / 128);
    MARL_REQUIRE(slices <= 65535, "marl_wgrad_tf32x3: R too large");
    wg::Args a;
    a.dY = d_dY; a.X = d_X; a.partial = static_cast<float *>(d_workspace); a.lddy = lddy; a.ldx = ldx; a.R = R;
    a.bias_partial = d_dbias ? a.partial + (size_t)slices * tiles * 128 * 128 : nullptr;
    const int64_t kbs = (R + wg::BK - 1) / wg::BK;
    a.rows_per_slice = ((kbs + slices - 1) / slices) * wg::BK;
    a.n_tiles_k = K_in / 128;
    cudaError_t e = cudaFuncSetAttribute(wg::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("wgrad_kernel: smem %d: %s", wg::SMEM_BYTES, cudaGetErrorString(e)); return MARL_ECUDA; }
    wg::wgrad_kernel<<<dim3((unsigned)tiles, (unsigned)slices), wg::THREADS, wg::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    int rc = check_launch("wgrad_kernel");
    if (rc) return rc;
    wg::wgrad_reduce_kernel<<<dim3(16, (unsigned)tiles), 256, 0, (cudaStream_t)stream>>>(a.partial, (int)slices, tiles, a.n_tiles_k, d_dW, lddw, accumulate);
    rc = check_launch("wgrad_reduce_kernel");
    if (rc || !d_dbias) return rc;
    wg::wgrad_bias_reduce_kernel<<<(N_out + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a.bias_partial, (int)slices, N_out, d_dbias, accumulate);
    return check_launch("wgrad_bias_reduce_kernel");
}
