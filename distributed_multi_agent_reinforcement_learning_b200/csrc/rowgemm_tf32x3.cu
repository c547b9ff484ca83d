// rowgemm_tf32x3.cu — the dense layers of TRAINING (forward and input-gradient GEMMs) as a persistent row-tile kernel:
//     C[M, N] = act([A1 | A2][M, K1+K2] * W^T + bias + D),   M ~ 5e5 rows, N <= 512, K1 + K2 <= 512, all multiples of 128
// (every E-wide nn.Linear / GRU input projection of DHGN/mappo_parallel.py:116-545 in MAPPO.train, and their dX = dY W).
// Same machinery as the fused rollout kernel (policy_fused.cu): a CTA owns 128 rows at a time, the activation chunk
// [128 rows x 128 k] lives in shared memory as hi/lo TF32 planes (K-major SWIZZLE_128B), the weights are pre-split / pre-swizzled
// once per weight version (marl_rowgemm_pack) and streamed by one thread with cp.async.bulk, products are 3xTF32
// tcgen05.mma.kind::tf32 with fp32 accumulation in TMEM (one 128-column accumulator per 128 output features).
// Differences from gemm_tf32x3.cu (one 128x128 output tile per CTA, producers splitting A AND W per k-block): the A chunk is
// loaded and split once and reused for all N tiles, weights cost no ALU work, CTAs are persistent (barriers / TMEM set up once),
// and every global access is a coalesced 512-byte row segment (the epilogue stages the fp32 tile through shared memory).
#include "tc_common.cuh"

namespace marl {
namespace rg {

using namespace tc;

constexpr int ROWS = 128, NSTAGE = 3;
constexpr int TILE = ROWS * 128;             // one plane of a k-block
constexpr int XKB = 2 * TILE;
constexpr int X_BYTES = 4 * XKB;             // [128 rows x K=128], hi + lo
constexpr int WSTAGE = 2 * TILE;
constexpr int UNIT_BYTES = 4 * WSTAGE;
constexpr int SMEM_BYTES = X_BYTES + NSTAGE * WSTAGE + 256;
constexpr int THREADS = 320;
constexpr int BAR_A_READY = 2 * NSTAGE, BAR_MMA_DONE = 2 * NSTAGE + 1;

struct Args {
    const float *A1, *A2, *bias, *D;
    float *C;
    const unsigned char *packed;             // units (kc, nt), kc-major
    int64_t lda1, lda2, ldd, ldc, M;
    int KC1, KC, NT, relu, n_tiles;
};

__device__ __forceinline__ int x_off(int row, int c4) { return (c4 >> 3) * XKB + row * 128 + (((c4 & 7) ^ (row & 7)) << 4); }
__device__ __forceinline__ void x_store4(unsigned char *X, int row, int c4, float4 v)
{
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    const int off = x_off(row, c4);
    *reinterpret_cast<float4 *>(X + off) = h;
    *reinterpret_cast<float4 *>(X + off + TILE) = l;
}
__device__ __forceinline__ float4 x_load4(const unsigned char *X, int row, int c4)
{
    const int off = x_off(row, c4);
    const float4 h = *reinterpret_cast<const float4 *>(X + off), l = *reinterpret_cast<const float4 *>(X + off + TILE);
    return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
rowgemm_kernel(const __grid_constant__ Args a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    unsigned char *X = smem, *Wst = smem + X_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + X_BYTES + NSTAGE * WSTAGE);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[NSTAGE + s]), 1);
        }
        mbar_init(smem_u32(&bars[BAR_A_READY]), 256);
        mbar_init(smem_u32(&bars[BAR_MMA_DONE]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *slot;
    const int n_units = a.KC * a.NT;

    if (warp < 8) {
        // ================================================================================= workers
        int group = 0;
        float4 nxt[16];
        bool have_next = false;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            const int64_t row0 = (int64_t)tile * ROWS;
            for (int kc = 0; kc < a.KC; ++kc) {
                // activation chunk kc -> X: warp w loads rows 16w..16w+15, 512 contiguous bytes per row, all loads in flight first
                const float *src = kc < a.KC1 ? a.A1 + kc * 128 : a.A2 + (kc - a.KC1) * 128;
                const int64_t ld = kc < a.KC1 ? a.lda1 : a.lda2;
                float4 v[16];
                if (kc == 0 && have_next) {                     // chunk 0 of this tile was prefetched during the previous epilogue
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) v[rr] = nxt[rr];
                } else {
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        const int64_t row = row0 + 16 * warp + rr;
                        v[rr] = row < a.M ? __ldg(reinterpret_cast<const float4 *>(src + row * ld) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (kc > 0) {                                   // X is still being read by the previous chunk's MMAs
                    mbar_wait(smem_u32(&bars[BAR_MMA_DONE]), (uint32_t)((group - 1) & 1));
                    fence_after();
                }
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) x_store4(X, 16 * warp + rr, lane, v[rr]);
                fence_async_smem();
                fence_before();
                mbar_arrive(smem_u32(&bars[BAR_A_READY]));
                ++group;
            }
            // prefetch chunk 0 of the next tile: its HBM latency hides behind this tile's MMAs and epilogue
            have_next = tile + (int)gridDim.x < a.n_tiles;
            if (have_next) {
                const int64_t nrow0 = (int64_t)(tile + gridDim.x) * ROWS;
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) {
                    const int64_t row = nrow0 + 16 * warp + rr;
                    nxt[rr] = row < a.M ? __ldg(reinterpret_cast<const float4 *>(a.A1 + row * a.lda1) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            mbar_wait(smem_u32(&bars[BAR_MMA_DONE]), (uint32_t)((group - 1) & 1));
            fence_after();
            // epilogue: per 128-column output tile, raw accumulator -> X (thread = row), then coalesced rows -> global with
            // bias / additive input / ReLU applied on the way out
            const int row = 32 * (warp & 3) + lane, hh = warp >> 2;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
            for (int nt = 0; nt < a.NT; ++nt) {
#pragma unroll 1
                for (int c0 = 64 * hh; c0 < 64 * hh + 64; c0 += 32) {
                    uint32_t acc[32];
                    TC_TMEM_LD16(acc, taddr + (uint32_t)(nt * 128 + c0));
                    TC_TMEM_LD16((acc + 16), taddr + (uint32_t)(nt * 128 + c0 + 16));
                    if (a.NT <= 2) {
                        uint32_t cor[32];
                        TC_TMEM_LD16(cor, taddr + (uint32_t)(256 + nt * 128 + c0));
                        TC_TMEM_LD16((cor + 16), taddr + (uint32_t)(256 + nt * 128 + c0 + 16));
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(cor[j]));
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        x_store4(X, row, (c0 + j) >> 2, make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                                                      __uint_as_float(acc[j + 3])));
                }
                worker_sync();
                const float4 bb = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias + nt * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int rr = 0; rr < 16; ++rr) {
                    const int r = 16 * warp + rr;
                    const int64_t grow = row0 + r;
                    if (grow < a.M) {
                        float4 o = x_load4(X, r, lane);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        if (a.D) {
                            const float4 d = __ldg(reinterpret_cast<const float4 *>(a.D + grow * a.ldd + nt * 128) + lane);
                            o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
                        }
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *(reinterpret_cast<float4 *>(a.C + grow * a.ldc + nt * 128) + lane) = o;
                    }
                }
                worker_sync();
            }
        }
    } else if (warp == 8) {
        // ================================================================================= weight loader
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
                for (int u = 0; u < n_units; ++u)
                    for (int kb = 0; kb < 4; ++kb, ++it) {
                        const int s = it % NSTAGE, round = it / NSTAGE;
                        if (round > 0) mbar_wait(smem_u32(&bars[NSTAGE + s]), (uint32_t)((round - 1) & 1));
                        mbar_expect_tx(smem_u32(&bars[s]), WSTAGE);
                        bulk_g2s(smem_u32(Wst + s * WSTAGE), a.packed + (size_t)u * UNIT_BYTES + (size_t)kb * WSTAGE, WSTAGE, smem_u32(&bars[s]));
                    }
        }
    } else {
        // ================================================================================= MMA issuer (whole warp, elected lane issues)
        {
            const uint32_t idesc = idesc_tf32(128, 128);
            int it = 0, group = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x)
                for (int kc = 0; kc < a.KC; ++kc) {
                    mbar_wait(smem_u32(&bars[BAR_A_READY]), (uint32_t)(group & 1));
                    fence_after();
                    for (int nt = 0; nt < a.NT; ++nt) {
                        // up to two output tiles: the 2^-11-small correction products get their own accumulator (columns 256+),
                        // so the truncating fp32 accumulation of the main products is not disturbed (error ~1e-6 instead of ~3e-6
                        // at K = 384); with 3-4 tiles TMEM only has room for one accumulator per tile
                        const bool split_corr = a.NT <= 2;
                        const uint32_t acc = tmem_base + (uint32_t)(nt * 128);
                        const uint32_t corr = split_corr ? tmem_base + 256u + (uint32_t)(nt * 128) : acc;
                        for (int kb = 0; kb < 4; ++kb, ++it) {
                            const int s = it % NSTAGE, round = it / NSTAGE;
                            mbar_wait(smem_u32(&bars[s]), (uint32_t)(round & 1));
                            fence_after();
                            const uint32_t xa = smem_u32(X + kb * XKB), wb = smem_u32(Wst + s * WSTAGE);
                            const uint64_t a_hi0 = make_desc(xa), a_lo0 = make_desc(xa + TILE), b_hi0 = make_desc(wb), b_lo0 = make_desc(wb + TILE);
                            if (elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const uint32_t first = (kc | kb | kk) ? 1u : 0u;
                                    umma_tf32(acc, a_hi0 + 2 * kk, b_hi0 + 2 * kk, idesc, first);
                                    umma_tf32(corr, a_lo0 + 2 * kk, b_hi0 + 2 * kk, idesc, split_corr ? first : 1u);
                                    umma_tf32(corr, a_hi0 + 2 * kk, b_lo0 + 2 * kk, idesc, 1u);
                                }
                                umma_commit(smem_u32(&bars[NSTAGE + s]));
                            }
                            __syncwarp();
                        }
                    }
                    if (elect_one()) umma_commit(smem_u32(&bars[BAR_MMA_DONE]));
                    __syncwarp();
                    ++group;
                }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// unit (kc, nt): B[n][k] = W[(nt*128 + n) * stride_n + (kc*128 + k) * stride_k]
__global__ void __launch_bounds__(256)
pack_kernel(const float *__restrict__ W, int64_t stride_n, int64_t stride_k, int NT, unsigned char *__restrict__ out)
{
    const int u = blockIdx.y, kc = u / NT, nt = u % NT;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= 128 * 128) return;
    const int n = idx >> 7, k = idx & 127;
    float hi, lo;
    split_tf32(W[(int64_t)(nt * 128 + n) * stride_n + (int64_t)(kc * 128 + k) * stride_k], hi, lo);
    const int kb = k >> 5, kk = k & 31;
    const size_t off = (size_t)u * UNIT_BYTES + (size_t)kb * WSTAGE + (size_t)n * 128 + ((((kk >> 2) ^ (n & 7))) << 4) + (kk & 3) * 4;
    *reinterpret_cast<float *>(out + off) = hi;
    *reinterpret_cast<float *>(out + off + TILE) = lo;
}

}  // namespace rg
}  // namespace marl

using namespace marl;

extern "C" int64_t marl_rowgemm_pack_bytes(int32_t N, int32_t K) { return (N <= 0 || K <= 0 || N % 128 || K % 128) ? -1 : (int64_t)(N / 128) * (K / 128) * rg::UNIT_BYTES; }

// Packs the "weight" operand B[N, K] of C = A B^T given as B[n][k] = d_W[n*stride_n + k*stride_k]
// (nn.Linear weight: stride_n = ld, stride_k = 1; for dX = dY W use the transposed view: stride_n = 1, stride_k = ld).
extern "C" int marl_rowgemm_pack(const float *d_W, int64_t stride_n, int64_t stride_k, int32_t N, int32_t K, void *d_packed, void *stream)
{
    MARL_REQUIRE(d_W && d_packed && ((uintptr_t)d_packed & 1023) == 0 && N > 0 && K > 0 && (N % 128) == 0 && (K % 128) == 0,
                 "marl_rowgemm_pack: bad arguments (N=%d K=%d must be multiples of 128, 1024-byte aligned output)", N, K);
    rg::pack_kernel<<<dim3(64, (unsigned)((N / 128) * (K / 128))), 256, 0, (cudaStream_t)stream>>>(d_W, stride_n, stride_k, N / 128,
                                                                                                   static_cast<unsigned char *>(d_packed));
    return check_launch("rowgemm pack_kernel");
}

extern "C" int marl_rowgemm_tf32x3(int64_t M, int32_t N, int32_t K1, int32_t K2, const float *d_A1, int64_t lda1, const float *d_A2, int64_t lda2,
                                   const void *d_packed, const float *d_bias, const float *d_D, int64_t ldd, float *d_C, int64_t ldc, int32_t relu,
                                   void *stream)
{
    MARL_REQUIRE(M > 0 && N > 0 && N <= 512 && (N % 128) == 0 && K1 > 0 && (K1 % 128) == 0 && K2 >= 0 && (K2 % 128) == 0 && K1 + K2 <= 1024,
                 "marl_rowgemm_tf32x3: M=%lld N=%d K1=%d K2=%d", (long long)M, N, K1, K2);
    MARL_REQUIRE(d_A1 && d_packed && d_C && (K2 == 0 || d_A2), "marl_rowgemm_tf32x3: null pointer");
    auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    MARL_REQUIRE(al(d_A1) && al(d_A2) && al(d_bias) && al(d_D) && al(d_C) && (lda1 % 4) == 0 && (K2 == 0 || (lda2 % 4) == 0) && (ldc % 4) == 0 &&
                     (!d_D || (ldd % 4) == 0), "marl_rowgemm_tf32x3: pointers / leading dimensions must be 16-byte aligned");
    rg::Args a;
    a.A1 = d_A1; a.A2 = d_A2; a.bias = d_bias; a.D = d_D; a.C = d_C; a.packed = static_cast<const unsigned char *>(d_packed);
    a.lda1 = lda1; a.lda2 = lda2; a.ldd = ldd; a.ldc = ldc; a.M = M;
    a.KC1 = K1 / 128; a.KC = (K1 + K2) / 128; a.NT = N / 128; a.relu = relu;
    a.n_tiles = (int)((M + rg::ROWS - 1) / rg::ROWS);
    cudaError_t e = cudaFuncSetAttribute(rg::rowgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("rowgemm_kernel: smem %d: %s", rg::SMEM_BYTES, cudaGetErrorString(e)); return MARL_ECUDA; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.n_tiles < sms ? a.n_tiles : sms;
    rg::rowgemm_kernel<<<grid, rg::THREADS, rg::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("rowgemm_kernel");
}
