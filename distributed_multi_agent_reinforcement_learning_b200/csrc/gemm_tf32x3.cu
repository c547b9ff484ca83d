// gemm_tf32x3.cu — the dense E-wide layers of the actor/critic on the 5th-generation tensor cores (tcgen05 + TMEM),
// with fp32-level accuracy: C = act(A1 * W[:, :K1]^T + A2 * W[:, K1:]^T + bias + D).
//
// Why not plain TF32: parity with the reference is specified at 1e-5 (activations) / 1e-4 (loss, gradients); a single
// TF32 product carries 2^-11 relative error.  Every fp32 operand is therefore split on the fly into hi + lo TF32 parts
// (hi = fp32 rounded to 10 mantissa bits, lo = fp32(a - hi) rounded likewise, both with the discarded bits zeroed so the
// result does not depend on how the tensor core truncates) and the product is accumulated as
//     A_hi*B_hi + A_lo*B_hi + A_hi*B_lo        (fp32 accumulation in TMEM; dropped terms are O(2^-21))
// i.e. three tcgen05.mma.kind::tf32 per K-slice — still ~10x the throughput of the SIMT sgemm the library picks for
// fp32 on sm_100.  The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
// accumulation steps; to keep it at the 1e-6 level the hi*hi products alternate between two TMEM accumulators and the
// (2^-11 smaller) correction products go to a third, summed in round-to-nearest fp32 by the epilogue.
//
// Structure (one CTA = one 128x128 output tile, 5 warps):
//   warps 0-3  producers: coalesced LDG.128 of the A / W k-blocks (128 rows x 32 fp32), hi/lo split in registers,
//              st.shared into the K-major SWIZZLE_128B canonical layout (chunk16 ^= row&7), fence.proxy.async,
//              mbarrier arrive; afterwards the same warps run the epilogue (tcgen05.ld of their TMEM lane quarter,
//              bias / additive input / ReLU, st.global);
//   warp 4     allocates 512 TMEM columns (three 128-column accumulators); its elected lane waits on the "full" barriers and issues the MMAs,
//              releasing each smem stage with tcgen05.commit -> "empty" barrier and finally signalling the epilogue.
// 3 smem stages x (A_hi, A_lo, B_hi, B_lo) x 16 KB = 192 KB.
#include "common.cuh"

namespace marl {

namespace g3 {
constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB: 128 rows x 128 B
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/;
constexpr int THREADS = 160;

struct Args {
    const float *A1, *A2, *W, *bias, *D;
    float *C;
    int64_t lda1, lda2, ldw, ldd, ldc;
    int M, N, K1, K2, relu;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (SBO = 64 x 16 B), LBO = 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                       // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                                 // layout type SWIZZLE_128B, bits [61,64)
    return d;
}

// kind::tf32, fp32 accumulate, M=128, N=128, both operands K-major
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one_lane()      // see tc_common.cuh: elect_one
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// hi = a rounded to nearest at 10 mantissa bits (low 13 bits cleared), lo = (a - hi) rounded the same way
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo)
{
    const uint32_t u = __float_as_uint(a);
    hi = __uint_as_float((u + 0x1000u) & 0xFFFFE000u);
    const float r = a - hi;                                  // exact in fp32
    lo = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void store_split(unsigned char *hi_tile, unsigned char *lo_tile, int row, int chunk, float4 v)
{
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
    *reinterpret_cast<float4 *>(hi_tile + off) = h;
    *reinterpret_cast<float4 *>(lo_tile + off) = l;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(Args a)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);   // full[3], empty[3], tmem_full
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int KB = (a.K1 + a.K2) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 128);             // full: every producer thread arrives
            mbar_init(smem_u32(&bars[STAGES + s]), 1);      // empty: one tcgen05.commit
        }
        mbar_init(smem_u32(&bars[2 * STAGES]), 1);          // accumulator ready
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int t = threadIdx.x, chunk = t & 7, r0 = t >> 3;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES, round = kb / STAGES;
            if (round > 0) mbar_wait(smem_u32(&bars[STAGES + s]), (round - 1) & 1);
            unsigned char *st = smem + s * STAGE_BYTES;
            const int k0 = kb * BK;
            const float *Asrc;
            int64_t lda;
            int ka;
            if (k0 < a.K1) { Asrc = a.A1; lda = a.lda1; ka = k0; }
            else { Asrc = a.A2; lda = a.lda2; ka = k0 - a.K1; }
            float4 va[8], vb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + 16 * i;
                const int gm = m0 + row;
                va[i] = gm < a.M ? __ldg(reinterpret_cast<const float4 *>(Asrc + (int64_t)gm * lda + ka + 4 * chunk))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                vb[i] = __ldg(reinterpret_cast<const float4 *>(a.W + (int64_t)(n0 + row) * a.ldw + k0 + 4 * chunk));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + 16 * i;
                store_split(st, st + TILE_BYTES, row, chunk, va[i]);
                store_split(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, row, chunk, vb[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            mbar_arrive(smem_u32(&bars[s]));
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(smem_u32(&bars[2 * STAGES]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane, gm = m0 + row;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32], v1[32], v2[32];
#define MARL_TMEM_LD32(dst, addr)                                                                                       \
    asm volatile(                                                                                                       \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                       \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                       \
        : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]), "=r"(dst[7]),   \
          "=r"(dst[8]), "=r"(dst[9]), "=r"(dst[10]), "=r"(dst[11]), "=r"(dst[12]), "=r"(dst[13]), "=r"(dst[14]),            \
          "=r"(dst[15]), "=r"(dst[16]), "=r"(dst[17]), "=r"(dst[18]), "=r"(dst[19]), "=r"(dst[20]), "=r"(dst[21]),          \
          "=r"(dst[22]), "=r"(dst[23]), "=r"(dst[24]), "=r"(dst[25]), "=r"(dst[26]), "=r"(dst[27]), "=r"(dst[28]),          \
          "=r"(dst[29]), "=r"(dst[30]), "=r"(dst[31])                                                                      \
        : "r"(addr))
            MARL_TMEM_LD32(v, taddr + (uint32_t)c0);
            MARL_TMEM_LD32(v2, taddr + (uint32_t)(2 * BN + c0));
            if (KB > 1) MARL_TMEM_LD32(v1, taddr + (uint32_t)(BN + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float acc = __uint_as_float(v[j]);
                if (KB > 1) acc += __uint_as_float(v1[j]);
                v[j] = __float_as_uint(acc + __uint_as_float(v2[j]));
            }
            if (gm < a.M) {
                float *crow = a.C + (int64_t)gm * a.ldc + n0 + c0;
                const float *drow = a.D ? a.D + (int64_t)gm * a.ldd + n0 + c0 : nullptr;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
                    if (a.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + c0 + j));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (drow) {
                        const float4 d = __ldg(reinterpret_cast<const float4 *>(drow + j));
                        o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
                    }
                    if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4 *>(crow + j) = o;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES, round = kb / STAGES;
            mbar_wait(smem_u32(&bars[s]), round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
            if (elect_one_lane()) {
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    const uint64_t a_hi = make_desc(base + kk * 32), a_lo = make_desc(base + TILE_BYTES + kk * 32);
                    const uint64_t b_hi = make_desc(base + 2 * TILE_BYTES + kk * 32), b_lo = make_desc(base + 3 * TILE_BYTES + kk * 32);
                    // main accumulator kb&1 (columns 0 / 128), corrections in columns 256..383
                    umma_tf32(tmem_base + (uint32_t)(kb & 1) * BN, a_hi, b_hi, (kb >= 2 || kk) ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, a_lo, b_hi, (kb | kk) ? 1u : 0u);
                    umma_tf32(tmem_base + 2 * BN, a_hi, b_lo, 1u);
                }
                umma_commit(smem_u32(&bars[STAGES + s]));     // frees this smem stage once the MMAs have read it
            }
            __syncwarp();
        }
        if (elect_one_lane()) umma_commit(smem_u32(&bars[2 * STAGES]));   // accumulator complete -> epilogue
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace g3
}  // namespace marl

using namespace marl;

// C[M,N] = act(A1[M,K1] W[:, :K1]^T + A2[M,K2] W[:, K1:]^T + bias[N] + D[M,N]); W is [N, K1+K2] row-major (nn.Linear).
// N % 128 == 0, K1 % 32 == 0, K2 % 32 == 0 (K2 may be 0), all leading dimensions multiples of 4 floats, 16-B aligned.
extern "C" int marl_gemm_tf32x3(int32_t M, int32_t N, int32_t K1, int32_t K2, const float *d_A1, int64_t lda1,
                                const float *d_A2, int64_t lda2, const float *d_W, int64_t ldw, const float *d_bias,
                                const float *d_D, int64_t ldd, float *d_C, int64_t ldc, int32_t relu, void *stream)
{
    MARL_REQUIRE(M > 0 && N > 0 && (N % g3::BN) == 0 && K1 > 0 && (K1 % g3::BK) == 0 && K2 >= 0 && (K2 % g3::BK) == 0,
                 "marl_gemm_tf32x3: M=%d N=%d K1=%d K2=%d (need N%%128==0, K%%32==0)", M, N, K1, K2);
    MARL_REQUIRE(d_A1 && d_W && d_C && (K2 == 0 || d_A2), "marl_gemm_tf32x3: null pointer");
    MARL_REQUIRE((lda1 % 4) == 0 && (ldw % 4) == 0 && (ldc % 4) == 0 && (K2 == 0 || (lda2 % 4) == 0) && (!d_D || (ldd % 4) == 0),
                 "marl_gemm_tf32x3: leading dimensions must be multiples of 4 floats");
    auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    MARL_REQUIRE(al(d_A1) && al(d_W) && al(d_C) && al(d_A2) && al(d_bias) && al(d_D), "marl_gemm_tf32x3: pointers must be 16-byte aligned");
    g3::Args a;
    a.A1 = d_A1; a.A2 = d_A2; a.W = d_W; a.bias = d_bias; a.D = d_D; a.C = d_C;
    a.lda1 = lda1; a.lda2 = lda2; a.ldw = ldw; a.ldd = ldd; a.ldc = ldc;
    a.M = M; a.N = N; a.K1 = K1; a.K2 = K2; a.relu = relu;
    cudaError_t e = cudaFuncSetAttribute(g3::gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g3::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_tf32x3: smem %d: %s", g3::SMEM_BYTES, cudaGetErrorString(e)); return MARL_ECUDA; }
    dim3 grid((M + g3::BM - 1) / g3::BM, N / g3::BN);
    g3::gemm_tf32x3_kernel<<<grid, g3::THREADS, g3::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("gemm_tf32x3_kernel");
}
