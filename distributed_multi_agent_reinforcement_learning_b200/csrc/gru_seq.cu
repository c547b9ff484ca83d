// gru_seq.cu — the recurrence of one nn.GRU layer over a whole sequence as ONE persistent kernel per direction
// (torch.nn.GRU inside SharedActor/SharedCritic.forward mode 1, DHGN/mappo_parallel.py:401-437,496-545; its autograd backward).
//
// The input projection gi = x W_ih^T + b_ih of all T steps is one big GEMM done by the caller; what remains is inherently
// sequential per row and, as separate launches, was 2 x T tiny GEMMs + 2 x T pointwise kernels per layer (12 000 launches of
// ~30 us per PPO epoch).  Here a CTA owns NR rows for all T steps and keeps the recurrent state on chip:
//   forward :  gh_t = h_{t-1} W_hh^T  (3 gate units, tcgen05 kind::tf32 3xTF32, accumulators in TMEM) -> cell -> h_t
//   backward:  cell backward -> dgi_t, dgh_t ;  dh_{t-1} = dh_t * z_t + dgh_t W_hh   (3 units accumulated into one accumulator)
// Transposed formulation: D^T[feature, row] = W[feature, k] * state[row, k]^T, i.e. the WEIGHTS are the M=128 operand and the
// recurrent state is the N=NR operand.  The accumulator then has lane = feature, column = row, so an epilogue thread owns one
// feature: its bias is a register, and for a fixed row the 32 lanes of a warp touch 32 consecutive floats of gi / out / saves
// (coalesced 128-byte lines) and one 128-byte swizzled row chunk of the state tile in shared memory (conflict-free).
// W_hh (hi/lo pre-split, pre-swizzled by marl_gru_pack) is streamed from L2 every step with cp.async.bulk (384 KB per step per
// CTA; it cannot stay resident: 3 gates x 128 KB).  10 warps: 8 workers, loader, MMA issuer — same hand-over protocol as
// policy_fused.cu.  Element arithmetic is the one of gru_cell_{fwd,bwd}_kernel (policy_kernels.cu).
#include "tc_common.cuh"

namespace marl {
namespace gs {


// MUFU.EX2 / MUFU.RCP directly (the approximations __expf / __fdividef are built on, without their range handling: 4 / 6
// instructions per sigmoid / tanh instead of 11 / 13); the reciprocal's argument is 1 + 2^y >= 1, the limits come out right.
__device__ __forceinline__ float gs_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gs_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gs_sigmoid(float x) { return gs_rcp(1.f + gs_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float gs_tanh(float x) { return 1.f - 2.f * gs_rcp(1.f + gs_ex2(2.8853900817779268f * x)); }

using namespace tc;

constexpr int E = 128;
constexpr int WPLANE = 128 * 128;            // one plane of a weight k-block: 128 rows x 128 B
constexpr int WSTAGE = 2 * WPLANE;           // hi + lo
constexpr int UNIT_BYTES = 4 * WSTAGE;       // K = 128
constexpr int THREADS = 320;

template <int NR>
struct Tile {                                // [NR rows x K=128] operand: 4 k-blocks x (hi plane, lo plane)
    static constexpr int PLANE = NR * 128, KB = 2 * PLANE, BYTES = 4 * KB;
    __device__ static __forceinline__ int off(int n, int f) { return (f >> 5) * KB + n * 128 + ((((f & 31) >> 2) ^ (n & 7)) << 4) + (f & 3) * 4; }
    __device__ static __forceinline__ void store(unsigned char *X, int n, int f, float v)
    {
        float hi, lo;
        split_tf32(v, hi, lo);
        const int o = off(n, f);
        *reinterpret_cast<float *>(X + o) = hi;
        *reinterpret_cast<float *>(X + o + PLANE) = lo;
    }
    __device__ static __forceinline__ float load(const unsigned char *X, int n, int f)
    {
        const int o = off(n, f);
        return *reinterpret_cast<const float *>(X + o) + *reinterpret_cast<const float *>(X + o + PLANE);
    }
};

struct FwdArgs {
    int T;
    int64_t R;
    const float *gi, *h0, *b_hh;
    const unsigned char *packed;             // 3 units: gate g, A[m = f][k] = W_hh[g*E + f][k]
    float *out, *saves;                      // [T,R,E], [4,T,R,E] (r, z, n, hn) or null
};

struct BwdArgs {
    int T;
    int64_t R;
    const float *dout, *saves, *out, *h0;
    const unsigned char *packed;             // 3 units: gate g, A[m = f][k] = W_hh[g*E + k][f]
    float *dgi, *dgh, *dh0;                  // [T,R,3E], [T,R,3E], [R,E]
};

template <int NR, int NTILES /* state tiles in X */, int NSTAGE>
struct Smem {
    static constexpr int X_BYTES = NTILES * Tile<NR>::BYTES;
    static constexpr int BAR_OFF = X_BYTES + NSTAGE * WSTAGE;
    static constexpr int BYTES = BAR_OFF + 256;
};

// loader: streams the 3 units (12 stages) once per step
template <int NSTAGE>
__device__ __forceinline__ void loader_loop(const unsigned char *packed, int T, unsigned char *Wst, uint64_t *bars)
{
    int it = 0;
    for (int t = 0; t < T; ++t)
        for (int s12 = 0; s12 < 12; ++s12, ++it) {
            const int s = it % NSTAGE, round = it / NSTAGE;
            if (round > 0) mbar_wait(smem_u32(&bars[NSTAGE + s]), (uint32_t)((round - 1) & 1));
            mbar_expect_tx(smem_u32(&bars[s]), WSTAGE);
            bulk_g2s(smem_u32(Wst + s * WSTAGE), packed + (size_t)s12 * WSTAGE, WSTAGE, smem_u32(&bars[s]));
        }
}

// issuer: per step, 3 units; unit g multiplies state tile (g_tile ? g : 0) and writes accumulator column acc_col(g)
template <int NR, bool BWD, int NSTAGE>
__device__ __forceinline__ void issuer_loop(int T, unsigned char *X, unsigned char *Wst, uint64_t *bars, uint32_t tmem_base)
{
    const uint32_t idesc = idesc_tf32(128, NR);
    const int A_READY = 2 * NSTAGE, MMA_DONE = 2 * NSTAGE + 1;
    int it = 0;
    for (int t = 0; t < T; ++t) {
        mbar_wait(smem_u32(&bars[A_READY]), (uint32_t)(t & 1));
        fence_after();
        for (int g = 0; g < 3; ++g) {
            const uint32_t acc = tmem_base + (BWD ? 0u : (uint32_t)(g * NR));
            const uint32_t corr = acc + (BWD ? (uint32_t)NR : (uint32_t)(3 * NR));      // lo*hi + hi*lo terms, summed by the epilogue
            const unsigned char *Xg = X + (BWD ? g * Tile<NR>::BYTES : 0);
            for (int kb = 0; kb < 4; ++kb, ++it) {
                const int s = it % NSTAGE, round = it / NSTAGE;
                mbar_wait(smem_u32(&bars[s]), (uint32_t)(round & 1));
                fence_after();
                const uint32_t wa = smem_u32(Wst + s * WSTAGE), xb = smem_u32(Xg + kb * Tile<NR>::KB);
                const uint64_t a_hi0 = make_desc(wa), a_lo0 = make_desc(wa + WPLANE), b_hi0 = make_desc(xb), b_lo0 = make_desc(xb + Tile<NR>::PLANE);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t accum = (BWD ? (g | kb | kk) : (kb | kk)) ? 1u : 0u;
                        umma_tf32(acc, a_hi0 + 2 * kk, b_hi0 + 2 * kk, idesc, accum);
                        umma_tf32(corr, a_lo0 + 2 * kk, b_hi0 + 2 * kk, idesc, accum);
                        umma_tf32(corr, a_hi0 + 2 * kk, b_lo0 + 2 * kk, idesc, 1u);
                    }
                    umma_commit(smem_u32(&bars[NSTAGE + s]));
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(smem_u32(&bars[MMA_DONE]));
        __syncwarp();
    }
}

template <int NSTAGE>
__device__ __forceinline__ void setup_barriers(uint64_t *bars)
{
    for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(smem_u32(&bars[s]), 1);
        mbar_init(smem_u32(&bars[NSTAGE + s]), 1);
    }
    mbar_init(smem_u32(&bars[2 * NSTAGE]), 256);
    mbar_init(smem_u32(&bars[2 * NSTAGE + 1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ uint32_t tmem_alloc_all(int warp, uint32_t *slot)
{
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    return *slot;
}

// ------------------------------------------------------------------------------------------------------------ forward
template <int NR, int NSTAGE>
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_fwd_kernel(const __grid_constant__ FwdArgs a)
{
    using TL = Tile<NR>;
    using SM = Smem<NR, 1, NSTAGE>;
    constexpr int TMEM_COLS = NR == 64 ? 512 : 256;            // 3 gates x (main + correction) x NR columns
    constexpr int HALF = NR / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    unsigned char *X = smem, *Wst = smem + SM::X_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::BAR_OFF);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) setup_barriers<NSTAGE>(bars);
    const uint32_t tmem_base = tmem_alloc_all<TMEM_COLS>(warp, slot);
    const int64_t row0 = (int64_t)blockIdx.x * NR;
    const int A_READY = 2 * NSTAGE, MMA_DONE = 2 * NSTAGE + 1;

    if (warp < 8) {
        const int f = 32 * (warp & 3) + lane, half = warp >> 2;
        const float bh_r = a.b_hh[f], bh_z = a.b_hh[E + f], bh_n = a.b_hh[2 * E + f];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(half * HALF);
#pragma unroll 4
        for (int j = 0; j < HALF; ++j) {
            const int n = half * HALF + j;
            const int64_t row = row0 + n;
            TL::store(X, n, f, row < a.R ? a.h0[row * E + f] : 0.f);
        }
        fence_async_smem();
        fence_before();
        mbar_arrive(smem_u32(&bars[A_READY]));
        const int64_t TRE = (int64_t)a.T * a.R * E;
        static_assert(HALF == 16, "one x16 chunk per thread and step (the gi prefetch below holds 48 registers)");
        for (int t = 0; t < a.T; ++t) {
            // this step's input projections do not depend on the MMAs: their HBM latency hides behind the wait for gh
            const float *__restrict__ gi_t = a.gi + (int64_t)t * a.R * 3 * E;
            const int64_t o_t = (int64_t)t * a.R * E;
            float gir[16], giz[16], gin[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t row = row0 + half * HALF + j;
                gir[j] = giz[j] = gin[j] = 0.f;
                if (row < a.R) {
                    const float *g = gi_t + row * 3 * E + f;
                    gir[j] = __ldg(g); giz[j] = __ldg(g + E); gin[j] = __ldg(g + 2 * E);
                }
            }
            mbar_wait(smem_u32(&bars[MMA_DONE]), (uint32_t)(t & 1));
            fence_after();
            {
                uint32_t ar[16], az[16], ahn[16], cr[16], cz[16], chn[16];
                TC_TMEM_LD16(ar, taddr);
                TC_TMEM_LD16(az, taddr + (uint32_t)NR);
                TC_TMEM_LD16(ahn, taddr + (uint32_t)(2 * NR));
                TC_TMEM_LD16(cr, taddr + (uint32_t)(3 * NR));
                TC_TMEM_LD16(cz, taddr + (uint32_t)(4 * NR));
                TC_TMEM_LD16(chn, taddr + (uint32_t)(5 * NR));
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = half * HALF + j;
                    const int64_t row = row0 + n;
                    const float hp = TL::load(X, n, f);
                    // ex2-based exp and the fast reciprocal: ~40 instructions per element instead of ~210 with expf / tanhf / IEEE
                    // division (the epilogue is instruction-bound); absolute error ~2e-7, far inside the 1e-5 / 1e-4 tolerances
                    const float r = gs_sigmoid(gir[j] + ((__uint_as_float(ar[j]) + __uint_as_float(cr[j])) + bh_r));
                    const float z = gs_sigmoid(giz[j] + ((__uint_as_float(az[j]) + __uint_as_float(cz[j])) + bh_z));
                    const float hn = (__uint_as_float(ahn[j]) + __uint_as_float(chn[j])) + bh_n;
                    const float nn = gs_tanh(gin[j] + r * hn);
                    const float h = (1.f - z) * nn + z * hp;
                    TL::store(X, n, f, h);
                    if (row < a.R) {
                        const int64_t idx = o_t + row * E + f;
                        a.out[idx] = h;
                        if (a.saves) { a.saves[idx] = r; a.saves[TRE + idx] = z; a.saves[2 * TRE + idx] = nn; a.saves[3 * TRE + idx] = hn; }
                    }
                }
            }
            fence_async_smem();
            fence_before();
            mbar_arrive(smem_u32(&bars[A_READY]));
        }
    } else if (warp == 8) {
        if (lane == 0) loader_loop<NSTAGE>(a.packed, a.T, Wst, bars);
    } else {
        issuer_loop<NR, false, NSTAGE>(a.T, X, Wst, bars, tmem_base);      // whole warp, elected lane issues
    }
    fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------------------ backward
template <int NR, int NSTAGE>
__global__ void __launch_bounds__(THREADS, 1)
gru_seq_bwd_kernel(const __grid_constant__ BwdArgs a)
{
    using TL = Tile<NR>;
    using SM = Smem<NR, 3, NSTAGE>;
    constexpr int TMEM_COLS = 2 * NR;                           // main + correction
    constexpr int HALF = NR / 2;
    static_assert(HALF == 16, "one x16 TMEM load per thread and step");
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    unsigned char *X = smem, *Wst = smem + SM::X_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::BAR_OFF);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) setup_barriers<NSTAGE>(bars);
    const uint32_t tmem_base = tmem_alloc_all<TMEM_COLS>(warp, slot);
    const int64_t row0 = (int64_t)blockIdx.x * NR;
    const int A_READY = 2 * NSTAGE, MMA_DONE = 2 * NSTAGE + 1;

    if (warp < 8) {
        const int f = 32 * (warp & 3) + lane, half = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(half * HALF);
        const int64_t TRE = (int64_t)a.T * a.R * E;
        float dhz[HALF];
#pragma unroll
        for (int j = 0; j < HALF; ++j) dhz[j] = 0.f;
        for (int tt = 0; tt < a.T; ++tt) {
            const int t = a.T - 1 - tt;
            const int64_t o_t = (int64_t)t * a.R * E;
            const float *hprev = t > 0 ? a.out + (int64_t)(t - 1) * a.R * E : a.h0;
            // everything this step reads from HBM is independent of the pending MMAs: issue it all before waiting for them
            float v_do[HALF], v_r[HALF], v_z[HALF], v_n[HALF], v_hn[HALF], v_hp[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                const int64_t row = row0 + half * HALF + j;
                v_do[j] = v_r[j] = v_z[j] = v_n[j] = v_hn[j] = v_hp[j] = 0.f;
                if (row < a.R) {
                    const int64_t idx = o_t + row * E + f;
                    v_do[j] = __ldg(a.dout + idx);
                    v_r[j] = __ldg(a.saves + idx); v_z[j] = __ldg(a.saves + TRE + idx);
                    v_n[j] = __ldg(a.saves + 2 * TRE + idx); v_hn[j] = __ldg(a.saves + 3 * TRE + idx);
                    v_hp[j] = __ldg(hprev + row * E + f);
                }
            }
            uint32_t acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0u;
            if (tt > 0) {
                mbar_wait(smem_u32(&bars[MMA_DONE]), (uint32_t)((tt - 1) & 1));
                fence_after();
                uint32_t cor[16];
                TC_TMEM_LD16(acc, taddr);
                TC_TMEM_LD16(cor, taddr + (uint32_t)NR);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(cor[j]));
            }
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                const int n = half * HALF + j;
                const int64_t row = row0 + n;
                float dpr = 0.f, dpz = 0.f, dpn = 0.f, dpnr = 0.f;
                if (row < a.R) {
                    const float dh = v_do[j] + (dhz[j] + __uint_as_float(acc[j]));
                    const float r = v_r[j], z = v_z[j], nn = v_n[j], hn = v_hn[j], hp = v_hp[j];
                    const float dn = dh * (1.f - z);
                    const float dz = dh * (hp - nn);
                    dpn = dn * (1.f - nn * nn);
                    dpz = dz * z * (1.f - z);
                    dpr = dpn * hn * r * (1.f - r);
                    dpnr = dpn * r;
                    dhz[j] = dh * z;
                    float *gi = a.dgi + ((int64_t)t * a.R + row) * 3 * E + f, *gh = a.dgh + ((int64_t)t * a.R + row) * 3 * E + f;
                    gi[0] = dpr; gi[E] = dpz; gi[2 * E] = dpn;
                    gh[0] = dpr; gh[E] = dpz; gh[2 * E] = dpnr;
                }
                TL::store(X, n, f, dpr);
                TL::store(X + TL::BYTES, n, f, dpz);
                TL::store(X + 2 * TL::BYTES, n, f, dpnr);
            }
            fence_async_smem();
            fence_before();
            mbar_arrive(smem_u32(&bars[A_READY]));
        }
        // dh0 = dh_0 * z_0 + dgh_0 W_hh
        mbar_wait(smem_u32(&bars[MMA_DONE]), (uint32_t)((a.T - 1) & 1));
        fence_after();
        uint32_t acc[16], cor[16];
        TC_TMEM_LD16(acc, taddr);
        TC_TMEM_LD16(cor, taddr + (uint32_t)NR);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int64_t row = row0 + half * HALF + j;
            if (row < a.R && a.dh0) a.dh0[row * E + f] = dhz[j] + (__uint_as_float(acc[j]) + __uint_as_float(cor[j]));
        }
    } else if (warp == 8) {
        if (lane == 0) loader_loop<NSTAGE>(a.packed, a.T, Wst, bars);
    } else {
        issuer_loop<NR, true, NSTAGE>(a.T, X, Wst, bars, tmem_base);       // whole warp, elected lane issues
    }
    fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// packs A[m][k] = W[m*stride_m + k*stride_k] (m, k < 128) into 4 k-blocks x (hi plane, lo plane), SWIZZLE_128B image
__global__ void __launch_bounds__(256)
pack_kernel(const float *__restrict__ W, int64_t stride_m, int64_t stride_k, unsigned char *__restrict__ out)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= 128 * 128) return;
    const int m = idx >> 7, k = idx & 127;
    float hi, lo;
    split_tf32(W[m * stride_m + k * stride_k], hi, lo);
    const int kb = k >> 5, kk = k & 31;
    const size_t off = (size_t)kb * WSTAGE + (size_t)m * 128 + ((((kk >> 2) ^ (m & 7))) << 4) + (kk & 3) * 4;
    *reinterpret_cast<float *>(out + off) = hi;
    *reinterpret_cast<float *>(out + off + WPLANE) = lo;
}

constexpr int FWD_NR = 32, BWD_NR = 32, FWD_STAGES = 6, BWD_STAGES = 4;

}  // namespace gs
}  // namespace marl

using namespace marl;

extern "C" int64_t marl_gru_pack_bytes(void) { return 6 * (int64_t)gs::UNIT_BYTES; }

// d_packed: 1024-byte aligned, marl_gru_pack_bytes() bytes: units 0-2 serve the forward kernel, 3-5 the backward kernel
extern "C" int marl_gru_pack(const float *d_w_hh, void *d_packed, void *stream)
{
    MARL_REQUIRE(d_w_hh && d_packed && ((uintptr_t)d_packed & 1023) == 0, "marl_gru_pack: null or misaligned pointer");
    unsigned char *out = static_cast<unsigned char *>(d_packed);
    for (int g = 0; g < 3; ++g) {
        gs::pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(d_w_hh + (size_t)g * gs::E * gs::E, gs::E, 1, out + (size_t)g * gs::UNIT_BYTES);
        gs::pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(d_w_hh + (size_t)g * gs::E * gs::E, 1, gs::E, out + (size_t)(3 + g) * gs::UNIT_BYTES);
    }
    return check_launch("gru pack_kernel");
}

extern "C" int marl_gru_seq_fwd(int32_t T, int64_t R, int32_t E, const float *d_gi, const float *d_h0, const void *d_packed,
                                const float *d_b_hh, float *d_out, float *d_saves, void *stream)
{
    MARL_REQUIRE(E == gs::E, "marl_gru_seq_fwd: hidden size %d (the sequence kernel is built for 128)", E);
    MARL_REQUIRE(T > 0 && R > 0 && d_gi && d_h0 && d_packed && d_b_hh && d_out, "marl_gru_seq_fwd: bad arguments");
    gs::FwdArgs a{T, R, d_gi, d_h0, d_b_hh, static_cast<const unsigned char *>(d_packed), d_out, d_saves};
    using SM = gs::Smem<gs::FWD_NR, 1, gs::FWD_STAGES>;
    cudaError_t e = cudaFuncSetAttribute(gs::gru_seq_fwd_kernel<gs::FWD_NR, gs::FWD_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    if (e != cudaSuccess) { set_error("gru_seq_fwd_kernel: smem %d: %s", SM::BYTES, cudaGetErrorString(e)); return MARL_ECUDA; }
    const unsigned grid = (unsigned)((R + gs::FWD_NR - 1) / gs::FWD_NR);
    gs::gru_seq_fwd_kernel<gs::FWD_NR, gs::FWD_STAGES><<<grid, gs::THREADS, SM::BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("gru_seq_fwd_kernel");
}

extern "C" int marl_gru_seq_bwd(int32_t T, int64_t R, int32_t E, const float *d_dout, const float *d_saves, const float *d_out,
                                const float *d_h0, const void *d_packed, float *d_dgi, float *d_dgh, float *d_dh0, void *stream)
{
    MARL_REQUIRE(E == gs::E, "marl_gru_seq_bwd: hidden size %d (the sequence kernel is built for 128)", E);
    MARL_REQUIRE(T > 0 && R > 0 && d_dout && d_saves && d_out && d_h0 && d_packed && d_dgi && d_dgh, "marl_gru_seq_bwd: bad arguments");
    gs::BwdArgs a{T, R, d_dout, d_saves, d_out, d_h0, static_cast<const unsigned char *>(d_packed) + 3 * (size_t)gs::UNIT_BYTES, d_dgi, d_dgh, d_dh0};
    using SM = gs::Smem<gs::BWD_NR, 3, gs::BWD_STAGES>;
    cudaError_t e = cudaFuncSetAttribute(gs::gru_seq_bwd_kernel<gs::BWD_NR, gs::BWD_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    if (e != cudaSuccess) { set_error("gru_seq_bwd_kernel: smem %d: %s", SM::BYTES, cudaGetErrorString(e)); return MARL_ECUDA; }
    const unsigned grid = (unsigned)((R + gs::BWD_NR - 1) / gs::BWD_NR);
    gs::gru_seq_bwd_kernel<gs::BWD_NR, gs::BWD_STAGES><<<grid, gs::THREADS, SM::BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("gru_seq_bwd_kernel");
}
