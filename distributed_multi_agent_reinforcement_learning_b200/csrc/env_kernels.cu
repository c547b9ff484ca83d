// env_kernels.cu — kernels 1 (batched pursuer step), 2 (observations) and the fused env-only rollout.
// sm_100a.  HBM-bound integer/fp64 work: coalesced vectorised state traffic, shared-memory staging of the
// per-step exchange tiles and of the occupancy bitmap, grids sized to fill 148 SMs.  No tensor cores here.
#include "env_group.cuh"
#include "evader_move.cuh"
#include <type_traits>

namespace marl {

static constexpr int kWarpsPerBlock = 4;
static constexpr int kThreads = kWarpsPerBlock * 32;

// Welford update for one agent (DHGN/normalization.py:11-22 + :29-35), returns float32((x-mean)/(std+1e-8)).
__device__ __forceinline__ float welford_one(int x, long long n, double &mean, double &S, double &sd)
{
    const double xd = (double)x;
    if (n == 1) {
        mean = xd;
        sd = xd;   // `self.std = x` on the first sample
    } else {
        const double old = mean;
        const DivBy by_n = make_divby((double)n);
        mean = dadd(old, ddiv(dsub(xd, old), by_n));
        S = dadd(S, dmul(dsub(xd, old), dsub(xd, mean)));
        sd = sqrt(ddiv(S, by_n));
    }
    const double num = dsub(xd, mean);
    return num == 0.0 ? 0.0f : (float)ddiv(num, dadd(sd, 1e-8));     // +0 / positive = +0 without the library's slow path
}

// ---- kernel 1 --------------------------------------------------------------------------------------------
template <int G, int APL>
__global__ void __launch_bounds__(kThreads)
env_step_kernel(EnvDev c, int B, double *__restrict__ p_state, const double *__restrict__ e_state,
                const int32_t *__restrict__ action, const uint32_t *__restrict__ grid_bits,
                const int32_t *__restrict__ map_id, const double *__restrict__ action_table,
                int32_t *__restrict__ reward, uint8_t *__restrict__ can_apply, uint8_t *__restrict__ collision,
                int32_t *__restrict__ time_step, uint8_t *__restrict__ done)
{
    using Gp = Group<G, APL>;
    __shared__ double2 s_tiles[kWarpsPerBlock][2][32 * APL];
    __shared__ double s_table[2 * MARL_NUM_ACTIONS];
    if (threadIdx.x < 2 * MARL_NUM_ACTIONS) s_table[threadIdx.x] = action_table[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * kWarpsPerBlock + warp);
    const bool env_ok = g.env < B;
    const int N = c.N;
    AgentState st[APL];
    int act[APL], rew[APL];
    bool can[APL];
    double ex = 0.0, ey = 0.0;
    const uint32_t *grid = grid_bits;
    if (env_ok) {
        const int m = map_id ? map_id[g.env] : (int)g.env;
        grid = grid_bits + (size_t)m * c.W * c.HW;
        const double2 e = *reinterpret_cast<const double2 *>(e_state + 4 * g.env);
        ex = e.x;
        ey = e.y;
    }
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        act[a] = 8;
        st[a] = AgentState{0, 0, 0, 0};
        if (env_ok && i < N) {
            st[a] = load_state(p_state + (g.env * N + i) * 4);
            act[a] = action[g.env * N + i];
        }
    }
    bool rejected;
    double2 *s_raw = &s_tiles[warp][0][g.sub * Gp::SLOTS], *s_fin = &s_tiles[warp][1][g.sub * Gp::SLOTS];
    step_group<G, APL>(c, g, env_ok, grid, s_table, s_raw, s_fin, st, act, ex, ey, rew, can, rejected);
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) {
            const int64_t idx = g.env * N + i;
            if (can[a]) store_state(p_state + idx * 4, st[a]);
            reward[idx] = rew[a];
            can_apply[idx] = can[a] ? 1 : 0;
        }
    }
    if (env_ok && g.gl == 0) {
        if (rejected) collision[g.env] = 1;
        const int ts = time_step[g.env] + 1;
        time_step[g.env] = ts;
        done[g.env] = ts >= c.max_steps ? 1 : 0;
    }
}

// ---- observation writer shared by kernel 2 and the fused rollout -----------------------------------------
struct ObsOut {
    uint32_t *p_adj_bits;   // [B,N,NW]
    uint8_t *e_adj;         // [B,N]
    uint32_t *o_adj_bits;   // [B,N,OW]
    float *p_adj_f32;       // [B,N,N]
    float *e_adj_f32;       // [B,N,1]
    float *o_adj_f32;       // [B,N,O]
};

// s_words: per-warp staging [EPW*SLOTS][OW + NW] words.  Packed rows go out as per-agent word runs (agents of a
// warp are contiguous in memory: env-major, agent-minor); dense fp32 rows are expanded from the staged words by
// the whole warp with fully coalesced (float4 where aligned) stores.
template <int G, int APL>
__device__ __forceinline__ void write_obs(const EnvDev &c, const Group<G, APL> &g, bool env_ok, int B, const ObsOut &o,
                                          uint32_t *s_words, const uint32_t (&padj)[APL][4], const bool (&eadj)[APL],
                                          const uint32_t *(&orow)[APL])
{
    using Gp = Group<G, APL>;
    const int N = c.N, OW = c.OW, NW = c.NW, RW = OW + NW;
    const int64_t env0 = g.env - g.sub;                               // first env of this warp
    int64_t n_env64 = (int64_t)B - env0;
    if (n_env64 > Gp::EPW) n_env64 = Gp::EPW;
    if (n_env64 < 0) n_env64 = 0;
    const int n_env = (int)n_env64;
    const int n_agents = n_env * N;
    const int64_t agent0 = env0 * N;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) {
            const int la = g.sub * N + i;
            const int64_t idx = g.env * N + i;
            uint32_t *row = s_words + la * RW;
            for (int w = 0; w < OW; ++w) {
                const uint32_t v = __ldg(orow[a] + w);
                row[w] = v;
                if (o.o_adj_bits) o.o_adj_bits[idx * OW + w] = v;
            }
            for (int w = 0; w < NW; ++w) {
                row[OW + w] = padj[a][w];
                if (o.p_adj_bits) o.p_adj_bits[idx * NW + w] = padj[a][w];
            }
            if (o.e_adj) o.e_adj[idx] = eadj[a] ? 1 : 0;
            if (o.e_adj_f32) o.e_adj_f32[idx] = eadj[a] ? 1.0f : 0.0f;
        }
    }
    __syncwarp();
    if (o.o_adj_f32 && n_agents > 0) {
        const int O = c.O;
        float *dst = o.o_adj_f32 + agent0 * O;
        const int total = n_agents * O;
        if ((O & 3) == 0) {
            for (int e = g.lane * 4; e < total; e += 128) {
                const int la = e / O, k = e - la * O;
                const uint32_t *row = s_words + la * RW;
                float4 v;   // 4 consecutive bits never straddle a word because k % 4 == 0
                const uint32_t bits = row[k >> 5] >> (k & 31);
                v.x = (bits & 1u) ? 1.0f : 0.0f;
                v.y = (bits & 2u) ? 1.0f : 0.0f;
                v.z = (bits & 4u) ? 1.0f : 0.0f;
                v.w = (bits & 8u) ? 1.0f : 0.0f;
                *reinterpret_cast<float4 *>(dst + e) = v;
            }
        } else {
            for (int e = g.lane; e < total; e += 32) {
                const int la = e / O, k = e - la * O;
                dst[e] = ((s_words[la * RW + (k >> 5)] >> (k & 31)) & 1u) ? 1.0f : 0.0f;
            }
        }
    }
    if (o.p_adj_f32 && n_agents > 0) {
        float *dst = o.p_adj_f32 + agent0 * N;
        const int total = n_agents * N;
        for (int e = g.lane; e < total; e += 32) {
            const int la = e / N, k = e - la * N;
            dst[e] = ((s_words[la * RW + OW + (k >> 5)] >> (k & 31)) & 1u) ? 1.0f : 0.0f;
        }
    }
    __syncwarp();
}

// ---- kernel 2 --------------------------------------------------------------------------------------------
template <int G, int APL>
__global__ void __launch_bounds__(kThreads)
env_observe_kernel(EnvDev c, int B, const double *__restrict__ p_state, const double *__restrict__ e_state,
                   const uint32_t *__restrict__ grid_bits, const uint32_t *__restrict__ raser_bits,
                   const int32_t *__restrict__ map_id, ObsOut out)
{
    using Gp = Group<G, APL>;
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *s_pos_all = reinterpret_cast<double2 *>(smem);                                  // [warps][32*APL]
    uint32_t *s_words_all = reinterpret_cast<uint32_t *>(s_pos_all + kWarpsPerBlock * 32 * APL);
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * kWarpsPerBlock + warp);
    const bool env_ok = g.env < B;
    const int N = c.N;
    const uint32_t *grid = grid_bits, *raser = raser_bits;
    double ex = 0.0, ey = 0.0;
    if (env_ok) {
        const int m = map_id ? map_id[g.env] : (int)g.env;
        grid = grid_bits + (size_t)m * c.W * c.HW;
        raser = raser_bits + (size_t)m * c.W * c.H * c.OW;
        const double2 e = *reinterpret_cast<const double2 *>(e_state + 4 * g.env);
        ex = e.x;
        ey = e.y;
    }
    AgentState st[APL];
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        st[a] = AgentState{0, 0, 0, 0};
        if (env_ok && i < N) st[a] = load_state(p_state + (g.env * N + i) * 4);
    }
    uint32_t padj[APL][4];
    bool eadj[APL];
    const uint32_t *orow[APL];
    observe_group<G, APL>(c, g, env_ok, grid, raser, s_pos_all + warp * 32 * APL + g.sub * Gp::SLOTS, st, ex, ey, padj,
                          eadj, orow);
    uint32_t *s_words = s_words_all + (size_t)warp * (32 * APL) * (c.OW + c.NW);
    write_obs<G, APL>(c, g, env_ok, B, out, s_words, padj, eadj, orow);
}

// ---- fused env-only rollout ------------------------------------------------------------------------------
struct RolloutArgs {
    int B, T, t0, K;
    int env0;                    // global index of this sub-batch's first env (keys the action generator)
    int B_stride;                // envs per time slab of the arena (>= B; == B unless a sub-batch is being rolled)
    double *p_state;             // [B,N,4] in/out
    const double *e_tape;        // [K+1,B,4]
    const int32_t *action_tape;  // [K,B,N] or null
    uint64_t seed;
    const uint32_t *grid_bits, *raser_bits;
    const int32_t *map_id;
    const double *action_table;
    long long *wf_n;
    double *wf_mean, *wf_S, *wf_std;
    uint8_t *collision;
    int32_t *time_step;
    marl_rollout_records rec;    // time-major [T,B,N,...]
    // closed loop (CLOSED=true): the evader's per-step move is done here by the group leader; replanning is a
    // separate launch at every `difficulty` boundary, so K never crosses one.
    double *e_state;             // [B,4] in/out
    int32_t *target;             // [B,2] in/out
    const int16_t *path;         // [B,path_cap,2]
    int32_t *path_len;           // [B] in/out
    int path_cap, tape_len;
    const uint32_t *inflated_bits;
    const int32_t *target_tape;  // [B,tape_len,2]
    int32_t *tape_pos;           // [B] in/out
    int32_t *ev_status;          // [B] OR-ed
};

#ifndef MARL_ROLLOUT_MIN_BLOCKS
#define MARL_ROLLOUT_MIN_BLOCKS 5
#endif
template <int G, int APL, bool CLOSED>
__global__ void __launch_bounds__(kThreads, (APL == 1 ? MARL_ROLLOUT_MIN_BLOCKS : 1))
rollout_kernel(EnvDev c, RolloutArgs r)
{
    using Gp = Group<G, APL>;
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: tiles [warps][2][32*APL] double2 | table [18] double | grid [warps][EPW][W*HW] u32 | words [warps][32*APL][OW+NW]
    double2 *s_tiles = reinterpret_cast<double2 *>(smem);
    double *s_table = reinterpret_cast<double *>(s_tiles + kWarpsPerBlock * 2 * 32 * APL);
    uint32_t *s_grid_all = reinterpret_cast<uint32_t *>(s_table + 2 * MARL_NUM_ACTIONS + 2);
    const int grid_words = c.W * c.HW;
    uint32_t *s_words_all = s_grid_all + kWarpsPerBlock * Gp::EPW * grid_words;
    if (threadIdx.x < 2 * MARL_NUM_ACTIONS) s_table[threadIdx.x] = r.action_table[threadIdx.x];
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * kWarpsPerBlock + warp);
    const int B = r.B, N = c.N;
    const bool env_ok = g.env < B;
    // stage this group's occupancy bitmap once: it is read 9x per agent per step plus the Bresenham walks
    uint32_t *s_grid = s_grid_all + (warp * Gp::EPW + g.sub) * grid_words;
    const uint32_t *raser = r.raser_bits;
    if (env_ok) {
        const int m = r.map_id ? r.map_id[g.env] : (int)g.env;
        const uint32_t *src = r.grid_bits + (size_t)m * grid_words;
        for (int w = g.gl; w < grid_words; w += G) s_grid[w] = src[w];
        raser = r.raser_bits + (size_t)m * c.W * c.H * c.OW;
    }
    __syncthreads();
    AgentState st[APL];
    double wmean[APL], wS[APL], wsd[APL];
    long long wn = 0;
    int ts = 0;
    bool coll = false;
    if (env_ok) {
        wn = r.wf_n ? r.wf_n[g.env] : 0;
        ts = r.time_step[g.env];
    }
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        st[a] = AgentState{0, 0, 0, 0};
        wmean[a] = wS[a] = wsd[a] = 0.0;
        if (env_ok && i < N) {
            const int64_t idx = g.env * N + i;
            st[a] = load_state(r.p_state + idx * 4);
            if (r.wf_n) {
                wmean[a] = r.wf_mean[idx];
                wS[a] = r.wf_S[idx];
                wsd[a] = r.wf_std[idx];
            }
        }
    }
    double2 *s_raw = s_tiles + (warp * 2 + 0) * 32 * APL + g.sub * Gp::SLOTS;
    double2 *s_fin = s_tiles + (warp * 2 + 1) * 32 * APL + g.sub * Gp::SLOTS;
    uint32_t *s_words = s_words_all + (size_t)warp * (32 * APL) * (c.OW + c.NW);
    const int64_t BN = (int64_t)B * N, BN_rec = (int64_t)r.B_stride * N;
    double2 e_cur = make_double2(0.0, 0.0), e_vel = make_double2(0.0, 0.0);
    const double *e_src = CLOSED ? r.e_state : r.e_tape;
    if (env_ok) {
        e_cur = *reinterpret_cast<const double2 *>(e_src + 4 * g.env);
        e_vel = *reinterpret_cast<const double2 *>(e_src + 4 * g.env + 2);
    }
    EvaderRegs ev;
    ev.x = ev.y = ev.vx = ev.vy = 0.0;
    ev.tx = ev.ty = ev.plen = ev.tape_pos = ev.status = 0;
    const int leader = g.sub * G;
    const bool is_leader = env_ok && g.gl == 0;
    if (CLOSED && is_leader) {
        ev.tx = r.target[2 * g.env];
        ev.ty = r.target[2 * g.env + 1];
        ev.plen = r.path_len[g.env];
        ev.tape_pos = r.tape_pos[g.env];
    }
    for (int k = 0; k < r.K; ++k) {
        const int t = r.t0 + k;
        const int64_t tb = (int64_t)t * BN_rec;   // time-major record offset in agents
        // prefetch the evader state after this iteration's attacker_step and the actions (independent of observe)
        double2 e_nxt = e_cur, e_nvel = e_vel;
        int act[APL];
        if (CLOSED) {
            // attacker_step's move (pursuit_env.py:84-100) by the group leader, then broadcast to the group
            if (is_leader) {
                if (k > 0 && (ts % c.difficulty) == 0) ev.status |= EV_MISSED_REPLAN;
                ev.x = e_cur.x; ev.y = e_cur.y; ev.vx = e_vel.x; ev.vy = e_vel.y;
                evader_move(c, s_grid, r.inflated_bits + (size_t)(r.map_id ? r.map_id[g.env] : (int)g.env) * grid_words,
                            r.path + (size_t)g.env * r.path_cap * 2, r.target_tape + (size_t)g.env * r.tape_len * 2,
                            r.tape_len, ev);
            }
            e_nxt.x = __shfl_sync(0xffffffffu, ev.x, leader);
            e_nxt.y = __shfl_sync(0xffffffffu, ev.y, leader);
            e_nvel.x = __shfl_sync(0xffffffffu, ev.vx, leader);
            e_nvel.y = __shfl_sync(0xffffffffu, ev.vy, leader);
        } else if (env_ok) {
            const double *ep = r.e_tape + ((int64_t)(k + 1) * B + g.env) * 4;
            e_nxt = *reinterpret_cast<const double2 *>(ep);
            e_nvel = *reinterpret_cast<const double2 *>(ep + 2);
        }
#pragma unroll
        for (int a = 0; a < APL; ++a) {
            const int i = g.agent(a);
            act[a] = 8;
            if (env_ok && i < N) {
                const int64_t idx = g.env * N + i;
                act[a] = r.action_tape ? r.action_tape[(int64_t)k * BN_rec + idx] : rand_action(r.seed, idx + (int64_t)r.env0 * N, t);
            }
        }
        // -- observe (state before the step, evader before attacker_step): mappo_parallel.py:759-763
        uint32_t padj[APL][4];
        bool eadj[APL];
        const uint32_t *orow[APL];
        observe_group<G, APL>(c, g, env_ok, s_grid, raser, s_raw, st, e_cur.x, e_cur.y, padj, eadj, orow);
        ObsOut o;
        o.p_adj_bits = r.rec.p_adj_bits ? r.rec.p_adj_bits + tb * c.NW : nullptr;
        o.e_adj = r.rec.e_adj ? r.rec.e_adj + tb : nullptr;
        o.o_adj_bits = r.rec.o_adj_bits ? r.rec.o_adj_bits + tb * c.OW : nullptr;
        o.p_adj_f32 = r.rec.p_adj_f32 ? r.rec.p_adj_f32 + tb * N : nullptr;
        o.e_adj_f32 = r.rec.e_adj_f32 ? r.rec.e_adj_f32 + tb : nullptr;
        o.o_adj_f32 = r.rec.o_adj_f32 ? r.rec.o_adj_f32 + tb * c.O : nullptr;
        write_obs<G, APL>(c, g, env_ok, B, o, s_words, padj, eadj, orow);
#pragma unroll
        for (int a = 0; a < APL; ++a) {
            const int i = g.agent(a);
            if (env_ok && i < N) {
                const int64_t idx = tb + g.env * N + i;
                if (r.rec.p_state_f32)
                    *reinterpret_cast<float4 *>(r.rec.p_state_f32 + idx * 4) =
                        make_float4((float)st[a].x, (float)st[a].y, (float)st[a].vx, (float)st[a].vy);
                if (r.rec.a_n) r.rec.a_n[idx] = (float)act[a];
                if (r.rec.active) r.rec.active[idx] = 1.0f;
            }
        }
        if (env_ok && g.gl == 0 && r.rec.e_state_f32)
            *reinterpret_cast<float4 *>(r.rec.e_state_f32 + ((int64_t)t * r.B_stride + g.env) * 4) =
                make_float4((float)e_cur.x, (float)e_cur.y, (float)e_vel.x, (float)e_vel.y);
        // -- step against the evader state AFTER attacker_step: mappo_parallel.py:765,793
        int rew[APL];
        bool can[APL], rejected;
        step_group<G, APL>(c, g, env_ok, s_grid, s_table, s_raw, s_fin, st, act, e_nxt.x, e_nxt.y, rew, can, rejected);
        coll |= rejected;
        ts += 1;
        wn += 1;
#pragma unroll
        for (int a = 0; a < APL; ++a) {
            const int i = g.agent(a);
            if (env_ok && i < N) {
                const int64_t idx = tb + g.env * N + i;
                const float rn = welford_one(rew[a], wn, wmean[a], wS[a], wsd[a]);
                if (r.rec.r) r.rec.r[idx] = rn;
                if (r.rec.raw_reward) r.rec.raw_reward[idx] = rew[a];
            }
        }
        e_cur = e_nxt;
        e_vel = e_nvel;
    }
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) {
            const int64_t idx = g.env * N + i;
            store_state(r.p_state + idx * 4, st[a]);
            if (r.wf_n) {
                r.wf_mean[idx] = wmean[a];
                r.wf_S[idx] = wS[a];
                r.wf_std[idx] = wsd[a];
            }
        }
    }
    if (env_ok && g.gl == 0) {
        if (r.wf_n) r.wf_n[g.env] = wn;
        r.time_step[g.env] = ts;
        if (coll && r.collision) r.collision[g.env] = 1;
        if (CLOSED) {
            *reinterpret_cast<double2 *>(r.e_state + 4 * g.env) = e_cur;
            *reinterpret_cast<double2 *>(r.e_state + 4 * g.env + 2) = e_vel;
            r.target[2 * g.env] = ev.tx;
            r.target[2 * g.env + 1] = ev.ty;
            r.path_len[g.env] = ev.plen;
            r.tape_pos[g.env] = ev.tape_pos;
            if (r.ev_status && ev.status) r.ev_status[g.env] |= ev.status;
        }
    }
}

// ---- host-side dispatch ------------------------------------------------------------------------------------
template <typename F>
static int dispatch_group(int N, F &&f)
{
    if (N <= 2) return f(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{});
    if (N <= 4) return f(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
    if (N <= 8) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
    if (N <= 16) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
    if (N <= 32) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
    if (N <= 64) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
    return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
}

static inline int blocks_for(int B, int epw)
{
    const int64_t warps = ((int64_t)B + epw - 1) / epw;
    return (int)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
}

}  // namespace marl

using namespace marl;

extern "C" int marl_env_step(const marl_env_params *p, int32_t B, int32_t M, double *d_p_state, const double *d_e_state,
                             const int32_t *d_action, const uint32_t *d_grid_bits, const int32_t *d_map_id,
                             const double *d_action_table, int32_t *d_reward, uint8_t *d_can_apply,
                             uint8_t *d_collision, int32_t *d_time_step, uint8_t *d_done, void *stream)
{
    EnvDev c;
    int rc = make_env_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && M > 0, "marl_env_step: B=%d M=%d", B, M);
    MARL_REQUIRE(d_p_state && d_e_state && d_action && d_grid_bits && d_action_table && d_reward && d_can_apply &&
                     d_collision && d_time_step && d_done, "marl_env_step: null pointer");
    MARL_REQUIRE(d_map_id || M >= B, "marl_env_step: map_id is NULL but M < B");
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_group(c.N, [&](auto Gc, auto Ac) -> int {
        constexpr int G = decltype(Gc)::value, APL = decltype(Ac)::value;
        env_step_kernel<G, APL><<<blocks_for(B, 32 / G), kThreads, 0, s>>>(
            c, B, d_p_state, d_e_state, d_action, d_grid_bits, d_map_id, d_action_table, d_reward, d_can_apply,
            d_collision, d_time_step, d_done);
        return check_launch("env_step_kernel");
    });
}

extern "C" int marl_env_observe(const marl_env_params *p, int32_t B, int32_t M, const double *d_p_state,
                                const double *d_e_state, const uint32_t *d_grid_bits, const uint32_t *d_raser_bits,
                                const int32_t *d_map_id, uint32_t *d_p_adj_bits, uint8_t *d_e_adj,
                                uint32_t *d_o_adj_bits, float *d_p_adj_f32, float *d_e_adj_f32, float *d_o_adj_f32,
                                void *stream)
{
    EnvDev c;
    int rc = make_env_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && M > 0, "marl_env_observe: B=%d M=%d", B, M);
    MARL_REQUIRE(d_p_state && d_e_state && d_grid_bits && d_raser_bits, "marl_env_observe: null input pointer");
    MARL_REQUIRE(d_map_id || M >= B, "marl_env_observe: map_id is NULL but M < B");
    ObsOut o{d_p_adj_bits, d_e_adj, d_o_adj_bits, d_p_adj_f32, d_e_adj_f32, d_o_adj_f32};
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_group(c.N, [&](auto Gc, auto Ac) -> int {
        constexpr int G = decltype(Gc)::value, APL = decltype(Ac)::value;
        const size_t smem = (size_t)kWarpsPerBlock * 32 * APL * (sizeof(double2) + sizeof(uint32_t) * (c.OW + c.NW));
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(env_observe_kernel<G, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("env_observe: smem %zu: %s", smem, cudaGetErrorString(e)); return (int)MARL_ECUDA; }
        }
        env_observe_kernel<G, APL><<<blocks_for(B, 32 / G), kThreads, smem, s>>>(c, B, d_p_state, d_e_state, d_grid_bits,
                                                                                 d_raser_bits, d_map_id, o);
        return check_launch("env_observe_kernel");
    });
}

static int launch_rollout(const EnvDev &c, RolloutArgs &r, bool closed, cudaStream_t s)
{
    return dispatch_group(c.N, [&](auto Gc, auto Ac) -> int {
        constexpr int G = decltype(Gc)::value, APL = decltype(Ac)::value;
        constexpr int EPW = 32 / G;
        const size_t smem = (size_t)kWarpsPerBlock * 2 * 32 * APL * sizeof(double2) + (2 * MARL_NUM_ACTIONS + 2) * sizeof(double) +
                            (size_t)kWarpsPerBlock * EPW * c.W * c.HW * sizeof(uint32_t) +
                            (size_t)kWarpsPerBlock * 32 * APL * (c.OW + c.NW) * sizeof(uint32_t);
        MARL_REQUIRE(smem <= 227 * 1024, "marl_rollout: %zu B shared memory needed", smem);
        auto go = [&](auto kernel) -> int {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("rollout: smem %zu: %s", smem, cudaGetErrorString(e)); return (int)MARL_ECUDA; }
            }
            kernel<<<blocks_for(r.B, EPW), kThreads, smem, s>>>(c, r);
            return check_launch("rollout_kernel");
        };
        return closed ? go(rollout_kernel<G, APL, true>) : go(rollout_kernel<G, APL, false>);
    });
}

extern "C" int marl_rollout_steps(const marl_env_params *p, int32_t B, int32_t M, int32_t T, int32_t t0, int32_t K,
                                  double *d_p_state, const double *d_e_tape, const int32_t *d_action_tape, uint64_t seed,
                                  const uint32_t *d_grid_bits, const uint32_t *d_raser_bits, const int32_t *d_map_id,
                                  const double *d_action_table, int64_t *d_wf_n, double *d_wf_mean, double *d_wf_S,
                                  double *d_wf_std, uint8_t *d_collision, int32_t *d_time_step,
                                  const marl_rollout_records *rec, void *stream)
{
    EnvDev c;
    int rc = make_env_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && M > 0 && K > 0 && t0 >= 0 && t0 + K <= T, "marl_rollout_steps: B=%d M=%d T=%d t0=%d K=%d", B, M, T, t0, K);
    MARL_REQUIRE(d_p_state && d_e_tape && d_grid_bits && d_raser_bits && d_action_table && d_time_step && rec,
                 "marl_rollout_steps: null pointer");
    MARL_REQUIRE(d_map_id || M >= B, "marl_rollout_steps: map_id is NULL but M < B");
    MARL_REQUIRE(!d_wf_n || (d_wf_mean && d_wf_S && d_wf_std), "marl_rollout_steps: partial Welford state");
    RolloutArgs r = {};
    r.B = B; r.T = T; r.t0 = t0; r.K = K; r.B_stride = B;
    r.p_state = d_p_state; r.e_tape = d_e_tape; r.action_tape = d_action_tape; r.seed = seed;
    r.grid_bits = d_grid_bits; r.raser_bits = d_raser_bits; r.map_id = d_map_id; r.action_table = d_action_table;
    r.wf_n = (long long *)d_wf_n; r.wf_mean = d_wf_mean; r.wf_S = d_wf_S; r.wf_std = d_wf_std;
    r.collision = d_collision; r.time_step = d_time_step; r.rec = *rec;
    return launch_rollout(c, r, false, (cudaStream_t)stream);
}

extern "C" int marl_rollout_closed(const marl_env_params *p, int32_t B, int32_t B_stride, int32_t env0, int32_t M, int32_t T,
                                   int32_t t0, int32_t K, double *d_p_state, double *d_e_state, int32_t *d_target, const int16_t *d_path,
                                   int32_t *d_path_len, int32_t path_cap, const uint32_t *d_inflated_bits,
                                   const int32_t *d_target_tape, int32_t tape_len, int32_t *d_tape_pos,
                                   int32_t *d_evader_status, const int32_t *d_action_tape, uint64_t seed,
                                   const uint32_t *d_grid_bits, const uint32_t *d_raser_bits, const int32_t *d_map_id,
                                   const double *d_action_table, int64_t *d_wf_n, double *d_wf_mean, double *d_wf_S,
                                   double *d_wf_std, uint8_t *d_collision, int32_t *d_time_step,
                                   const marl_rollout_records *rec, void *stream)
{
    EnvDev c;
    int rc = make_env_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && M > 0 && K > 0 && t0 >= 0 && t0 + K <= T, "marl_rollout_closed: B=%d M=%d T=%d t0=%d K=%d", B, M, T, t0, K);
    MARL_REQUIRE((B_stride == 0 && env0 == 0) || (B_stride >= env0 + B && env0 >= 0), "marl_rollout_closed: B_stride=%d env0=%d B=%d", B_stride, env0, B);
    MARL_REQUIRE(K <= c.difficulty, "marl_rollout_closed: K=%d crosses a replanning boundary (difficulty=%d)", K, c.difficulty);
    MARL_REQUIRE(d_p_state && d_e_state && d_target && d_path && d_path_len && d_inflated_bits && d_tape_pos &&
                     (d_target_tape || tape_len == 0) && d_grid_bits && d_raser_bits && d_action_table && d_time_step && rec,
                 "marl_rollout_closed: null pointer");
    MARL_REQUIRE(path_cap >= 2 && tape_len >= 0, "marl_rollout_closed: path_cap=%d tape_len=%d", path_cap, tape_len);
    MARL_REQUIRE(d_map_id || M >= B, "marl_rollout_closed: map_id is NULL but M < B");
    MARL_REQUIRE(!d_wf_n || (d_wf_mean && d_wf_S && d_wf_std), "marl_rollout_closed: partial Welford state");
    RolloutArgs r = {};
    r.B = B; r.T = T; r.t0 = t0; r.K = K; r.B_stride = B_stride > 0 ? B_stride : B; r.env0 = env0;
    r.p_state = d_p_state; r.e_tape = nullptr; r.action_tape = d_action_tape; r.seed = seed;
    r.grid_bits = d_grid_bits; r.raser_bits = d_raser_bits; r.map_id = d_map_id; r.action_table = d_action_table;
    r.wf_n = (long long *)d_wf_n; r.wf_mean = d_wf_mean; r.wf_S = d_wf_S; r.wf_std = d_wf_std;
    r.collision = d_collision; r.time_step = d_time_step; r.rec = *rec;
    r.e_state = d_e_state; r.target = d_target; r.path = d_path; r.path_len = d_path_len; r.path_cap = path_cap;
    r.tape_len = tape_len; r.inflated_bits = d_inflated_bits; r.target_tape = d_target_tape; r.tape_pos = d_tape_pos;
    r.ev_status = d_evader_status;
    return launch_rollout(c, r, true, (cudaStream_t)stream);
}
