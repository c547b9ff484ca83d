// env3d_kernels.cu — the 3-D particle env (environment/env_3d/particle_env.py) on sm_100a.
// Same lane-group mapping as the 2-D env (env_group.cuh): one env = G lanes of one warp (G = next pow2 >= N), N > 32 =>
// the warp owns one env and each lane carries APL agents.  All cross-agent work (pairwise kill-radius tests, alive
// counts, the evader's verdict) is warp-local: ballots, __reduce_add_sync and a per-warp shared tile of positions.
// fp64, operation by operation in the reference's order (never contracted); cos/sin are CUDA's double routines, so
// continuous state agrees with numpy to ~1e-15 relative while every discrete output is exact away from 1-ulp knife
// edges of the distance thresholds.
#include "env_group.cuh"

namespace marl {

struct Env3dDev {
    int N, NW, max_step;
    double p_vmax, e_vmax, ang_lmt, v_lmt, step;
    double thr2_kill, thr2_comm, thr2_sen;   // largest s with sqrt(s) <= r  (see common.cuh)
};

struct P6 {
    double x, y, z, phi, gamma, v;
};

__device__ __forceinline__ P6 load_p6(const double *__restrict__ p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);   // 48-byte records, 16-byte aligned
    const double2 a = q[0], b = q[1], c = q[2];
    return P6{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void store_p6(double *__restrict__ p, const P6 &s)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(s.x, s.y);
    q[1] = make_double2(s.z, s.phi);
    q[2] = make_double2(s.gamma, s.v);
}

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
__device__ __forceinline__ double signd(double v) { return (double)((v > 0.0) - (v < 0.0)); }
// np.linalg.norm of a 3-vector squared, as numpy/OpenBLAS evaluates it (FMA-contracted ddot tail)
__device__ __forceinline__ double sqnorm3(double a, double b, double c) { return __fma_rn(c, c, __fma_rn(b, b, __dmul_rn(a, a))); }

// particle_env.py:25-57 Point.step (active point)
__device__ __forceinline__ void point_step(P6 &s, double a0, double a1, double a2, double v_max, double ang_lmt, double v_lmt,
                                           double step)
{
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    const double phi = dmul(a0, pi);
    const double gamma = dmul(dmul(a1, pi), 0.5);
    const double v = dmul(dmul(dadd(a2, 1.0), 0.5), v_max);
    s.gamma = dadd(s.gamma, clipd(dsub(gamma, s.gamma), -ang_lmt, ang_lmt));
    s.v = dadd(s.v, clipd(dsub(v, s.v), -v_lmt, v_lmt));
    const double diff = dsub(phi, s.phi);
    double dphi;
    if (signd(dmul(phi, s.phi)) >= 0.0) {
        dphi = clipd(diff, -ang_lmt, ang_lmt);
    } else {
        const double d = fabs(diff);
        if (d < dsub(two_pi, d)) dphi = clipd(diff, -ang_lmt, ang_lmt);
        else dphi = dmul(clipd(dsub(two_pi, d), 0.0, ang_lmt), -signd(diff));
    }
    s.phi = dadd(s.phi, dphi);
    if (s.phi > pi) s.phi = dsub(s.phi, two_pi);
    else if (s.phi < -pi) s.phi = dadd(s.phi, two_pi);
    const double cg = cos(s.gamma), sg = sin(s.gamma);
    s.x = dadd(s.x, dmul(dmul(dmul(s.v, cg), cos(phi)), step));
    s.y = dadd(s.y, dmul(dmul(dmul(s.v, cg), sin(phi)), step));
    s.z = dadd(s.z, dmul(dmul(s.v, sg), step));
}

__device__ __forceinline__ void park(P6 &s) { s = P6{1000.0, 1000.0, 1000.0, 0.0, 0.0, 0.0}; }

// uniform in [-1,1) from the counter RNG (throughput runs; tests re-implement it in numpy)
__host__ __device__ __forceinline__ double rand_pm1(uint64_t seed, int64_t agent_linear, int t, int comp)
{
    const uint64_t h = splitmix64(splitmix64(seed ^ splitmix64((uint64_t)agent_linear * 0x100000001B3ull + (uint64_t)t)) + (uint64_t)comp);
    return (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

struct Env3dArgs {
    int B, T, t0, K;
    double *p_state;
    uint8_t *p_active;
    double *e_state;
    uint8_t *e_active;
    const double *target;
    int32_t *time_step;
    const double *action_tape;     // [K,B,N,3] or null
    const double *e_action_tape;   // [K,B,3] or null (null with move_evader => RNG)
    uint64_t seed;
    int move_evader, want_adj;
    int32_t *reward_out;           // per-call API: [B,N]
    uint8_t *done_out;             // per-call API: [B]
    marl_env3d_records rec;
};

// s_pos: tile of SLOTS double4 (x,y,z, active as 1.0/0.0)
template <int G, int APL>
__device__ __forceinline__ void adjacency_group(const Env3dDev &c, const Group<G, APL> &g, bool env_ok, double4 *s_pos,
                                                const P6 (&st)[APL], const bool (&act)[APL], const P6 &ev,
                                                uint32_t (&pp)[APL][4], bool (&pe)[APL])
{
    const int N = c.N;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) s_pos[i] = make_double4(st[a].x, st[a].y, st[a].z, act[a] ? 1.0 : 0.0);
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        pp[a][0] = pp[a][1] = pp[a][2] = pp[a][3] = 0u;
        pe[a] = false;
        if (env_ok && g.agent(a) < N && act[a]) {
            for (int k = 0; k < N; ++k) {
                const double4 q = s_pos[k];
                if (sqnorm3(dsub(st[a].x, q.x), dsub(st[a].y, q.y), dsub(st[a].z, q.z)) <= c.thr2_comm) pp[a][k >> 5] |= 1u << (k & 31);
            }
            pe[a] = sqnorm3(dsub(st[a].x, ev.x), dsub(st[a].y, ev.y), dsub(st[a].z, ev.z)) <= c.thr2_sen;
        }
    }
    __syncwarp();
}

// ParticleEnv.step for one group (particle_env.py:204-321).  Returns reward per owned agent; updates st/act/ev/e_act.
template <int G, int APL>
__device__ __forceinline__ void step3d_group(const Env3dDev &c, const Group<G, APL> &g, bool env_ok, double4 *s_pos,
                                             P6 (&st)[APL], bool (&act)[APL], const double (&cmd)[APL][3], P6 &ev, bool &e_act,
                                             int (&reward)[APL], int &alive_after)
{
    const int N = c.N;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) {
            if (act[a]) point_step(st[a], cmd[a][0], cmd[a][1], cmd[a][2], c.p_vmax, c.ang_lmt, c.v_lmt, c.step);
            s_pos[i] = make_double4(st[a].x, st[a].y, st[a].z, act[a] ? 1.0 : 0.0);
        }
    }
    __syncwarp();
    bool lane_hits_evader = false, lane_dead[APL];
    int lane_alive = 0;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        reward[a] = 0;
        lane_dead[a] = false;
        if (env_ok && g.agent(a) < N && act[a]) {
            int inner = 0;
            for (int k = 0; k < N; ++k) {
                const double4 q = s_pos[k];
                inner += (q.w != 0.0 && sqnorm3(dsub(st[a].x, q.x), dsub(st[a].y, q.y), dsub(st[a].z, q.z)) <= c.thr2_kill) ? 1 : 0;
            }
            const int hit = (e_act && sqnorm3(dsub(st[a].x, ev.x), dsub(st[a].y, ev.y), dsub(st[a].z, ev.z)) <= c.thr2_kill) ? 1 : 0;
            reward[a] = hit - (inner - 1);
            lane_dead[a] = (inner + hit - 1) != 0;
            // the evader's own verdict uses ev - q; (ev-q)^2 == (q-ev)^2 term by term, so one test serves both
            lane_hits_evader |= hit != 0;
        }
    }
    const bool e_dead = (__ballot_sync(0xffffffffu, lane_hits_evader) & g.gmask) != 0u;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        if (lane_dead[a]) { act[a] = false; park(st[a]); }
        lane_alive += (env_ok && g.agent(a) < N && act[a]) ? 1 : 0;
    }
    if (e_dead) { e_act = false; park(ev); }
    if (G == 32) alive_after = __reduce_add_sync(0xffffffffu, lane_alive);
    else {
        int v = lane_alive;
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        alive_after = v;
    }
    __syncwarp();
}

template <int G, int APL>
__global__ void __launch_bounds__(128)
env3d_kernel(Env3dDev c, Env3dArgs a)
{
    using Gp = Group<G, APL>;
    __shared__ double4 s_tiles[4][32 * APL];
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * 4 + warp);
    const bool env_ok = g.env < a.B;
    const int N = c.N, NW = c.NW;
    double4 *s_pos = &s_tiles[warp][g.sub * Gp::SLOTS];
    P6 st[APL], ev = P6{0, 0, 0, 0, 0, 0};
    bool act[APL], e_act = false;
    double tx = 0, ty = 0, tz = 0;
    int ts = 0;
    if (env_ok) {
        ev = load_p6(a.e_state + 6 * g.env);
        e_act = a.e_active[g.env] != 0;
        tx = a.target[3 * g.env]; ty = a.target[3 * g.env + 1]; tz = a.target[3 * g.env + 2];
        ts = a.time_step[g.env];
    }
#pragma unroll
    for (int q = 0; q < APL; ++q) {
        const int i = g.agent(q);
        act[q] = false;
        st[q] = P6{0, 0, 0, 0, 0, 0};
        if (env_ok && i < N) {
            st[q] = load_p6(a.p_state + (g.env * N + i) * 6);
            act[q] = a.p_active[g.env * N + i] != 0;
        }
    }
    for (int k = 0; k < a.K; ++k) {
        const int t = a.t0 + k;
        const int64_t slab = (int64_t)t * a.B;
        if (a.want_adj) {
            uint32_t pp[APL][4];
            bool pe[APL];
            adjacency_group<G, APL>(c, g, env_ok, s_pos, st, act, ev, pp, pe);
#pragma unroll
            for (int q = 0; q < APL; ++q) {
                const int i = g.agent(q);
                if (env_ok && i < N) {
                    const int64_t idx = (slab + g.env) * N + i;
                    if (a.rec.pp_adj_bits) for (int w = 0; w < NW; ++w) a.rec.pp_adj_bits[idx * NW + w] = pp[q][w];
                    if (a.rec.pe_adj) a.rec.pe_adj[idx] = pe[q] ? 1 : 0;
                }
            }
        }
        // records of the pre-step state (what a replay buffer stores)
#pragma unroll
        for (int q = 0; q < APL; ++q) {
            const int i = g.agent(q);
            if (env_ok && i < N) {
                const int64_t idx = (slab + g.env) * N + i;
                if (a.rec.p_state_f32) {
                    float2 *o = reinterpret_cast<float2 *>(a.rec.p_state_f32 + idx * 6);
                    o[0] = make_float2((float)st[q].x, (float)st[q].y);
                    o[1] = make_float2((float)st[q].z, (float)st[q].phi);
                    o[2] = make_float2((float)st[q].gamma, (float)st[q].v);
                }
                if (a.rec.active_f32) a.rec.active_f32[idx] = act[q] ? 1.f : 0.f;
            }
        }
        if (env_ok && g.gl == 0 && a.rec.e_state_f32) {
            float2 *o = reinterpret_cast<float2 *>(a.rec.e_state_f32 + (slab + g.env) * 6);
            o[0] = make_float2((float)ev.x, (float)ev.y);
            o[1] = make_float2((float)ev.z, (float)ev.phi);
            o[2] = make_float2((float)ev.gamma, (float)ev.v);
        }
        // evader move (every lane of the group computes the same thing: no broadcast needed)
        if (a.move_evader && env_ok && e_act) {
            double e0, e1, e2;
            if (a.e_action_tape) {
                const double *ea = a.e_action_tape + ((int64_t)k * a.B + g.env) * 3;
                e0 = ea[0]; e1 = ea[1]; e2 = ea[2];
            } else {
                const int64_t lin = -(g.env + 1);
                e0 = rand_pm1(a.seed, lin, t, 0); e1 = rand_pm1(a.seed, lin, t, 1); e2 = rand_pm1(a.seed, lin, t, 2);
            }
            point_step(ev, e0, e1, e2, c.e_vmax, c.ang_lmt, c.v_lmt, c.step);
        }
        double cmd[APL][3];
#pragma unroll
        for (int q = 0; q < APL; ++q) {
            const int i = g.agent(q);
            cmd[q][0] = cmd[q][1] = cmd[q][2] = 0.0;
            if (env_ok && i < N) {
                if (a.action_tape) {
                    const double *pa = a.action_tape + (((int64_t)k * a.B + g.env) * N + i) * 3;
                    cmd[q][0] = pa[0]; cmd[q][1] = pa[1]; cmd[q][2] = pa[2];
                } else {
                    const int64_t lin = g.env * N + i;
                    cmd[q][0] = rand_pm1(a.seed, lin, t, 0); cmd[q][1] = rand_pm1(a.seed, lin, t, 1); cmd[q][2] = rand_pm1(a.seed, lin, t, 2);
                }
            }
        }
        int rew[APL], alive;
        step3d_group<G, APL>(c, g, env_ok, s_pos, st, act, cmd, ev, e_act, rew, alive);
        ts += 1;
        const bool at_target = sqnorm3(dsub(ev.x, tx), dsub(ev.y, ty), dsub(ev.z, tz)) <= c.thr2_kill;
        const bool done = at_target || alive == 0 || !e_act || ts >= c.max_step;
#pragma unroll
        for (int q = 0; q < APL; ++q) {
            const int i = g.agent(q);
            if (env_ok && i < N) {
                if (a.rec.reward) a.rec.reward[(slab + g.env) * N + i] = rew[q];
                if (a.reward_out) a.reward_out[g.env * N + i] = rew[q];
            }
        }
        if (env_ok && g.gl == 0) {
            if (a.rec.done) a.rec.done[slab + g.env] = done ? 1 : 0;
            if (a.done_out) a.done_out[g.env] = done ? 1 : 0;
        }
    }
    // state back
#pragma unroll
    for (int q = 0; q < APL; ++q) {
        const int i = g.agent(q);
        if (env_ok && i < N && a.K > 0) {
            store_p6(a.p_state + (g.env * N + i) * 6, st[q]);
            a.p_active[g.env * N + i] = act[q] ? 1 : 0;
        }
    }
    if (env_ok && g.gl == 0 && (a.K > 0 || a.move_evader)) {
        store_p6(a.e_state + 6 * g.env, ev);
        a.e_active[g.env] = e_act ? 1 : 0;
        a.time_step[g.env] = ts;
    }
}

__global__ void __launch_bounds__(128)
env3d_evader_kernel(Env3dDev c, int B, double *__restrict__ e_state, const uint8_t *__restrict__ e_active,
                    const double *__restrict__ e_action)
{
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= B || !e_active[b]) return;
    P6 ev = load_p6(e_state + 6 * (int64_t)b);
    point_step(ev, e_action[3 * b], e_action[3 * b + 1], e_action[3 * b + 2], c.e_vmax, c.ang_lmt, c.v_lmt, c.step);
    store_p6(e_state + 6 * (int64_t)b, ev);
}

template <int G, int APL>
__global__ void __launch_bounds__(128)
env3d_adjacency_kernel(Env3dDev c, int B, const double *__restrict__ p_state, const uint8_t *__restrict__ p_active,
                       const double *__restrict__ e_state, uint32_t *__restrict__ pp_bits, uint8_t *__restrict__ pe,
                       float *__restrict__ pp_f32, float *__restrict__ pe_f32)
{
    using Gp = Group<G, APL>;
    __shared__ double4 s_tiles[4][32 * APL];
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * 4 + warp);
    const bool env_ok = g.env < B;
    const int N = c.N, NW = c.NW;
    P6 st[APL], ev = P6{0, 0, 0, 0, 0, 0};
    bool act[APL];
    if (env_ok) ev = load_p6(e_state + 6 * g.env);
#pragma unroll
    for (int q = 0; q < APL; ++q) {
        const int i = g.agent(q);
        act[q] = false;
        st[q] = P6{0, 0, 0, 0, 0, 0};
        if (env_ok && i < N) {
            st[q] = load_p6(p_state + (g.env * N + i) * 6);
            act[q] = p_active[g.env * N + i] != 0;
        }
    }
    uint32_t ppw[APL][4];
    bool pev[APL];
    adjacency_group<G, APL>(c, g, env_ok, &s_tiles[warp][g.sub * Gp::SLOTS], st, act, ev, ppw, pev);
#pragma unroll
    for (int q = 0; q < APL; ++q) {
        const int i = g.agent(q);
        if (env_ok && i < N) {
            const int64_t idx = g.env * N + i;
            if (pp_bits) for (int w = 0; w < NW; ++w) pp_bits[idx * NW + w] = ppw[q][w];
            if (pe) pe[idx] = pev[q] ? 1 : 0;
            if (pe_f32) pe_f32[idx] = pev[q] ? 1.f : 0.f;
            if (pp_f32) for (int k = 0; k < N; ++k) pp_f32[idx * N + k] = ((ppw[q][k >> 5] >> (k & 31)) & 1u) ? 1.f : 0.f;
        }
    }
}

static int make_env3d_dev(const marl_env3d_params *p, Env3dDev *o)
{
    MARL_REQUIRE(p != nullptr, "env3d params is NULL");
    MARL_REQUIRE(p->N >= 1 && p->N <= MARL_MAX_AGENTS, "env3d: N=%d unsupported (1..%d)", p->N, MARL_MAX_AGENTS);
    MARL_REQUIRE(p->max_step >= 1 && p->step_size > 0 && p->kill_radius >= 0 && p->comm_range >= 0 && p->sen_range >= 0 &&
                     p->ang_lmt >= 0 && p->v_lmt >= 0, "env3d: bad scalar configuration");
    o->N = p->N; o->NW = (p->N + 31) / 32; o->max_step = p->max_step;
    o->p_vmax = p->p_vmax; o->e_vmax = p->e_vmax; o->ang_lmt = p->ang_lmt; o->v_lmt = p->v_lmt; o->step = p->step_size;
    o->thr2_kill = sq_threshold(p->kill_radius, false);
    o->thr2_comm = sq_threshold(p->comm_range, false);
    o->thr2_sen = sq_threshold(p->sen_range, false);
    return MARL_OK;
}

template <typename F>
static int dispatch_n(int N, F &&f)
{
    if (N <= 2) return f(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{});
    if (N <= 4) return f(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
    if (N <= 8) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
    if (N <= 16) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
    if (N <= 32) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
    if (N <= 64) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
    return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
}

static int launch_env3d(const Env3dDev &c, const Env3dArgs &a, cudaStream_t st)
{
    return dispatch_n(c.N, [&](auto g, auto apl) {
        constexpr int G = decltype(g)::value, APL = decltype(apl)::value;
        const int64_t envs_per_block = 4 * (32 / G);
        const unsigned grid = (unsigned)((a.B + envs_per_block - 1) / envs_per_block);
        env3d_kernel<G, APL><<<grid, 128, 0, st>>>(c, a);
        return check_launch("env3d_kernel");
    });
}

}  // namespace marl

using namespace marl;

extern "C" int marl_env3d_step(const marl_env3d_params *p, int32_t B, double *d_p_state, uint8_t *d_p_active, double *d_e_state,
                               uint8_t *d_e_active, const double *d_target, const double *d_action, int32_t *d_time_step,
                               int32_t *d_reward, uint8_t *d_done, void *stream)
{
    Env3dDev c;
    int rc = make_env3d_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_p_state && d_p_active && d_e_state && d_e_active && d_target && d_action && d_time_step && d_reward && d_done,
                 "marl_env3d_step: null pointer or B<=0");
    Env3dArgs a{};
    a.B = B; a.T = 1; a.t0 = 0; a.K = 1;
    a.p_state = d_p_state; a.p_active = d_p_active; a.e_state = d_e_state; a.e_active = d_e_active; a.target = d_target;
    a.time_step = d_time_step; a.action_tape = d_action; a.reward_out = d_reward; a.done_out = d_done;
    return launch_env3d(c, a, (cudaStream_t)stream);
}

extern "C" int marl_env3d_evader_step(const marl_env3d_params *p, int32_t B, double *d_e_state, const uint8_t *d_e_active,
                                      const double *d_e_action, void *stream)
{
    Env3dDev c;
    int rc = make_env3d_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_e_state && d_e_active && d_e_action, "marl_env3d_evader_step: null pointer or B<=0");
    env3d_evader_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, B, d_e_state, d_e_active, d_e_action);
    return check_launch("env3d_evader_kernel");
}

extern "C" int marl_env3d_adjacency(const marl_env3d_params *p, int32_t B, const double *d_p_state, const uint8_t *d_p_active,
                                    const double *d_e_state, uint32_t *d_pp_adj_bits, uint8_t *d_pe_adj, float *d_pp_adj_f32,
                                    float *d_pe_adj_f32, void *stream)
{
    Env3dDev c;
    int rc = make_env3d_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_p_state && d_p_active && d_e_state, "marl_env3d_adjacency: null pointer or B<=0");
    return dispatch_n(c.N, [&](auto g, auto apl) {
        constexpr int G = decltype(g)::value, APL = decltype(apl)::value;
        const int64_t envs_per_block = 4 * (32 / G);
        const unsigned grid = (unsigned)((B + envs_per_block - 1) / envs_per_block);
        env3d_adjacency_kernel<G, APL><<<grid, 128, 0, (cudaStream_t)stream>>>(c, B, d_p_state, d_p_active, d_e_state, d_pp_adj_bits,
                                                                             d_pe_adj, d_pp_adj_f32, d_pe_adj_f32);
        return check_launch("env3d_adjacency_kernel");
    });
}

extern "C" int marl_env3d_rollout(const marl_env3d_params *p, int32_t B, int32_t T, int32_t t0, int32_t K, double *d_p_state,
                                  uint8_t *d_p_active, double *d_e_state, uint8_t *d_e_active, const double *d_target,
                                  int32_t *d_time_step, const double *d_action_tape, const double *d_e_action_tape, uint64_t seed,
                                  const marl_env3d_records *rec, void *stream)
{
    Env3dDev c;
    int rc = make_env3d_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && K > 0 && t0 >= 0 && t0 + K <= T, "marl_env3d_rollout: B=%d T=%d t0=%d K=%d", B, T, t0, K);
    MARL_REQUIRE(d_p_state && d_p_active && d_e_state && d_e_active && d_target && d_time_step, "marl_env3d_rollout: null pointer");
    Env3dArgs a{};
    a.B = B; a.T = T; a.t0 = t0; a.K = K;
    a.p_state = d_p_state; a.p_active = d_p_active; a.e_state = d_e_state; a.e_active = d_e_active; a.target = d_target;
    a.time_step = d_time_step; a.action_tape = d_action_tape; a.e_action_tape = d_e_action_tape; a.seed = seed;
    a.move_evader = 1; a.want_adj = (rec && (rec->pp_adj_bits || rec->pe_adj)) ? 1 : 0;
    if (rec) a.rec = *rec;
    return launch_env3d(c, a, (cudaStream_t)stream);
}
