// envn2n_kernels.cu — the 2-D N-pursuers-vs-E-evaders particle env (environment/env_n2n/particle_env.py) on
// sm_100a.  Same lane-group mapping as the other env families (env_group.cuh): one env = G lanes of one warp with
// G = next pow2 >= max(N, E) <= 32; lane i of the group owns pursuer i AND evader i, so every cross-agent rule (kill-radius tests in
// both directions, team collisions, alive counts, the nearest-evader assignment) is warp-local: two per-warp shared tiles of
// positions, ballots and __reduce_add_sync.  fp64, operation by operation in the reference's order (never contracted); cos / sin
// are CUDA's double routines, so continuous state agrees with numpy to ~1e-15 relative while every discrete output is exact away
// from 1-ulp knife edges of the distance thresholds.
#include "env_group.cuh"

namespace marl {

struct N2nDev {
    int N, E, episode_limit;
    double p_vmax, ang_lmt, step;
    double thr2_kill, thr2_comm, thr2_sen;   // largest s with sqrt(s) <= r  (see common.cuh)
    double sen_range;                        // choose_evader compares the distance itself (strictly below)
};

struct S4 {
    double x, y, phi, v;
};

__device__ __forceinline__ S4 load_s4(const double *__restrict__ p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);   // 32-byte records
    const double2 a = q[0], b = q[1];
    return S4{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store_s4(double *__restrict__ p, const S4 &s)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(s.x, s.y);
    q[1] = make_double2(s.phi, s.v);
}
__device__ __forceinline__ double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }

// particle_env.py:41-55 / 73-85: signed, rate-limited turn from phi towards the commanded angle
__device__ __forceinline__ double turn(double ang, double phi, double ang_lmt)
{
    const double two_pi = 6.283185307179586;
    const double diff = dsub(ang, phi), ad = fabs(diff);
    double delta, sign;
    if (sgn(dmul(ang, phi)) >= 0.0) { delta = ad; sign = sgn(diff); }
    else if (ad < dsub(two_pi, ad)) { delta = ad; sign = sgn(diff); }
    else { delta = dsub(two_pi, ad); sign = -sgn(diff); }
    return dmul(sign, fmin(fmax(delta, 0.0), ang_lmt));
}
__device__ __forceinline__ double wrap_pi(double phi)
{
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    if (phi > pi) return dsub(phi, two_pi);
    if (phi < -pi) return dadd(phi, two_pi);
    return phi;
}
// Pursuer.step (particle_env.py:34-63): the heading turns whether or not the pursuer is active; only an active one moves
__device__ __forceinline__ void pursuer_step(S4 &s, bool active, int a, const N2nDev &c)
{
    const double pi = 3.141592653589793, two_pi = 6.283185307179586;
    double v = 0.0;
    if (a != 0) {
        v = c.p_vmax;
        double ang = dmul(dmul((double)a, pi), 0.25);
        if (ang > pi) ang = dsub(ang, two_pi);
        s.phi = wrap_pi(dadd(s.phi, turn(ang, s.phi, c.ang_lmt)));
    }
    if (active) {
        s.x = dadd(s.x, dmul(dmul(v, cos(s.phi)), c.step));
        s.y = dadd(s.y, dmul(dmul(v, sin(s.phi)), c.step));
        s.v = v;
    }
}
// Evader.step (particle_env.py:70-93) of an active evader: moves along the OLD heading, then turns
__device__ __forceinline__ void evader_step(S4 &s, double a, const N2nDev &c)
{
    const double pi = 3.141592653589793;
    const double d = turn(dmul(a, pi), s.phi, c.ang_lmt);
    s.x = dadd(s.x, dmul(dmul(s.v, cos(s.phi)), c.step));
    s.y = dadd(s.y, dmul(dmul(s.v, sin(s.phi)), c.step));
    s.phi = wrap_pi(dadd(s.phi, d));
}
__device__ __forceinline__ void park(S4 &s) { s.x = 1000.0; s.y = 1000.0; s.phi = 0.0; }

// uniform in [-1,1) from the counter RNG (throughput runs; tests re-implement it in numpy)
__host__ __device__ __forceinline__ double n2n_rand_pm1(uint64_t seed, int64_t agent_linear, int t)
{
    const uint64_t h = splitmix64(splitmix64(seed ^ splitmix64((uint64_t)agent_linear * 0x100000001B3ull + (uint64_t)t)) + 0x51ull);
    return (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

struct N2nArgs {
    int B, T, t0, K;
    double *p_state;
    uint8_t *p_active;
    double *e_state;
    uint8_t *e_active;
    const double *target;
    int32_t *time_step;
    const int32_t *action_tape;    // [K,B,N] or null (null => counter RNG, uniform{0..8})
    const double *e_action_tape;   // [K,B,E] or null (null with move_evader => counter RNG)
    uint64_t seed;
    int move_evader, do_step, want_obs;
    int32_t *reward_out;           // per-call API: [B,N]
    uint8_t *done_out;             // per-call API: [B]
    uint32_t *pp_out, *pe_out;     // per-call observe API: [B,N] words
    int8_t *assign_out;            // per-call observe API: [B,N]
    marl_envn2n_records rec;
};

// s_p / s_e: this group's tiles of G double4 (x, y, active as 1.0 / 0.0, unused)
template <int G>
__device__ __forceinline__ void publish(const Group<G, 1> &g, bool env_ok, const N2nDev &c, double4 *s_p, double4 *s_e, const S4 &ps,
                                        bool pa, const S4 &es, bool ea)
{
    __syncwarp();
    if (env_ok && g.gl < c.N) s_p[g.gl] = make_double4(ps.x, ps.y, pa ? 1.0 : 0.0, 0.0);
    if (env_ok && g.gl < c.E) s_e[g.gl] = make_double4(es.x, es.y, ea ? 1.0 : 0.0, 0.0);
    __syncwarp();
}

template <int G>
__global__ void __launch_bounds__(128)
envn2n_kernel(N2nDev c, N2nArgs a)
{
    using Gp = Group<G, 1>;
    __shared__ double4 s_tiles[4][2][32];
    const int warp = threadIdx.x >> 5;
    Gp g;
    g.init((int64_t)blockIdx.x * 4 + warp);
    const bool env_ok = g.env < a.B;
    const int N = c.N, E = c.E, i = g.gl;
    double4 *s_p = &s_tiles[warp][0][g.sub * G], *s_e = &s_tiles[warp][1][g.sub * G];
    const bool is_p = env_ok && i < N, is_e = env_ok && i < E;
    S4 ps = S4{0, 0, 0, 0}, es = S4{0, 0, 0, 0};
    bool pa = false, ea = false;
    double tx = 0, ty = 0;
    int ts = 0;
    if (is_p) { ps = load_s4(a.p_state + (g.env * N + i) * 4); pa = a.p_active[g.env * N + i] != 0; }
    if (is_e) { es = load_s4(a.e_state + (g.env * E + i) * 4); ea = a.e_active[g.env * E + i] != 0; }
    if (env_ok && a.target) { tx = a.target[2 * g.env]; ty = a.target[2 * g.env + 1]; }
    if (env_ok && a.time_step) ts = a.time_step[g.env];
    for (int k = 0; k < a.K; ++k) {
        const int t = a.t0 + k;
        const int64_t slab = (int64_t)t * a.B + g.env;
        if (a.want_obs) {
            // get_adj_mat for pursuer-pursuer (comm range) and pursuer-evader (sensor range), choose_evader('actor')
            publish<G>(g, env_ok, c, s_p, s_e, ps, pa, es, ea);
            uint32_t pp = 0u, pe = 0u;
            int best = -1;
            double best_d = 0.0;
            if (is_p && pa) {
                for (int j = 0; j < N; ++j) {
                    const double4 q = s_p[j];
                    pp |= (sqnorm2(dsub(ps.x, q.x), dsub(ps.y, q.y)) <= c.thr2_comm ? 1u : 0u) << j;
                }
                for (int j = 0; j < E; ++j) {
                    const double4 q = s_e[j];
                    const double s2 = sqnorm2(dsub(ps.x, q.x), dsub(ps.y, q.y));
                    pe |= (s2 <= c.thr2_sen ? 1u : 0u) << j;
                    const double d = sqrt(s2);
                    if (q.z != 0.0 && d < c.sen_range && (best < 0 || d < best_d)) { best = j; best_d = d; }
                }
            }
            if (is_p) {
                if (a.rec.pp_adj_bits) a.rec.pp_adj_bits[slab * N + i] = pp;
                if (a.rec.pe_adj_bits) a.rec.pe_adj_bits[slab * N + i] = pe;
                if (a.rec.assign) a.rec.assign[slab * N + i] = (int8_t)best;
                if (a.pp_out) a.pp_out[g.env * N + i] = pp;
                if (a.pe_out) a.pe_out[g.env * N + i] = pe;
                if (a.assign_out) a.assign_out[g.env * N + i] = (int8_t)best;
            }
        }
        // records of the pre-step state (what a replay buffer stores)
        if (is_p) {
            if (a.rec.p_state_f32)
                *reinterpret_cast<float4 *>(a.rec.p_state_f32 + (slab * N + i) * 4) = make_float4((float)ps.x, (float)ps.y, (float)ps.phi, (float)ps.v);
            if (a.rec.p_active) a.rec.p_active[slab * N + i] = pa ? 1 : 0;
        }
        if (is_e) {
            if (a.rec.e_state_f32)
                *reinterpret_cast<float4 *>(a.rec.e_state_f32 + (slab * E + i) * 4) = make_float4((float)es.x, (float)es.y, (float)es.phi, (float)es.v);
            if (a.rec.e_active) a.rec.e_active[slab * E + i] = ea ? 1 : 0;
        }
        // evader_step (particle_env.py:179-193): every ACTIVE evader takes its commanded heading
        if (a.move_evader && is_e && ea) {
            const double cmd = a.e_action_tape ? a.e_action_tape[((int64_t)k * a.B + g.env) * E + i]
                                               : n2n_rand_pm1(a.seed, -(g.env * E + i + 1), t);
            evader_step(es, cmd, c);
        }
        if (!a.do_step) continue;
        // ParticleEnv.step (particle_env.py:164-177)
        int act = 0;
        if (is_p) {
            act = a.action_tape ? a.action_tape[((int64_t)k * a.B + g.env) * N + i] : rand_action(a.seed, g.env * N + i, t);
            pursuer_step(ps, pa, act, c);
        }
        publish<G>(g, env_ok, c, s_p, s_e, ps, pa, es, ea);
        int reward = 0;
        bool p_dead = false, e_dead = false;
        if (is_p && pa) {
            int inner = 0, hit = 0;
            for (int j = 0; j < N; ++j) {
                const double4 q = s_p[j];
                inner += (q.z != 0.0 && sqnorm2(dsub(ps.x, q.x), dsub(ps.y, q.y)) <= c.thr2_kill) ? 1 : 0;
            }
            for (int j = 0; j < E; ++j) {
                const double4 q = s_e[j];
                hit += (q.z != 0.0 && sqnorm2(dsub(ps.x, q.x), dsub(ps.y, q.y)) <= c.thr2_kill) ? 1 : 0;
            }
            reward = hit - (inner - 1);                     // agent_reward (:263-276)
            p_dead = (inner + hit - 1) != 0;                // update_agent_active (:287-300)
        }
        if (is_e && ea) {
            for (int j = 0; j < N; ++j) {
                const double4 q = s_p[j];
                e_dead |= q.z != 0.0 && sqnorm2(dsub(es.x, q.x), dsub(es.y, q.y)) <= c.thr2_kill;
            }
        }
        if (p_dead) { pa = false; park(ps); }
        if (e_dead) { ea = false; park(es); }
        ts += 1;
        const bool e_at_target = is_e && sqnorm2(dsub(es.x, tx), dsub(es.y, ty)) <= c.thr2_kill;      // get_done (:264-285): every evader
        const unsigned m_target = __ballot_sync(0xffffffffu, e_at_target) & g.gmask;
        const unsigned m_p = __ballot_sync(0xffffffffu, is_p && pa) & g.gmask;
        const unsigned m_e = __ballot_sync(0xffffffffu, is_e && ea) & g.gmask;
        const bool done = m_target != 0u || m_p == 0u || m_e == 0u || ts >= c.episode_limit;
        if (is_p) {
            if (a.rec.action) a.rec.action[slab * N + i] = act;
            if (a.rec.reward) a.rec.reward[slab * N + i] = reward;
            if (a.reward_out) a.reward_out[g.env * N + i] = reward;
        }
        if (env_ok && i == 0) {
            if (a.rec.done) a.rec.done[slab] = done ? 1 : 0;
            if (a.done_out) a.done_out[g.env] = done ? 1 : 0;
        }
    }
    if (a.do_step && is_p) {
        store_s4(a.p_state + (g.env * N + i) * 4, ps);
        a.p_active[g.env * N + i] = pa ? 1 : 0;
    }
    if ((a.do_step || a.move_evader) && is_e) {
        store_s4(a.e_state + (g.env * E + i) * 4, es);
        a.e_active[g.env * E + i] = ea ? 1 : 0;
    }
    if (a.do_step && env_ok && i == 0) a.time_step[g.env] = ts;
}

static int make_n2n_dev(const marl_envn2n_params *p, N2nDev *o)
{
    MARL_REQUIRE(p != nullptr, "envn2n params is NULL");
    MARL_REQUIRE(p->N >= 1 && p->N <= 32 && p->E >= 1 && p->E <= 32, "envn2n: N=%d E=%d unsupported (1..32 each)", p->N, p->E);
    MARL_REQUIRE(p->episode_limit >= 1 && p->step_size > 0 && p->kill_radius >= 0 && p->comm_range >= 0 && p->sen_range >= 0 &&
                     p->ang_lmt >= 0, "envn2n: bad scalar configuration");
    o->N = p->N; o->E = p->E; o->episode_limit = p->episode_limit;
    o->p_vmax = p->p_vmax; o->ang_lmt = p->ang_lmt; o->step = p->step_size; o->sen_range = p->sen_range;
    o->thr2_kill = sq_threshold(p->kill_radius, false);
    o->thr2_comm = sq_threshold(p->comm_range, false);
    o->thr2_sen = sq_threshold(p->sen_range, false);
    return MARL_OK;
}

template <int G>
static int launch_g(const N2nDev &c, const N2nArgs &a, cudaStream_t st)
{
    const int64_t envs_per_block = 4 * (32 / G);
    envn2n_kernel<G><<<(unsigned)((a.B + envs_per_block - 1) / envs_per_block), 128, 0, st>>>(c, a);
    return check_launch("envn2n_kernel");
}
static int launch_n2n(const N2nDev &c, const N2nArgs &a, cudaStream_t st)
{
    const int m = c.N > c.E ? c.N : c.E;
    if (m <= 2) return launch_g<2>(c, a, st);
    if (m <= 4) return launch_g<4>(c, a, st);
    if (m <= 8) return launch_g<8>(c, a, st);
    if (m <= 16) return launch_g<16>(c, a, st);
    return launch_g<32>(c, a, st);
}

}  // namespace marl

using namespace marl;

extern "C" int marl_envn2n_step(const marl_envn2n_params *p, int32_t B, double *d_p_state, uint8_t *d_p_active, double *d_e_state,
                                uint8_t *d_e_active, const double *d_target, const int32_t *d_action, int32_t *d_time_step,
                                int32_t *d_reward, uint8_t *d_done, void *stream)
{
    N2nDev c;
    int rc = make_n2n_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_p_state && d_p_active && d_e_state && d_e_active && d_target && d_action && d_time_step && d_reward && d_done,
                 "marl_envn2n_step: null pointer or B<=0");
    N2nArgs a{};
    a.B = B; a.T = 1; a.t0 = 0; a.K = 1; a.do_step = 1;
    a.p_state = d_p_state; a.p_active = d_p_active; a.e_state = d_e_state; a.e_active = d_e_active; a.target = d_target;
    a.time_step = d_time_step; a.action_tape = d_action; a.reward_out = d_reward; a.done_out = d_done;
    return launch_n2n(c, a, (cudaStream_t)stream);
}

extern "C" int marl_envn2n_evader_step(const marl_envn2n_params *p, int32_t B, double *d_e_state, uint8_t *d_e_active,
                                       const double *d_e_action, void *stream)
{
    N2nDev c;
    int rc = make_n2n_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_e_state && d_e_active && d_e_action, "marl_envn2n_evader_step: null pointer or B<=0");
    N2nArgs a{};
    a.B = B; a.T = 1; a.t0 = 0; a.K = 1; a.move_evader = 1;
    a.e_state = d_e_state; a.e_active = d_e_active; a.e_action_tape = d_e_action;
    c.N = 0;                       // the pursuer side is not touched: no lane is a pursuer, its pointers stay null
    return launch_n2n(c, a, (cudaStream_t)stream);
}

extern "C" int marl_envn2n_observe(const marl_envn2n_params *p, int32_t B, const double *d_p_state, const uint8_t *d_p_active,
                                   const double *d_e_state, const uint8_t *d_e_active, uint32_t *d_pp_adj_bits, uint32_t *d_pe_adj_bits,
                                   int8_t *d_assign, void *stream)
{
    N2nDev c;
    int rc = make_n2n_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && d_p_state && d_p_active && d_e_state && d_e_active, "marl_envn2n_observe: null pointer or B<=0");
    N2nArgs a{};
    a.B = B; a.T = 1; a.t0 = 0; a.K = 1; a.want_obs = 1;
    a.p_state = const_cast<double *>(d_p_state); a.p_active = const_cast<uint8_t *>(d_p_active);
    a.e_state = const_cast<double *>(d_e_state); a.e_active = const_cast<uint8_t *>(d_e_active);
    a.pp_out = d_pp_adj_bits; a.pe_out = d_pe_adj_bits; a.assign_out = d_assign;
    return launch_n2n(c, a, (cudaStream_t)stream);
}

extern "C" int marl_envn2n_rollout(const marl_envn2n_params *p, int32_t B, int32_t T, int32_t t0, int32_t K, double *d_p_state,
                                   uint8_t *d_p_active, double *d_e_state, uint8_t *d_e_active, const double *d_target,
                                   int32_t *d_time_step, const int32_t *d_action_tape, const double *d_e_action_tape, uint64_t seed,
                                   const marl_envn2n_records *rec, void *stream)
{
    N2nDev c;
    int rc = make_n2n_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && K > 0 && t0 >= 0 && t0 + K <= T, "marl_envn2n_rollout: B=%d T=%d t0=%d K=%d", B, T, t0, K);
    MARL_REQUIRE(d_p_state && d_p_active && d_e_state && d_e_active && d_target && d_time_step, "marl_envn2n_rollout: null pointer");
    N2nArgs a{};
    a.B = B; a.T = T; a.t0 = t0; a.K = K; a.do_step = 1; a.move_evader = 1;
    a.p_state = d_p_state; a.p_active = d_p_active; a.e_state = d_e_state; a.e_active = d_e_active; a.target = d_target;
    a.time_step = d_time_step; a.action_tape = d_action_tape; a.e_action_tape = d_e_action_tape; a.seed = seed;
    a.want_obs = (rec && (rec->pp_adj_bits || rec->pe_adj_bits || rec->assign)) ? 1 : 0;
    if (rec) a.rec = *rec;
    return launch_n2n(c, a, (cudaStream_t)stream);
}
