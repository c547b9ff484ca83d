// env_group.cuh — the per-environment "lane group" machinery shared by the step / observe / fused-rollout kernels.
//
// Mapping (B200: 148 SMs, 32-wide warps).  One environment is owned by a group of G lanes of ONE warp
// (G = next power of two >= N, G <= 32), so 32/G environments share a warp; for N > 32 the whole warp owns one
// environment and every lane carries APL = ceil(N/32) pursuers (agent index = lane + 32*a).  All cross-agent
// work (pairwise distances, the ordered clip resolution, ballots) is therefore warp-local: no __syncthreads,
// no atomics.  Proposed positions are exchanged through a per-warp shared-memory tile read with broadcast
// LDS.128.
#pragma once
#include "common.cuh"

namespace marl {

template <int G, int APL>
struct Group {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "G must be a power of two <= 32");
    static_assert(APL == 1 || G == 32, "several agents per lane only when the warp owns one env");
    static constexpr int EPW = 32 / G;      // environments per warp
    static constexpr int SLOTS = G * APL;   // agent slots per environment
    int lane, sub, gl;                      // lane in warp, env slot in warp, lane in group
    unsigned gmask;                         // lanes of this group
    int64_t env;                            // environment index (may be >= B: inactive group)
    __device__ __forceinline__ void init(int64_t warp_global)
    {
        lane = threadIdx.x & 31;
        sub = lane / G;
        gl = lane % G;
        gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
        env = warp_global * EPW + sub;
    }
    __device__ __forceinline__ int agent(int a) const { return gl + a * G; }
};

struct AgentState {
    double x, y, vx, vy;
};

__device__ __forceinline__ AgentState load_state(const double *__restrict__ p)
{
    // 32-byte aligned record -> two LDG.128
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = q[0], b = q[1];
    return AgentState{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store_state(double *__restrict__ p, const AgentState &s)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(s.x, s.y);
    q[1] = make_double2(s.vx, s.vy);
}

// ---------------------------------------------------------------------------------------------------------
// Pursuit_Env.step + defender_reward for one environment group (pursuit_env.py:104-149).
// s_raw / s_fin: this group's shared tiles of SLOTS double2 (proposed positions before / after the in-place clip).
// The in-place clip of pursuit_env.py:143-145 is visible to LATER agents only.  Agent j therefore sees
//     pos(k) = k < j ? fin[k] : raw[k],      fin[k] = clip(raw[k]) if agent k passed, else raw[k]
// and fin differs from raw only for agents whose proposal left [0,W-1]x[0,H-1].  Fast path (no such agent in
// the env): everything is evaluated in parallel.  Slow path: only the out-of-range agents are resolved in index
// order (cooperatively by the group), then the others are evaluated in parallel with the k<j rule.
template <int G, int APL>
__device__ __forceinline__ void step_group(const EnvDev &c, const Group<G, APL> &g, bool env_ok,
                                           const uint32_t *__restrict__ grid, const double *__restrict__ s_table,
                                           double2 *s_raw, double2 *s_fin, AgentState (&st)[APL], const int (&act)[APL],
                                           double ex, double ey, int (&reward)[APL], bool (&can)[APL], bool &any_rejected)
{
    const int N = c.N;
    AgentState nx[APL];
    bool valid[APL], oob[APL], obst[APL];
    int inner[APL];
    bool lane_oob = false;
    const DivBy by_tau = make_divby(c.d_tau), by_six = make_divby(6.0);
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        valid[a] = env_ok && i < N;
        oob[a] = false;
        obst[a] = false;
        inner[a] = 0;
        if (valid[a]) {
            const double ux = s_table[2 * act[a]], uy = s_table[2 * act[a] + 1];
            nx[a].vx = rk4_axis(st[a].vx, ux, by_tau, by_six, c.d_step);
            nx[a].vy = rk4_axis(st[a].vy, uy, by_tau, by_six, c.d_step);
            nx[a].x = dadd(st[a].x, dmul(nx[a].vx, c.d_step));
            nx[a].y = dadd(st[a].y, dmul(nx[a].vy, c.d_step));
            obst[a] = obstacle_collision(c, grid, nx[a].x, nx[a].y);
            oob[a] = (nx[a].x < 0.0) | (nx[a].x > c.x_hi) | (nx[a].y < 0.0) | (nx[a].y > c.y_hi);
            lane_oob |= oob[a];
            s_raw[i] = make_double2(nx[a].x, nx[a].y);
            s_fin[i] = make_double2(nx[a].x, nx[a].y);
        }
    }
    __syncwarp();
    const bool group_oob = (__ballot_sync(0xffffffffu, lane_oob) & g.gmask) != 0u;
    if (!group_oob) {
        if (env_ok) {
            for (int k = 0; k < N; ++k) {
                const double2 pk = s_raw[k];
#pragma unroll
                for (int a = 0; a < APL; ++a)
                    inner[a] += (sqnorm2(dsub(pk.x, nx[a].x), dsub(pk.y, nx[a].y)) <= c.thr2_collision) ? 1 : 0;
            }
        }
    } else {
        // ordered resolution of the out-of-range agents (group-uniform branch: every lane of the group is here)
#pragma unroll
        for (int a = 0; a < APL; ++a) {
            unsigned m = (__ballot_sync(g.gmask, valid[a] && oob[a]) & g.gmask) >> (g.sub * G);
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int i = j + a * G;
                const double2 ri = s_raw[i];
                unsigned cnt = 0;
#pragma unroll
                for (int a2 = 0; a2 < APL; ++a2) {
                    const int k = g.agent(a2);
                    if (valid[a2]) {
                        const double2 pk = (k < i) ? s_fin[k] : s_raw[k];
                        cnt += (sqnorm2(dsub(pk.x, ri.x), dsub(pk.y, ri.y)) <= c.thr2_collision) ? 1u : 0u;
                    }
                }
                const unsigned total = __reduce_add_sync(g.gmask, cnt);
                const int ob_i = __shfl_sync(g.gmask, (int)obst[a], j + g.sub * G);
                const bool pass = ((int)total - 1 + ob_i) == 0;
                if (g.gl == j) {
                    inner[a] = (int)total;
                    if (pass) s_fin[i] = make_double2(fmin(fmax(ri.x, 0.0), c.x_hi), fmin(fmax(ri.y, 0.0), c.y_hi));
                }
                __syncwarp(g.gmask);
            }
        }
#pragma unroll
        for (int a = 0; a < APL; ++a) {
            if (valid[a] && !oob[a]) {
                const int i = g.agent(a);
                int cnt = 0;
                for (int k = 0; k < N; ++k) {
                    const double2 pk = (k < i) ? s_fin[k] : s_raw[k];
                    cnt += (sqnorm2(dsub(pk.x, nx[a].x), dsub(pk.y, nx[a].y)) <= c.thr2_collision) ? 1 : 0;
                }
                inner[a] = cnt;
            }
        }
    }
    bool lane_rej = false;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        reward[a] = 0;
        can[a] = false;
        if (valid[a]) {
            int r = -(inner[a] - 1) - (obst[a] ? 1 : 0);
            if (r < 0) {
                lane_rej = true;
            } else {
                const double xc = fmin(fmax(nx[a].x, 0.0), c.x_hi), yc = fmin(fmax(nx[a].y, 0.0), c.y_hi);
                r += (sqnorm2(dsub(ex, xc), dsub(ey, yc)) <= c.thr2_collision) ? 1 : 0;
                st[a] = AgentState{xc, yc, nx[a].vx, nx[a].vy};
                can[a] = true;
            }
            reward[a] = r;
        }
    }
    any_rejected = (__ballot_sync(0xffffffffu, lane_rej) & g.gmask) != 0u;
    __syncwarp();   // tiles may be reused by the caller
}

// ---------------------------------------------------------------------------------------------------------
// Pursuit_Env.communicate + sensor for one environment group (pursuit_env.py:182-209).
// s_pos: this group's tile of SLOTS double2 holding the CURRENT positions (written here).
// Outputs per owned agent: p_adj row as NW (<=4) words, e_adj bit, pointer to the raser row (OW words).
template <int G, int APL>
__device__ __forceinline__ void observe_group(const EnvDev &c, const Group<G, APL> &g, bool env_ok,
                                              const uint32_t *__restrict__ grid, const uint32_t *__restrict__ raser,
                                              double2 *s_pos, const AgentState (&st)[APL], double ex, double ey,
                                              uint32_t (&padj)[APL][4], bool (&eadj)[APL],
                                              const uint32_t *(&orow)[APL])
{
    const int N = c.N;
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        if (env_ok && i < N) s_pos[i] = make_double2(st[a].x, st[a].y);
    }
    __syncwarp();
    const int exi = pyround(ex), eyi = pyround(ey);
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = g.agent(a);
        padj[a][0] = padj[a][1] = padj[a][2] = padj[a][3] = 0u;
        eadj[a] = false;
        orow[a] = nullptr;
        if (env_ok && i < N) {
            // adj[i,j]=1 for i<=j within comm range; adj[j,1]=1 for every j because (j,j) always qualifies
            padj[a][0] = 2u;
            if constexpr (APL == 1) {
                // N <= G <= 32: one word; every lane walks all G slots (uniform trip count, unrolled, no divergence) and masks
                // k < i / k >= N afterwards (slots >= N hold stale positions: masked, never stored)
                uint32_t near = 0u;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const double2 pk = s_pos[k];
                    near |= (sqnorm2(dsub(st[a].x, pk.x), dsub(st[a].y, pk.y)) <= c.thr2_comm ? 1u : 0u) << k;
                }
                const uint32_t keep = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) & ~((1u << i) - 1u);
                padj[a][0] |= near & keep;
            } else {
                for (int k = i; k < N; ++k) {
                    const double2 pk = s_pos[k];
                    if (sqnorm2(dsub(st[a].x, pk.x), dsub(st[a].y, pk.y)) <= c.thr2_comm) padj[a][k >> 5] |= 1u << (k & 31);
                }
            }
            eadj[a] = line_of_sight(c, grid, pyround(st[a].x), pyround(st[a].y), exi, eyi);
            const int cx = __double2int_rz(st[a].x), cy = __double2int_rz(st[a].y);   // int(): truncation
            orow[a] = raser + ((size_t)cx * c.H + cy) * c.OW;
        }
    }
    __syncwarp();
}

}  // namespace marl
