// evader_move.cuh — the per-step part of Pursuit_Env.attacker_step (pursuit_env.py:84-100): waypoint following,
// heading, first-order-lag dynamics, occupancy test and target resampling.  Scalar work done by ONE lane per env;
// shared by the stand-alone evader kernel and the fused closed-loop rollout kernel.
#pragma once
#include "common.cuh"

namespace marl {

enum { EV_HEAP_OVERFLOW = 1, EV_PATH_OVERFLOW = 2, EV_TAPE_EXHAUSTED = 4, EV_MISSED_REPLAN = 8 };

struct EvaderRegs {
    double x, y, vx, vy;
    int tx, ty, plen, tape_pos, status;
};

// grid: this env's occupancy words (shared or global), infl: its 2-inflated map (global), path: its waypoint list
// [goal, ..., start] of which path[plen-1] is the next waypoint, tape: its candidate targets.
__device__ __forceinline__ void evader_move(const EnvDev &c, const uint32_t *grid, const uint32_t *__restrict__ infl,
                                            const int16_t *__restrict__ path, const int32_t *__restrict__ tape,
                                            int tape_len, EvaderRegs &s)
{
    int plen = s.plen;
    if (plen >= 2) {   // pursuit_env.py:84-88: drop the waypoint once it is closer than `resolution`
        const double lx = (double)path[2 * (plen - 1)], ly = (double)path[2 * (plen - 1) + 1];
        if (sqnorm2(dsub(s.x, lx), dsub(s.y, ly)) <= c.thr2_resolution_lt) --plen;
    }
    const double wx = (double)path[2 * (plen - 1)], wy = (double)path[2 * (plen - 1) + 1];
    // agent.py:261-271 waypoint2phi
    const double radius = sqrt(sqnorm2(dsub(wx, s.x), dsub(wy, s.y)));
    double phi = 0.0;
    if (!(fabs(radius) <= fmax(dmul(1e-9, fabs(radius)), 0.01))) {   // math.isclose(radius, 0.0, abs_tol=0.01)
        const double dy = dsub(wy, s.y);
        const double sg = (dy > 0.0) ? 1.0 : ((dy < 0.0) ? -1.0 : 0.0);
        phi = dmul(sg, acos(ddiv(dsub(wx, s.x), dadd(radius, 1e-3))));
    }
    const double ux = dmul(cos(phi), c.e_vmax), uy = dmul(sin(phi), c.e_vmax);
    const DivBy by_tau = make_divby(c.e_tau), by_six = make_divby(6.0);
    const double nvx = rk4_axis(s.vx, ux, by_tau, by_six, c.e_step), nvy = rk4_axis(s.vy, uy, by_tau, by_six, c.e_step);
    const double nx = dadd(s.x, dmul(nvx, c.e_step)), ny = dadd(s.y, dmul(nvy, c.e_step));
    const int xi = pyround(nx), yi = pyround(ny);
    if (xi >= 0 && xi < c.W && yi >= 0 && yi < c.H && !grid_bit(grid, c.HW, xi, yi)) {   // pursuit_env.py:96-97
        s.x = nx; s.y = ny; s.vx = nvx; s.vy = nvy;
    }
    // target reached (tested on the PROPOSED position, applied or not) -> init_target on the 2-inflated map
    if (sqnorm2(dsub((double)s.tx, nx), dsub((double)s.ty, ny)) <= c.thr2_e_capture) {
        int pos = s.tape_pos;
        for (;;) {
            if (pos >= tape_len) { s.status |= EV_TAPE_EXHAUSTED; break; }
            const int cx2 = tape[2 * pos], cy2 = tape[2 * pos + 1];
            ++pos;
            if (!grid_bit(infl, c.HW, cx2, cy2)) { s.tx = cx2; s.ty = cy2; break; }
        }
        s.tape_pos = pos;
    }
    s.plen = plen;
}

}  // namespace marl
