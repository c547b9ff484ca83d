// evader.cu — kernel 1c: the A*-driven evader (pursuit_env.py:75-102 attacker_step, agent.py:197-271,
// astar.py:26-161, Occupied_Grid_Map.py:119-191).
//
// One warp (= one CTA) per environment.  The whole search state lives in shared memory: the blocked-node
// bitmap as one 64-bit column per x, g (fp64, exact tie-breaking needs the reference's summation order),
// parent (u16) and a binary min-heap keyed by the reference's tuple order (f, x, y).  That order is total up to
// identical items, so ANY correct min-heap pops the same sequence as Python's heapq: no need to mimic its layout.
// The 8 neighbours of a popped node are relaxed by 8 lanes in parallel; heap maintenance is done by lane 0.
// Obstacle inflation / pursuer inflation / the half-open view window are separable bit operations on the columns.
#include "common.cuh"

namespace marl {

static constexpr int kHeapCap = 3072;

enum { EV_HEAP_OVERFLOW = 1, EV_PATH_OVERFLOW = 2, EV_TAPE_EXHAUSTED = 4 };

struct EvaderArgs {
    int B, path_cap, tape_len;
    double *e_state;
    const double *p_state;
    int32_t *target;
    int16_t *path;
    int32_t *path_len;
    const int32_t *time_step;
    const uint32_t *grid_bits, *inflated_bits;
    const int32_t *map_id;
    const int32_t *target_tape;
    int32_t *tape_pos;
    int32_t *status;
    double *e_tape2;   // optional [2,B,4]: evader state before / after this attacker_step
};

__device__ __forceinline__ uint64_t col_of(const uint32_t *bits, int HW, int x)
{
    uint64_t v = bits[x * HW];
    if (HW > 1) v |= (uint64_t)bits[x * HW + 1] << 32;
    return v;
}

// OR of the column shifted by -e..e (Chebyshev inflation along y), clipped to H bits
__device__ __forceinline__ uint64_t dilate_y(uint64_t c, int e, uint64_t hmask)
{
    uint64_t r = c;
    for (int d = 1; d <= e; ++d) r |= (c << d) | (c >> d);
    return r & hmask;
}

__device__ __forceinline__ bool heap_less(double fa, int na, double fb, int nb)
{   // node id = x*H1 + y with y < H1, so comparing ids == comparing (x, y) tuples
    return fa < fb || (fa == fb && na < nb);
}

__global__ void __launch_bounds__(32)
evader_kernel(EnvDev c, EvaderArgs r)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = c.W, H = c.H, W1 = W + 1, H1 = H + 1, NODES = W1 * H1;
    double *s_g = reinterpret_cast<double *>(smem);                    // [NODES]
    double *s_hf = s_g + NODES;                                        // [kHeapCap]
    uint64_t *s_stat = reinterpret_cast<uint64_t *>(s_hf + kHeapCap);  // [W1] static (inflated) columns
    uint64_t *s_mov = s_stat + W1;                                     // [W1] pursuer cells
    uint64_t *s_blk = s_mov + W1;                                      // [W1] blocked columns for A*
    uint16_t *s_par = reinterpret_cast<uint16_t *>(s_blk + W1);        // [NODES]
    uint16_t *s_hn = s_par + ((NODES + 3) & ~3);                       // [kHeapCap]
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= r.B) return;
    const int m = r.map_id ? r.map_id[b] : b;
    const uint32_t *grid = r.grid_bits + (size_t)m * W * c.HW;
    const uint64_t hmask = (H >= 64) ? ~0ull : ((1ull << H) - 1ull);
    double ex = r.e_state[4 * b], ey = r.e_state[4 * b + 1], evx = r.e_state[4 * b + 2], evy = r.e_state[4 * b + 3];
    int tx = r.target[2 * b], ty = r.target[2 * b + 1];
    int plen = r.path_len[b];
    int16_t *path = r.path + (size_t)b * r.path_cap * 2;
    int status = 0;
    const int ts = r.time_step[b];

    if (ts % c.difficulty == 0) {
        // ---- Evader.replan (agent.py:232-259) ----
        const int cx = pyround(ex), cy = pyround(ey);
        for (int x = lane; x < W1; x += 32) s_mov[x] = 0ull;
        __syncwarp();
        for (int i = lane; i < c.N; i += 32) {   // set_moving_obstacle (Occupied_Grid_Map.py:119-124)
            const int mx = pyround(r.p_state[((size_t)b * c.N + i) * 4]), my = pyround(r.p_state[((size_t)b * c.N + i) * 4 + 1]);
            atomicOr((unsigned long long *)&s_mov[mx], 1ull << my);
        }
        __syncwarp();
        int new_len = 1;
        for (int e = c.e_extend_dis; e >= 0; --e) {
            // ---- Evader.rescan (agent.py:202-230) as column algebra ----
            for (int x = lane; x < W1; x += 32) {
                uint64_t st = 0ull, pr = 0ull;
                for (int xx = max(0, x - e); xx <= min(W - 1, x + e); ++xx) {
                    st |= dilate_y(col_of(grid, c.HW, xx), e, hmask);
                    pr |= dilate_y(s_mov[xx], e, hmask);
                }
                if (x >= W) { st = 0ull; pr = 0ull; }
                // local_observation: half-open window [p-R, p+R) intersected with the disc d2 <= R^2
                uint64_t view = 0ull;
                const int R = c.e_sen_range, dx = cx - x;
                if (x < W && x >= cx - R && x < cx + R) {
                    const int rem = c.e_view2_floor - dx * dx;
                    if (rem >= 0) {
                        int k = 0;
                        while ((k + 1) * (k + 1) <= rem) ++k;
                        const int y0 = max(max(cy - k, cy - R), 0), y1 = min(min(cy + k, cy + R - 1), H - 1);
                        if (y1 >= y0) view = ((y1 - y0 + 1 >= 64) ? ~0ull : ((1ull << (y1 - y0 + 1)) - 1ull)) << y0;
                    }
                }
                s_stat[x] = st;
                s_blk[x] = st | (view & pr);
            }
            for (int n = lane; n < NODES; n += 32) s_g[n] = INFINITY;
            __syncwarp();
            // ---- AStar_2D.searching (astar.py:26-73) ----
            const int start = cx * H1 + cy, goal = tx * H1 + ty;
            int n_path = 1;
            const bool goal_blocked = (s_blk[tx] >> ty) & 1ull;
            if (!goal_blocked) {
                int hsize = 0;
                if (lane == 0) {
                    s_g[start] = 0.0;
                    s_par[start] = (uint16_t)start;
                    s_hf[0] = 0.0 + 2.5 * (double)(abs(tx - cx) + abs(ty - cy));
                    s_hn[0] = (uint16_t)start;
                }
                hsize = 1;
                __syncwarp();
                bool reached = false;
                const int ux = (lane == 0 || lane == 1 || lane == 7) ? -1 : ((lane >= 3 && lane <= 5) ? 1 : 0);
                const int uy = (lane >= 1 && lane <= 3) ? 1 : ((lane >= 5 && lane <= 7) ? -1 : 0);
                const double diag = sqrt(2.0);   // math.hypot(1, 1)
                while (hsize > 0) {
                    // pop (lane 0), broadcast
                    int s = 0;
                    if (lane == 0) {
                        s = s_hn[0];
                        --hsize;
                        if (hsize > 0) {
                            const double lf = s_hf[hsize];
                            const int ln = s_hn[hsize];
                            int i = 0;
                            for (;;) {
                                int ch = 2 * i + 1;
                                if (ch >= hsize) break;
                                double cf = s_hf[ch];
                                int cn = s_hn[ch];
                                if (ch + 1 < hsize) {
                                    const double cf2 = s_hf[ch + 1];
                                    const int cn2 = s_hn[ch + 1];
                                    if (heap_less(cf2, cn2, cf, cn)) { ++ch; cf = cf2; cn = cn2; }
                                }
                                if (!heap_less(cf, cn, lf, ln)) break;
                                s_hf[i] = cf;
                                s_hn[i] = (uint16_t)cn;
                                i = ch;
                            }
                            s_hf[i] = lf;
                            s_hn[i] = (uint16_t)ln;
                        }
                    }
                    s = __shfl_sync(0xffffffffu, s, 0);
                    hsize = __shfl_sync(0xffffffffu, hsize, 0);
                    if (s == goal) { reached = true; break; }
                    // relax the 8 neighbours (astar.py:56-65) on lanes 0..7, neighbour order of u_set
                    const int sx = s / H1, sy = s - sx * H1;
                    bool push = false;
                    double nf = 0.0;
                    int nn = 0;
                    if (lane < 8) {
                        const int nx = sx + ux, ny = sy + uy;
                        bool bad = ((s_blk[sx] >> sy) & 1ull) || nx < 0 || nx > W || ny < 0 || ny > H;
                        if (!bad) bad = (s_blk[nx] >> ny) & 1ull;
                        if (!bad) {
                            nn = nx * H1 + ny;
                            const double nc = dadd(s_g[s], (ux != 0 && uy != 0) ? diag : 1.0);
                            if (nc < s_g[nn]) {
                                s_g[nn] = nc;
                                s_par[nn] = (uint16_t)s;
                                nf = dadd(nc, dmul(2.5, (double)(abs(tx - nx) + abs(ty - ny))));
                                push = true;
                            }
                        }
                    }
                    unsigned pm = __ballot_sync(0xffffffffu, push);
                    while (pm) {
                        const int src = __ffs(pm) - 1;
                        pm &= pm - 1;
                        const double f = __shfl_sync(0xffffffffu, nf, src);
                        const int n = __shfl_sync(0xffffffffu, nn, src);
                        if (lane == 0) {
                            if (hsize >= kHeapCap) {
                                status |= EV_HEAP_OVERFLOW;
                            } else {
                                int i = hsize++;
                                while (i > 0) {
                                    const int par = (i - 1) >> 1;
                                    const double pf = s_hf[par];
                                    const int pn = s_hn[par];
                                    if (!heap_less(f, n, pf, pn)) break;
                                    s_hf[i] = pf;
                                    s_hn[i] = (uint16_t)pn;
                                    i = par;
                                }
                                s_hf[i] = f;
                                s_hn[i] = (uint16_t)n;
                            }
                        }
                    }
                    hsize = __shfl_sync(0xffffffffu, hsize, 0);
                    __syncwarp();
                }
                if (reached) {   // extract_path (astar.py:130-146): [goal, ..., start]
                    if (lane == 0) {
                        int cur = goal, n = 0;
                        path[0] = (int16_t)tx;
                        path[1] = (int16_t)ty;
                        n = 1;
                        for (;;) {
                            const int par = s_par[cur];
                            if (n < r.path_cap) {
                                path[2 * n] = (int16_t)(par / H1);
                                path[2 * n + 1] = (int16_t)(par % H1);
                            }
                            ++n;
                            cur = par;
                            if (cur == start) break;
                        }
                        if (n > r.path_cap) { status |= EV_PATH_OVERFLOW; n = r.path_cap; }
                        n_path = n;
                    }
                    n_path = __shfl_sync(0xffffffffu, n_path, 0);
                }
            }
            if (n_path < 2 && lane == 0) {
                path[0] = (int16_t)cx;
                path[1] = (int16_t)cy;
            }
            new_len = n_path;
            __syncwarp();
            if (n_path >= 2) break;
        }
        plen = new_len;
    }

    if (lane == 0) {
        // ---- waypoint following (pursuit_env.py:84-97, agent.py:261-271) ----
        if (plen >= 2) {
            const double lx = (double)path[2 * (plen - 1)], ly = (double)path[2 * (plen - 1) + 1];
            if (sqnorm2(dsub(ex, lx), dsub(ey, ly)) <= c.thr2_resolution_lt) --plen;
        }
        const double wx = (double)path[2 * (plen - 1)], wy = (double)path[2 * (plen - 1) + 1];
        const double radius = sqrt(sqnorm2(dsub(wx, ex), dsub(wy, ey)));
        double phi = 0.0;
        if (!(fabs(radius) <= fmax(dmul(1e-9, fabs(radius)), 0.01))) {   // math.isclose(radius, 0.0, abs_tol=0.01)
            const double dy = dsub(wy, ey);
            const double sg = (dy > 0.0) ? 1.0 : ((dy < 0.0) ? -1.0 : 0.0);
            phi = dmul(sg, acos(ddiv(dsub(wx, ex), dadd(radius, 1e-3))));
        }
        const double ux = dmul(cos(phi), c.e_vmax), uy = dmul(sin(phi), c.e_vmax);
        const double nvx = rk4_axis(evx, ux, c.e_tau, c.e_step), nvy = rk4_axis(evy, uy, c.e_tau, c.e_step);
        const double nx = dadd(ex, dmul(nvx, c.e_step)), ny = dadd(ey, dmul(nvy, c.e_step));
        const int xi = pyround(nx), yi = pyround(ny);
        if (xi >= 0 && xi < W && yi >= 0 && yi < H && !grid_bit(grid, c.HW, xi, yi)) {
            r.e_state[4 * b] = nx;
            r.e_state[4 * b + 1] = ny;
            r.e_state[4 * b + 2] = nvx;
            r.e_state[4 * b + 3] = nvy;
        }
        // target reached -> init_target on the 2-inflated map (pursuit_env.py:98-100, base_env.py:52-70)
        if (sqnorm2(dsub((double)tx, nx), dsub((double)ty, ny)) <= c.thr2_e_capture) {
            const uint32_t *infl = r.inflated_bits + (size_t)m * W * c.HW;
            int pos = r.tape_pos[b];
            for (;;) {
                if (pos >= r.tape_len) { status |= EV_TAPE_EXHAUSTED; break; }
                const int cx2 = r.target_tape[((size_t)b * r.tape_len + pos) * 2], cy2 = r.target_tape[((size_t)b * r.tape_len + pos) * 2 + 1];
                ++pos;
                if (!grid_bit(infl, c.HW, cx2, cy2)) { tx = cx2; ty = cy2; break; }
            }
            r.tape_pos[b] = pos;
            r.target[2 * b] = tx;
            r.target[2 * b + 1] = ty;
        }
        if (r.e_tape2) {   // the 2-slot tape the rollout kernel consumes: [0]=before, [1]=after attacker_step
            double *t0 = r.e_tape2 + 4 * (size_t)b, *t1 = r.e_tape2 + 4 * ((size_t)r.B + b);
            t0[0] = ex; t0[1] = ey; t0[2] = evx; t0[3] = evy;
            t1[0] = r.e_state[4 * b]; t1[1] = r.e_state[4 * b + 1]; t1[2] = r.e_state[4 * b + 2]; t1[3] = r.e_state[4 * b + 3];
        }
        r.path_len[b] = plen;
        if (r.status) r.status[b] |= status;
    }
}

static size_t evader_smem(const EnvDev &c)
{
    const int W1 = c.W + 1, H1 = c.H + 1, NODES = W1 * H1;
    return sizeof(double) * (NODES + kHeapCap) + sizeof(uint64_t) * 3 * W1 + sizeof(uint16_t) * (((NODES + 3) & ~3) + kHeapCap);
}

}  // namespace marl

using namespace marl;

extern "C" int marl_evader_step(const marl_env_params *p, int32_t B, int32_t M, double *d_e_state,
                                const double *d_p_state, int32_t *d_target, int16_t *d_path, int32_t *d_path_len,
                                int32_t path_cap, const int32_t *d_time_step, const uint32_t *d_grid_bits,
                                const uint32_t *d_inflated_bits, const int32_t *d_map_id, const int32_t *d_target_tape,
                                int32_t tape_len, int32_t *d_tape_pos, int32_t *d_status, double *d_e_tape2, void *stream)
{
    EnvDev c;
    int rc = make_env_dev(p, &c);
    if (rc) return rc;
    MARL_REQUIRE(B > 0 && M > 0 && path_cap >= 2 && tape_len >= 0, "marl_evader_step: B=%d M=%d path_cap=%d", B, M, path_cap);
    MARL_REQUIRE(d_e_state && d_p_state && d_target && d_path && d_path_len && d_time_step && d_grid_bits &&
                     d_inflated_bits && d_tape_pos && (d_target_tape || tape_len == 0), "marl_evader_step: null pointer");
    MARL_REQUIRE(d_map_id || M >= B, "marl_evader_step: map_id is NULL but M < B");
    if (c.H > 63 || c.W > 254 || (c.W + 1) * (c.H + 1) > 65535) {
        set_error("marl_evader_step: map %dx%d unsupported by the column-bitmap search (H <= 63, W <= 254)", c.W, c.H);
        return MARL_EUNSUPPORTED;
    }
    const size_t smem = evader_smem(c);
    MARL_REQUIRE(smem <= 227 * 1024, "marl_evader_step: %zu B shared memory needed", smem);
    cudaError_t e = cudaFuncSetAttribute(evader_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("evader: smem %zu: %s", smem, cudaGetErrorString(e)); return MARL_ECUDA; }
    EvaderArgs r;
    r.B = B; r.path_cap = path_cap; r.tape_len = tape_len;
    r.e_state = d_e_state; r.p_state = d_p_state; r.target = d_target; r.path = d_path; r.path_len = d_path_len;
    r.time_step = d_time_step; r.grid_bits = d_grid_bits; r.inflated_bits = d_inflated_bits; r.map_id = d_map_id;
    r.target_tape = d_target_tape; r.tape_pos = d_tape_pos; r.status = d_status; r.e_tape2 = d_e_tape2;
    evader_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(c, r);
    return check_launch("evader_kernel");
}
