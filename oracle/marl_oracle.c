/* marl_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's rollout hot path, used ONLY as the checker in tests/, in
 * __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` leg of bench.py.  The product
 * package never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against golden vectors
 * produced by executing the unmodified reference in the build container (oracle/gen_golden.py ->
 * tests/golden/ (npz files)).  The reference itself ships no tests.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 * Arithmetic is IEEE double evaluated operation by operation in the order of the Python source; compile with
 * -ffp-contract=off.  The only fused operation is the explicit fma() in orc_sqnorm2 (see there).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/marl_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* np.linalg.norm of a 2-vector == sqrt(x.dot(x)); numpy routes the dot through OpenBLAS ddot whose scalar
 * tail loop is compiled with FMA contraction: acc = fma(x0,x0,0); acc = fma(x1,x1,acc).  Verified against
 * numpy 2.3.5 in the build container on 27 931 inputs where the fused and unfused forms differ (all matched
 * the fused form).  This matters only for 1-ulp knife edges of the distance thresholds. */
static inline double orc_sqnorm2(double d0, double d1) { return fma(d1, d1, d0 * d0); }
static inline double orc_norm2(double d0, double d1) { return sqrt(orc_sqnorm2(d0, d1)); }

/* Python round(): half to even == rint() in the default rounding mode (Occupied_Grid_Map.py:65-69). */
static inline int orc_round(double v) { return (int)rint(v); }

static inline int orc_in_bound(const marl_env_params *p, double x, double y)
{ /* Occupied_Grid_Map.py:102-104: bounds tested AFTER rounding */
    int xi = orc_round(x), yi = orc_round(y);
    return xi < p->W && xi >= 0 && yi < p->H && yi >= 0;
}

static inline int orc_occupied(const marl_env_params *p, const uint8_t *grid, double x, double y)
{ /* Occupied_Grid_Map.py:86-92,109-115 */
    return grid[orc_round(x) * p->H + orc_round(y)] != 0;
}

/* agent.py:74-104 Agent.dynamic — RK4 on dv/dt=(u-v)/tau, literal operation order. out = x,y,vx,vy */
void orc_dynamic(const double s[4], double ux, double uy, double tau, double h, double out[4])
{
    double vx0 = s[2], vy0 = s[3];
    double k1 = (ux - vx0) / tau;
    double k2 = (ux - (vx0 + h * k1 / 2)) / tau;
    double k3 = (ux - (vx0 + h * k2 / 2)) / tau;
    double k4 = (ux - (vx0 + h * k3)) / tau;
    double vx = vx0 + (k1 + 2 * k2 + 2 * k3 + k4) * h / 6;
    k1 = (uy - vy0) / tau;
    k2 = (uy - (vy0 + h * k1 / 2)) / tau;
    k3 = (uy - (vy0 + h * k2 / 2)) / tau;
    k4 = (uy - (vy0 + h * k3)) / tau;
    double vy = vy0 + (k1 + 2 * k2 + 2 * k3 + k4) * h / 6;
    out[0] = s[0] + vx * h;
    out[1] = s[1] + vy * h;
    out[2] = vx;
    out[3] = vy;
}

/* pursuit_env.py:151-163 collision_detection(obstacle_type='obstacle') */
static int orc_obstacle_collision(const marl_env_params *p, const uint8_t *grid, double x, double y)
{
    for (int i = -1; i < 2; i++)
        for (int j = -1; j < 2; j++) {
            double px = x + i * p->d_collision_radius, py = y + j * p->d_collision_radius;
            if (orc_in_bound(p, px, py) && orc_occupied(p, grid, px, py)) return 1;
        }
    return 0;
}

/* pursuit_env.py:104-149 Pursuit_Env.step + defender_reward for ONE env.
 * p_state [N,4] in/out; e_state [4] (already advanced by attacker_step); grid u8 [W*H]. */
void orc_env_step(const marl_env_params *p, double *p_state, const double *e_state, const int32_t *action,
                  const uint8_t *grid, const double *action_table, int32_t *reward, uint8_t *can_apply,
                  uint8_t *collision, int32_t *time_step, uint8_t *done)
{
    int N = p->N;
    double *next = (double *)malloc(sizeof(double) * 4 * N);
    *time_step += 1;
    for (int i = 0; i < N; i++) {
        int a = action[i];
        orc_dynamic(p_state + 4 * i, action_table[2 * a], action_table[2 * a + 1], p->d_tau, p->d_step, next + 4 * i);
    }
    for (int i = 0; i < N; i++) {
        double *st = next + 4 * i;
        int r = 0, inner = 0;
        for (int j = 0; j < N; j++) /* :165-175, over the PROPOSED states incl. self */
            inner += orc_norm2(next[4 * j] - st[0], next[4 * j + 1] - st[1]) <= p->d_collision_radius;
        r -= (inner - 1);
        r -= orc_obstacle_collision(p, grid, st[0], st[1]);
        if (r < 0) { /* :138-141 */
            can_apply[i] = 0;
            *collision = 1;
            reward[i] = r;
            continue;
        }
        /* :143-147 in-place clip, visible to later agents */
        st[0] = fmin(fmax(st[0], 0.0), (double)(p->W - 1));
        st[1] = fmin(fmax(st[1], 0.0), (double)(p->H - 1));
        r += orc_norm2(e_state[0] - st[0], e_state[1] - st[1]) <= p->d_collision_radius;
        reward[i] = r;
        can_apply[i] = 1;
    }
    for (int i = 0; i < N; i++)
        if (can_apply[i]) memcpy(p_state + 4 * i, next + 4 * i, 4 * sizeof(double));
    *done = (*time_step >= p->max_steps);
    free(next);
}

/* pursuit_env.py:182-195 communicate, incl. `adj_mat[j, 1] = 1`. adj u8 [N,N] */
void orc_communicate(const marl_env_params *p, const double *p_state, uint8_t *adj)
{
    int N = p->N;
    memset(adj, 0, (size_t)N * N);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            if (i <= j && orc_norm2(p_state[4 * i] - p_state[4 * j], p_state[4 * i + 1] - p_state[4 * j + 1]) <= p->d_comm_range) {
                adj[i * N + j] = 1;
                adj[j * N + 1] = 1;
            }
}

/* agent.py:319-341 bresenham_line walked over the occupied grid (agent.py:157-169 find_attacker). */
static int orc_find_attacker(const marl_env_params *p, const uint8_t *grid, int x0, int y0, int x1, int y1)
{
    int ddx = x0 - x1, ddy = y0 - y1;
    if (sqrt((double)(ddx * ddx + ddy * ddy)) > p->d_sen_range) return 0;
    int dx = abs(x1 - x0), dy = abs(y1 - y0);
    int sx = x0 > x1 ? -1 : 1, sy = y0 > y1 ? -1 : 1;
    int err = dx - dy;
    for (;;) {
        if (grid[x0 * p->H + y0] == 1) return 0;
        if (x0 == x1 && y0 == y1) break;
        int e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x0 += sx; }
        if (e2 < dx) { err += dx; y0 += sy; }
    }
    return 1;
}

/* pursuit_env.py:197-209 sensor.  raser u8 [W*H, Ob]; o_adj u8 [N,O] (zero padded as replay_buffer.py:33,52);
 * e_adj u8 [N]. */
void orc_sensor(const marl_env_params *p, const double *p_state, const double *e_state, const uint8_t *grid,
                const uint8_t *raser, int32_t Ob, uint8_t *o_adj, uint8_t *e_adj)
{
    int N = p->N, O = p->O;
    memset(o_adj, 0, (size_t)N * O);
    for (int i = 0; i < N; i++) {
        int cx = (int)p_state[4 * i], cy = (int)p_state[4 * i + 1]; /* int(): truncation */
        const uint8_t *row = raser + ((size_t)cx * p->H + cy) * Ob;
        for (int k = 0; k < Ob && k < O; k++) o_adj[i * O + k] = row[k];
        e_adj[i] = (uint8_t)orc_find_attacker(p, grid, orc_round(p_state[4 * i]), orc_round(p_state[4 * i + 1]),
                                              orc_round(e_state[0]), orc_round(e_state[1]));
    }
}

/* pursuit_env.py:18-27 get_boundary_map; find_boundaries(mode='inner') of scikit-image 0.19.3 restated:
 * foreground cell whose edge-replicated 4-neighbourhood is not constant.  xy in np.argwhere (row-major) order.
 * Returns the number of boundary cells (may exceed cap; xy is truncated to cap). */
int32_t orc_boundary_map(const marl_env_params *p, const uint8_t *grid, uint8_t *boundary, int32_t *xy, int32_t cap)
{
    int W = p->W, H = p->H, n = 0;
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int c = grid[x * H + y] != 0, b = 0;
            if (c) {
                int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
                int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
                b = !(grid[xm * H + y] && grid[xp * H + y] && grid[x * H + ym] && grid[x * H + yp]);
            }
            boundary[x * H + y] = (uint8_t)b;
            if (b) {
                if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; }
                n++;
            }
        }
    return n;
}

/* pursuit_env.py:29-53 get_raser_map.  beam_dir f64 [beams,2] = (np.cos, np.sin)(beam*2*np.pi/beams) from the
 * caller.  cell_index i32 [W*H] = index of a boundary cell in the argwhere list (-1 elsewhere).
 * raser u8 [W*H, Ob]. */
void orc_raser_map(const marl_env_params *p, const uint8_t *boundary, const int32_t *xy, int32_t Ob,
                   const double *beam_dir, uint8_t *raser)
{
    int W = p->W, H = p->H;
    int32_t *cell_index = (int32_t *)malloc(sizeof(int32_t) * W * H);
    for (int i = 0; i < W * H; i++) cell_index[i] = -1;
    for (int k = 0; k < Ob; k++) cell_index[xy[2 * k] * H + xy[2 * k + 1]] = k;
    memset(raser, 0, (size_t)W * H * Ob);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++)
            for (int beam = 0; beam < p->sensor_beams; beam++) {
                double dx = beam_dir[2 * beam], dy = beam_dir[2 * beam + 1];
                for (int r = 0; r < p->sensor_radius; r++) {
                    double cx = x + r * dx, cy = y + r * dy;
                    if (cx < 0 || cx >= W || cy < 0 || cy >= H) break;
                    int ix = (int)cx, iy = (int)cy;
                    if (boundary[ix * H + iy]) {
                        raser[((size_t)x * H + y) * Ob + cell_index[ix * H + iy]] = 1;
                        break;
                    }
                }
            }
    free(cell_index);
}

/* Occupied_Grid_Map.py:157-166 extended_obstacles restricted to what later code can observe: the set of
 * in-bound cells within Chebyshev distance e of an obstacle cell. */
void orc_dilate(const marl_env_params *p, const uint8_t *grid, int e, uint8_t *out)
{
    int W = p->W, H = p->H;
    memset(out, 0, (size_t)W * H);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++)
            if (grid[x * H + y])
                for (int xx = x - e; xx <= x + e; xx++)
                    for (int yy = y - e; yy <= y + e; yy++)
                        if (xx >= 0 && xx < W && yy >= 0 && yy < H) out[xx * H + yy] = 1;
}

/* ---- evader ------------------------------------------------------------------------------------------ */
typedef struct { double f; int x, y; } orc_heap_item;

static inline int orc_heap_less(const orc_heap_item *a, const orc_heap_item *b)
{ /* Python tuple order (f, (x, y)) — total up to identical items, so any correct min-heap pops the same sequence */
    if (a->f != b->f) return a->f < b->f;
    if (a->x != b->x) return a->x < b->x;
    return a->y < b->y;
}

typedef struct { orc_heap_item *a; int n, cap; } orc_heap;

static void orc_heap_push(orc_heap *h, orc_heap_item it)
{
    if (h->n == h->cap) { h->cap *= 2; h->a = (orc_heap_item *)realloc(h->a, sizeof(orc_heap_item) * h->cap); }
    int i = h->n++;
    while (i > 0) {
        int par = (i - 1) >> 1;
        if (!orc_heap_less(&it, &h->a[par])) break;
        h->a[i] = h->a[par];
        i = par;
    }
    h->a[i] = it;
}

static orc_heap_item orc_heap_pop(orc_heap *h)
{
    orc_heap_item top = h->a[0], last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && orc_heap_less(&h->a[c + 1], &h->a[c])) c++;
        if (!orc_heap_less(&h->a[c], &last)) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = last;
    return top;
}

/* astar.py:26-161 AStar_2D.searching (weighted, e=2.5, manhattan).  Node domain is [0,W]x[0,H] INCLUSIVE
 * (astar.py:109-113 tests `> width`).  blocked u8 [(W+1)*(H+1)] (index x*(H+1)+y).
 * path i16 [cap,2] receives [goal, ..., start]; returns its length (1 => [start]); *n_closed = len(CLOSED). */
int32_t orc_astar(const marl_env_params *p, const uint8_t *blocked, int sx, int sy, int gx, int gy,
                  int16_t *path, int32_t cap, int32_t *n_closed)
{
    const int W1 = p->W + 1, H1 = p->H + 1;
    static const int ux[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, uy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    if (n_closed) *n_closed = 0;
    path[0] = (int16_t)sx; path[1] = (int16_t)sy;
    if (gx >= 0 && gx < W1 && gy >= 0 && gy < H1 && blocked[gx * H1 + gy]) return 1; /* astar.py:46-47 */
    /* g over a window that also covers the one-cell ring outside the domain (neighbours get g=inf entries) */
    const int GW = W1 + 2, GH = H1 + 2;
    double *g = (double *)malloc(sizeof(double) * GW * GH);
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * GW * GH);
    for (int i = 0; i < GW * GH; i++) { g[i] = INFINITY; parent[i] = -1; }
#define GI(x, y) (((x) + 1) * GH + (y) + 1)
    orc_heap h; h.n = 0; h.cap = 1024; h.a = (orc_heap_item *)malloc(sizeof(orc_heap_item) * h.cap);
    g[GI(sx, sy)] = 0.0;
    parent[GI(sx, sy)] = GI(sx, sy);
    orc_heap_item it = {0.0 + 2.5 * (double)(abs(gx - sx) + abs(gy - sy)), sx, sy};
    orc_heap_push(&h, it);
    int closed = 0, reached = 0;
    const double diag = sqrt(2.0); /* math.hypot(1,1) */
    while (h.n > 0) {
        orc_heap_item s = orc_heap_pop(&h);
        closed++;
        if (s.x == gx && s.y == gy) { reached = 1; break; }
        int s_bad = blocked[s.x * H1 + s.y]; /* popped nodes are always inside the domain */
        for (int k = 0; k < 8; k++) {
            int nx = s.x + ux[k], ny = s.y + uy[k];
            int bad = s_bad || nx < 0 || nx > p->W || ny < 0 || ny > p->H;
            if (!bad) bad = blocked[nx * H1 + ny];
            double c = bad ? INFINITY : ((ux[k] != 0 && uy[k] != 0) ? diag : 1.0);
            double nc = g[GI(s.x, s.y)] + c;
            if (nc < g[GI(nx, ny)]) {
                g[GI(nx, ny)] = nc;
                parent[GI(nx, ny)] = GI(s.x, s.y);
                orc_heap_item ni = {nc + 2.5 * (double)(abs(gx - nx) + abs(gy - ny)), nx, ny};
                orc_heap_push(&h, ni);
            }
        }
    }
    if (n_closed) *n_closed = closed;
    int32_t n = 1;
    if (reached) { /* astar.py:130-146 extract_path */
        n = 0;
        int cur = GI(gx, gy);
        path[0] = (int16_t)gx; path[1] = (int16_t)gy; n = 1;
        for (;;) {
            int par = parent[cur];
            if (n < cap) { path[2 * n] = (int16_t)(par / GH - 1); path[2 * n + 1] = (int16_t)(par % GH - 1); }
            n++;
            cur = par;
            if (cur == GI(sx, sy)) break;
        }
        if (n > cap) n = -n; /* overflow marker */
    }
#undef GI
    free(g); free(parent); free(h.a);
    return n;
}

/* agent.py:202-230 Evader.rescan reduced to the membership set A* consumes
 * (dynamic_map.obstacles + dynamic_map.ex_obstacles).  blocked u8 [(W+1)*(H+1)]. */
void orc_rescan(const marl_env_params *p, const uint8_t *grid, int e, const double *p_state, int px, int py,
                uint8_t *blocked)
{
    int W = p->W, H = p->H, H1 = H + 1, N = p->N, R = p->e_sen_range;
    uint8_t *stat = (uint8_t *)malloc((size_t)W * H), *pred = (uint8_t *)malloc((size_t)W * H);
    orc_dilate(p, grid, e, stat);
    memcpy(pred, stat, (size_t)W * H);
    for (int i = 0; i < N; i++) { /* set_moving_obstacle + extended_moving_obstacles (Occupied_Grid_Map.py:119-135) */
        int mx = orc_round(p_state[4 * i]), my = orc_round(p_state[4 * i + 1]);
        for (int xx = mx - e; xx <= mx + e; xx++)
            for (int yy = my - e; yy <= my + e; yy++)
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) pred[xx * H + yy] = 1;
    }
    memset(blocked, 0, (size_t)(W + 1) * H1);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) blocked[x * H1 + y] = stat[x * H + y];
    /* local_observation: half-open window [p-R, p+R) (Occupied_Grid_Map.py:186-187) */
    for (int x = px - R; x < px + R; x++)
        for (int y = py - R; y < py + R; y++) {
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
            int ddx = px - x, ddy = py - y;
            if (sqrt((double)(ddx * ddx + ddy * ddy)) > (double)R) continue;
            if (!stat[x * H + y] && pred[x * H + y]) blocked[x * H1 + y] = 1;
        }
    free(stat); free(pred);
}

/* agent.py:232-259 Evader.replan: try extend_dis, extend_dis-1, ..., 0; keep the first path with >= 2 nodes. */
static _Thread_local int64_t orc_tl_pops = 0;   /* diagnostics: |CLOSED| summed over the searches of this thread */
static int64_t *orc_pop_sink = NULL;            /* optional per-env accumulator set by orc_set_pop_sink */
void orc_set_pop_sink(int64_t *sink) { orc_pop_sink = sink; }

int32_t orc_replan(const marl_env_params *p, const uint8_t *grid, const double *e_state, const double *p_state,
                   const int32_t *target, int16_t *path, int32_t cap)
{
    int cx = orc_round(e_state[0]), cy = orc_round(e_state[1]);
    uint8_t *blocked = (uint8_t *)malloc((size_t)(p->W + 1) * (p->H + 1));
    int32_t n = 1;
    for (int e = p->e_extend_dis; e >= 0; e--) {
        orc_rescan(p, grid, e, p_state, cx, cy, blocked);
        int32_t nc = 0;
        n = orc_astar(p, blocked, cx, cy, target[0], target[1], path, cap, &nc);
        orc_tl_pops += nc;
        if (n >= 2 || n < 0) break;
    }
    free(blocked);
    return n;
}

/* pursuit_env.py:75-102 attacker_step for ONE env (single evader).
 * time_step is the value BEFORE this iteration's Pursuit_Env.step.  cos_sin_acos: optional callbacks are not
 * used — libm cos/sin/acos stand in for np.cos/np.sin/np.arccos (<= 1 ulp apart; float tolerance 1e-5 applies).
 * target tape: i32 [tape_len,2]; *tape_pos advances over rejected candidates too (base_env.py:63-70).
 * Returns 0, or -1 if the path overflowed cap, -2 if the tape ran out. */
int32_t orc_evader_step(const marl_env_params *p, double *e_state, const double *p_state, int32_t *target,
                        int16_t *path, int32_t *path_len, int32_t cap, int32_t time_step, const uint8_t *grid,
                        const uint8_t *inflated, const int32_t *tape, int32_t tape_len, int32_t *tape_pos)
{
    if (time_step % p->difficulty == 0) {
        int32_t n = orc_replan(p, grid, e_state, p_state, target, path, cap);
        if (n < 0) return -1;
        *path_len = n;
    }
    int n = *path_len;
    if (n >= 2) {
        double lx = path[2 * (n - 1)], ly = path[2 * (n - 1) + 1];
        if (orc_norm2(e_state[0] - lx, e_state[1] - ly) < p->resolution) { n--; *path_len = n; }
    }
    double wx = path[2 * (n - 1)], wy = path[2 * (n - 1) + 1];
    /* agent.py:261-271 waypoint2phi */
    double radius = orc_norm2(wx - e_state[0], wy - e_state[1]);
    double phi;
    if (fabs(radius) <= fmax(1e-9 * fabs(radius), 0.01)) phi = 0.0; /* math.isclose(radius, 0.0, abs_tol=0.01) */
    else {
        double dy = wy - e_state[1];
        double sg = (dy > 0) - (dy < 0);
        phi = sg * acos((wx - e_state[0]) / (radius + 1e-3));
    }
    double nxt[4];
    orc_dynamic(e_state, cos(phi) * p->e_vmax, sin(phi) * p->e_vmax, p->e_tau, p->e_step, nxt);
    if (orc_in_bound(p, nxt[0], nxt[1]) && !orc_occupied(p, grid, nxt[0], nxt[1])) memcpy(e_state, nxt, sizeof(nxt));
    if (orc_norm2((double)target[0] - nxt[0], (double)target[1] - nxt[1]) <= p->e_collision_radius) {
        for (;;) { /* base_env.py:52-70 init_target on self.inflated_map */
            if (*tape_pos >= tape_len) return -2;
            int tx = tape[2 * *tape_pos], ty = tape[2 * *tape_pos + 1];
            (*tape_pos)++;
            if (!inflated[tx * p->H + ty]) { target[0] = tx; target[1] = ty; break; }
        }
    }
    return 0;
}

/* DHGN/normalization.py:4-35 for ONE env: x = int rewards [N]; out f32 [N] (mappo_parallel.py:797 cast). */
void orc_welford(int32_t N, const int32_t *x, int64_t *n, double *mean, double *S, double *std, float *out, int update)
{
    if (update) {
        *n += 1;
        if (*n == 1) {
            for (int i = 0; i < N; i++) { mean[i] = (double)x[i]; std[i] = (double)x[i]; }
        } else {
            for (int i = 0; i < N; i++) {
                double xi = (double)x[i], old = mean[i];
                mean[i] = old + (xi - old) / (double)*n;
                S[i] = S[i] + (xi - old) * (xi - mean[i]);
                std[i] = sqrt(S[i] / (double)*n);
            }
        }
    }
    for (int i = 0; i < N; i++) out[i] = (float)(((double)x[i] - mean[i]) / (std[i] + 1e-8));
}

/* DHGN/mappo_parallel.py:643-658 GAE + adv-norm, fp32 like torch (statistics accumulated in double).
 * r, active [B,T,N]; v [B,T+1,N]; gl = float32(gamma*lamda) computed by the caller in double then cast. */
void orc_gae(int32_t B, int32_t T, int32_t N, const float *r, const float *v, const float *active, float gamma,
             float gl, int use_adv_norm, float *adv, float *v_target)
{
    for (int b = 0; b < B; b++)
        for (int n = 0; n < N; n++) {
            float gae = 0.0f;
            for (int t = T - 1; t >= 0; t--) {
                size_t i = ((size_t)b * T + t) * N + n;
                size_t iv = ((size_t)b * (T + 1) + t) * N + n;
                float delta = (r[i] + gamma * v[iv + N] - v[iv]) * active[i];
                gae = delta + gl * gae;
                adv[i] = gae;
                v_target[i] = gae + v[iv];
            }
        }
    if (use_adv_norm) {
        size_t M = (size_t)B * T * N;
        double s = 0;
        for (size_t i = 0; i < M; i++) s += adv[i];
        double mean = s / (double)M, q = 0;
        for (size_t i = 0; i < M; i++) { double d = adv[i] - mean; q += d * d; }
        float fm = (float)mean, fs = (float)sqrt(q / (double)(M - 1));
        for (size_t i = 0; i < M; i++) adv[i] = (adv[i] - fm) / (fs + 1e-5f) * active[i];
    }
}

/* ---- batched drivers (cpu_baseline / --impl reference legs of bench.py; all host threads via OpenMP) ---- */

/* One rollout iteration of the env-only hot loop for B envs: observe -> evader (tape) -> step -> welford.
 * Same per-env semantics as the reference loop body DHGN/mappo_parallel.py:758-801 minus the network.
 * Dense per-env maps: grid u8 [M,W*H], raser u8 [M,W*H,Ob_stride] with per-map Ob in ob_count. */
void orc_rollout_iteration(const marl_env_params *p, int32_t B, double *p_state, const double *e_before,
                           const double *e_after, const int32_t *action, const uint8_t *grid, const uint8_t *raser,
                           const int32_t *ob_count, int32_t ob_stride, const int32_t *map_id,
                           const double *action_table, uint8_t *p_adj, uint8_t *o_adj, uint8_t *e_adj,
                           int32_t *reward, uint8_t *can_apply, uint8_t *collision, int32_t *time_step,
                           uint8_t *done, int64_t *wf_n, double *wf_mean, double *wf_S, double *wf_std,
                           float *r_norm)
{
    int N = p->N, O = p->O, WH = p->W * p->H;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        int m = map_id ? map_id[b] : b;
        const uint8_t *g = grid + (size_t)m * WH;
        const uint8_t *rs = raser + (size_t)m * WH * ob_stride;
        double *ps = p_state + (size_t)b * N * 4;
        orc_communicate(p, ps, p_adj + (size_t)b * N * N);
        /* raser rows are ob_stride wide here: walk with the stride of this pool */
        {
            uint8_t *oa = o_adj + (size_t)b * N * O;
            memset(oa, 0, (size_t)N * O);
            for (int i = 0; i < N; i++) {
                int cx = (int)ps[4 * i], cy = (int)ps[4 * i + 1];
                const uint8_t *row = rs + ((size_t)cx * p->H + cy) * ob_stride;
                int ob = ob_count[m] < O ? ob_count[m] : O;
                memcpy(oa + (size_t)i * O, row, (size_t)ob);
                e_adj[(size_t)b * N + i] = (uint8_t)orc_find_attacker(p, g, orc_round(ps[4 * i]), orc_round(ps[4 * i + 1]),
                                                                     orc_round(e_before[4 * b]), orc_round(e_before[4 * b + 1]));
            }
        }
        orc_env_step(p, ps, e_after + 4 * (size_t)b, action + (size_t)b * N, g, action_table,
                     reward + (size_t)b * N, can_apply + (size_t)b * N, collision + b, time_step + b, done + b);
        orc_welford(N, reward + (size_t)b * N, wf_n + b, wf_mean + (size_t)b * N, wf_S + (size_t)b * N,
                    wf_std + (size_t)b * N, r_norm + (size_t)b * N, 1);
    }
}


/* Closed env-only rollout iteration WITH the A* evader (the reference loop body DHGN/mappo_parallel.py:758-801 minus
 * the network): observe -> attacker_step -> step -> reward-norm, for B envs on all host threads.
 * Evader state per env: e_state [B,4], target [B,2], path i16 [B,cap,2], path_len [B]; target tape [B,tape_len,2]. */
void orc_rollout_iteration_closed(const marl_env_params *p, int32_t B, double *p_state, double *e_state,
                                  int32_t *target, int16_t *path, int32_t *path_len, int32_t cap,
                                  const int32_t *action, const uint8_t *grid, const uint8_t *inflated,
                                  const uint8_t *raser, const int32_t *ob_count, int32_t ob_stride,
                                  const int32_t *map_id, const double *action_table, const int32_t *tape,
                                  int32_t tape_len, int32_t *tape_pos, uint8_t *p_adj, uint8_t *o_adj, uint8_t *e_adj,
                                  int32_t *reward, uint8_t *can_apply, uint8_t *collision, int32_t *time_step,
                                  uint8_t *done, int64_t *wf_n, double *wf_mean, double *wf_S, double *wf_std,
                                  float *r_norm, int32_t *status)
{
    int N = p->N, O = p->O, WH = p->W * p->H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < B; b++) {
        int m = map_id ? map_id[b] : b;
        const uint8_t *g = grid + (size_t)m * WH;
        const uint8_t *rs = raser + (size_t)m * WH * ob_stride;
        double *ps = p_state + (size_t)b * N * 4;
        double *es = e_state + (size_t)b * 4;
        orc_communicate(p, ps, p_adj + (size_t)b * N * N);
        uint8_t *oa = o_adj + (size_t)b * N * O;
        memset(oa, 0, (size_t)N * O);
        for (int i = 0; i < N; i++) {
            int cx = (int)ps[4 * i], cy = (int)ps[4 * i + 1];
            const uint8_t *row = rs + ((size_t)cx * p->H + cy) * ob_stride;
            int ob = ob_count[m] < O ? ob_count[m] : O;
            memcpy(oa + (size_t)i * O, row, (size_t)ob);
            e_adj[(size_t)b * N + i] = (uint8_t)orc_find_attacker(p, g, orc_round(ps[4 * i]), orc_round(ps[4 * i + 1]),
                                                                 orc_round(es[0]), orc_round(es[1]));
        }
        orc_tl_pops = 0;
        int32_t rc = orc_evader_step(p, es, ps, target + 2 * (size_t)b, path + (size_t)b * cap * 2, path_len + b, cap,
                                     time_step[b], g, inflated + (size_t)m * WH, tape + (size_t)b * tape_len * 2, tape_len,
                                     tape_pos + b);
        if (rc && status) status[b] = rc;
        if (orc_pop_sink) orc_pop_sink[b] += orc_tl_pops;
        orc_env_step(p, ps, es, action + (size_t)b * N, g, action_table, reward + (size_t)b * N,
                     can_apply + (size_t)b * N, collision + b, time_step + b, done + b);
        orc_welford(N, reward + (size_t)b * N, wf_n + b, wf_mean + (size_t)b * N, wf_S + (size_t)b * N,
                    wf_std + (size_t)b * N, r_norm + (size_t)b * N, 1);
    }
}

/* communicate + sensor for B envs (the observation the policy consumes before it acts), all host threads. */
void orc_observe_batch(const marl_env_params *p, int32_t B, const double *p_state, const double *e_state,
                       const uint8_t *grid, const uint8_t *raser, const int32_t *ob_count, int32_t ob_stride,
                       const int32_t *map_id, uint8_t *p_adj, uint8_t *o_adj, uint8_t *e_adj)
{
    int N = p->N, O = p->O, WH = p->W * p->H;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        int m = map_id ? map_id[b] : b;
        const uint8_t *g = grid + (size_t)m * WH;
        const uint8_t *rs = raser + (size_t)m * WH * ob_stride;
        const double *ps = p_state + (size_t)b * N * 4, *es = e_state + (size_t)b * 4;
        orc_communicate(p, ps, p_adj + (size_t)b * N * N);
        uint8_t *oa = o_adj + (size_t)b * N * O;
        memset(oa, 0, (size_t)N * O);
        for (int i = 0; i < N; i++) {
            int cx = (int)ps[4 * i], cy = (int)ps[4 * i + 1];
            int ob = ob_count[m] < O ? ob_count[m] : O;
            memcpy(oa + (size_t)i * O, rs + ((size_t)cx * p->H + cy) * ob_stride, (size_t)ob);
            e_adj[(size_t)b * N + i] = (uint8_t)orc_find_attacker(p, g, orc_round(ps[4 * i]), orc_round(ps[4 * i + 1]),
                                                                 orc_round(es[0]), orc_round(es[1]));
        }
    }
}

int32_t orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* =====================================================================================================
 * 3-D particle env (environment/env_3d/particle_env.py) — SURVEY §8 a23.  Pinned by tests/golden/env3d_*.npz
 * (oracle/gen_golden_env3d.py executes the unmodified reference).
 * ===================================================================================================== */

/* np.linalg.norm of a 3-vector: sqrt(ddot) with the FMA-contracted scalar tail (fma(c,c,fma(b,b,a*a)));
 * checked against numpy 2.3.5 in the build container on 50 000 random vectors (0 mismatches; the unfused form
 * mismatches on 5 321 of them). */
static inline double orc_norm3(double a, double b, double c) { return sqrt(fma(c, c, fma(b, b, a * a))); }
static inline double orc_clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
static inline double orc_sign(double v) { return (v > 0.0) - (v < 0.0); }

/* particle_env.py:25-57 Point.step for an ACTIVE point; s = (x,y,z,phi,gamma,v) */
void orc_point_step(double *s, const double *a, double v_max, double ang_lmt, double v_lmt, double step_size)
{
    const double pi = 3.141592653589793;
    double phi = a[0] * pi;
    double gamma = a[1] * pi / 2;
    double v = (a[2] + 1) / 2 * v_max;
    double delta_gamma = orc_clip(gamma - s[4], -ang_lmt, ang_lmt);
    double delta_v = orc_clip(v - s[5], -v_lmt, v_lmt);
    s[4] += delta_gamma;
    s[5] += delta_v;
    double delta_phi;
    if (orc_sign(phi * s[3]) >= 0) {
        delta_phi = orc_clip(phi - s[3], -ang_lmt, ang_lmt);
    } else {
        double d = fabs(phi - s[3]);
        if (d < 2 * pi - d) {
            delta_phi = orc_clip(phi - s[3], -ang_lmt, ang_lmt);
        } else {
            delta_phi = 2 * pi - d;
            double sign = -orc_sign(phi - s[3]);
            delta_phi = orc_clip(delta_phi, 0, ang_lmt) * sign;
        }
    }
    s[3] += delta_phi;
    if (s[3] > pi) s[3] -= 2 * pi;
    else if (s[3] < -pi) s[3] += 2 * pi;
    s[0] += s[5] * cos(s[4]) * cos(phi) * step_size;   /* commanded phi, updated gamma (:53-55) */
    s[1] += s[5] * cos(s[4]) * sin(phi) * step_size;
    s[2] += s[5] * sin(s[4]) * step_size;
}

static void orc_park(double *s)
{ /* particle_env.py:303-309 */
    s[0] = s[1] = s[2] = 1000.0;
    s[3] = s[4] = s[5] = 0.0;
}

/* particle_env.py:204-217 step (+ :219-238, :263-321, :336-346) for ONE env */
void orc_env3d_step(const marl_env3d_params *p, double *ps, uint8_t *pa, double *es, uint8_t *ea, const double *target,
                    const double *action, int32_t *time_step, int32_t *reward, uint8_t *done)
{
    const int N = p->N;
    *time_step += 1;
    for (int i = 0; i < N; ++i)
        if (pa[i]) orc_point_step(ps + 6 * i, action + 3 * i, p->p_vmax, p->ang_lmt, p->v_lmt, p->step_size);
    /* reward(True): only active evaders / active teammates are visible (get_team_state(rules=True)) */
    for (int i = 0; i < N; ++i) {
        int r = 0;
        if (pa[i]) {
            const double *s = ps + 6 * i;
            if (*ea && orc_norm3(s[0] - es[0], s[1] - es[1], s[2] - es[2]) <= p->kill_radius) r += 1;
            int inner = 0;
            for (int j = 0; j < N; ++j)
                if (pa[j]) {
                    const double *q = ps + 6 * j;
                    if (orc_norm3(s[0] - q[0], s[1] - q[1], s[2] - q[2]) <= p->kill_radius) inner += 1;
                }
            r -= (inner - 1);
        }
        reward[i] = r;
    }
    /* update_agent_active: all verdicts from the pre-update state, then applied */
    uint8_t dead[MARL_MAX_AGENTS];
    uint8_t e_dead = 0;
    for (int i = 0; i < N; ++i) {
        dead[i] = 0;
        if (!pa[i]) continue;
        const double *s = ps + 6 * i;
        int c = 0;
        for (int j = 0; j < N; ++j)
            if (pa[j]) {
                const double *q = ps + 6 * j;
                c += orc_norm3(s[0] - q[0], s[1] - q[1], s[2] - q[2]) <= p->kill_radius;
            }
        if (*ea) c += orc_norm3(s[0] - es[0], s[1] - es[1], s[2] - es[2]) <= p->kill_radius;
        dead[i] = (c - 1) != 0;
    }
    if (*ea) {
        int c = 1; /* itself */
        for (int j = 0; j < N; ++j)
            if (pa[j]) {
                const double *q = ps + 6 * j;
                c += orc_norm3(es[0] - q[0], es[1] - q[1], es[2] - q[2]) <= p->kill_radius;
            }
        e_dead = (c - 1) != 0;
    }
    for (int i = 0; i < N; ++i)
        if (dead[i]) { pa[i] = 0; orc_park(ps + 6 * i); }
    if (e_dead) { *ea = 0; orc_park(es); }
    /* get_done */
    int d = orc_norm3(es[0] - target[0], es[1] - target[1], es[2] - target[2]) <= p->kill_radius;
    int alive = 0;
    for (int i = 0; i < N; ++i) alive += pa[i];
    if (alive == 0) d = 1;
    if (!*ea) d = 1;
    *done = (d || *time_step >= p->max_step) ? 1 : 0;
}

/* particle_env.py:323-334 get_adj_mat: pursuer-pursuer (comm_range) and pursuer-evader (sen_range) */
void orc_env3d_adjacency(const marl_env3d_params *p, const double *ps, const uint8_t *pa, const double *es,
                         uint8_t *pp_adj, uint8_t *pe_adj)
{
    const int N = p->N;
    for (int i = 0; i < N; ++i) {
        const double *s = ps + 6 * i;
        for (int j = 0; j < N; ++j) {
            const double *q = ps + 6 * j;
            pp_adj[i * N + j] = pa[i] && orc_norm3(s[0] - q[0], s[1] - q[1], s[2] - q[2]) <= p->comm_range;
        }
        pe_adj[i] = pa[i] && orc_norm3(s[0] - es[0], s[1] - es[1], s[2] - es[2]) <= p->sen_range;
    }
}

/* one iteration of the batched loop the CUDA rollout runs: adjacency -> evader move -> step, all envs (OpenMP) */
void orc_env3d_iteration(const marl_env3d_params *p, int32_t B, double *ps, uint8_t *pa, double *es, uint8_t *ea,
                         const double *target, const double *action, const double *e_action, int32_t *time_step,
                         int32_t *reward, uint8_t *done, uint8_t *pp_adj, uint8_t *pe_adj)
{
    const int N = p->N;
#pragma omp parallel for schedule(static)
    for (int32_t b = 0; b < B; ++b) {
        orc_env3d_adjacency(p, ps + (size_t)b * N * 6, pa + (size_t)b * N, es + (size_t)b * 6,
                            pp_adj + (size_t)b * N * N, pe_adj + (size_t)b * N);
        if (ea[b]) orc_point_step(es + (size_t)b * 6, e_action + (size_t)b * 3, p->e_vmax, p->ang_lmt, p->v_lmt, p->step_size);
        orc_env3d_step(p, ps + (size_t)b * N * 6, pa + (size_t)b * N, es + (size_t)b * 6, ea + b, target + (size_t)b * 3,
                       action + (size_t)b * N * 3, time_step + b, reward + (size_t)b * N, done + b);
    }
}
