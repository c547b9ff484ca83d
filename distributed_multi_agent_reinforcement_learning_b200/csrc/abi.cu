// abi.cu — version, error string and validation of the scalar configuration.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace marl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MARL_ECUDA;
    }
    return MARL_OK;
}

// largest double s with sqrt(s) <= r  (strict=false)  or  sqrt(s) < r  (strict=true)
double sq_threshold(double r, bool strict)
{
    auto ok = [&](double s) { double q = sqrt(s); return strict ? (q < r) : (q <= r); };
    double s = r * r;
    while (!ok(s)) s = nextafter(s, -INFINITY);
    while (ok(nextafter(s, INFINITY))) s = nextafter(s, INFINITY);
    return s;
}

static int int_sq_floor(double r, int limit)
{   // largest integer d2 in [0,limit] with sqrt((double)d2) <= r, i.e. NOT (sqrt(d2) > r)
    int best = -1;
    for (int d2 = 0; d2 <= limit; ++d2)
        if (!(sqrt((double)d2) > r)) best = d2; else break;
    return best;
}

int make_env_dev(const marl_env_params *p, EnvDev *o)
{
    MARL_REQUIRE(p != nullptr, "params is NULL");
    MARL_REQUIRE(p->N >= 2 && p->N <= MARL_MAX_AGENTS, "N=%d unsupported (2..%d; the reference itself indexes column 1 of the NxN adjacency)", p->N, MARL_MAX_AGENTS);
    MARL_REQUIRE(p->W >= 2 && p->H >= 2 && (int64_t)p->W * p->H <= 16384, "map %dx%d unsupported (W*H <= 16384)", p->W, p->H);
    MARL_REQUIRE(p->O >= 1 && p->O <= 1024, "O=%d unsupported (1..1024)", p->O);
    MARL_REQUIRE(p->max_steps >= 1 && p->difficulty >= 1, "max_steps=%d difficulty=%d", p->max_steps, p->difficulty);
    MARL_REQUIRE(p->sensor_beams >= 1 && p->sensor_beams <= 360 && p->sensor_radius >= 1 && p->sensor_radius <= 64, "sensor %d beams radius %d", p->sensor_beams, p->sensor_radius);
    MARL_REQUIRE(p->d_tau > 0 && p->e_tau > 0 && p->d_step > 0 && p->e_step > 0, "tau/step must be positive");
    MARL_REQUIRE(p->d_collision_radius >= 0 && p->d_comm_range >= 0 && p->d_sen_range >= 0 && p->e_collision_radius >= 0 && p->resolution > 0, "negative radius");
    MARL_REQUIRE(p->e_extend_dis >= 0 && p->e_extend_dis <= 8 && p->e_sen_range >= 0 && p->e_sen_range <= 32, "evader extend_dis=%d sen_range=%d", p->e_extend_dis, p->e_sen_range);
    o->W = p->W; o->H = p->H; o->N = p->N; o->O = p->O;
    o->HW = (p->H + 31) / 32; o->OW = (p->O + 31) / 32; o->NW = (p->N + 31) / 32;
    o->max_steps = p->max_steps; o->difficulty = p->difficulty;
    o->sensor_beams = p->sensor_beams; o->sensor_radius = p->sensor_radius;
    o->e_extend_dis = p->e_extend_dis; o->e_sen_range = p->e_sen_range;
    o->d_step = p->d_step; o->d_tau = p->d_tau; o->d_vmax = p->d_vmax; o->d_radius = p->d_collision_radius;
    o->e_step = p->e_step; o->e_tau = p->e_tau; o->e_vmax = p->e_vmax; o->e_radius = p->e_collision_radius;
    o->resolution = p->resolution;
    o->thr2_collision = sq_threshold(p->d_collision_radius, false);
    o->thr2_comm = sq_threshold(p->d_comm_range, false);
    o->thr2_e_capture = sq_threshold(p->e_collision_radius, false);
    o->thr2_resolution_lt = sq_threshold(p->resolution, true);
    const int lim = p->W * p->W + p->H * p->H;
    o->sen_range2_floor = int_sq_floor(p->d_sen_range, lim);
    o->e_view2_floor = int_sq_floor((double)p->e_sen_range, lim);
    o->x_hi = (double)(p->W - 1);
    o->y_hi = (double)(p->H - 1);
    return MARL_OK;
}

}  // namespace marl

extern "C" int marl_version(void) { return MARL_ABI_VERSION; }
extern "C" const char *marl_last_error_string(void) { return marl::g_err; }
