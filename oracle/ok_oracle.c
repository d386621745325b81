/*
 * ok_oracle.c -- plain-C CPU restatement of OpenKitchen's per-tick hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ok_oracle.h).  Build: gcc -O2 -ffp-contract=off -fopenmp
 * (no -march, no fast-math: same arithmetic model as Environment/CMakeLists.txt:67, where the
 * x86-64 baseline ISA has no FMA so nothing can be contracted).
 *
 * Arithmetic model ("canonical"):
 *   - every float expression is evaluated in binary32, left to right, one rounding per op,
 *     exactly as the reference's host code does (Agent.cpp, CollisionChecker.cu host part);
 *   - cos/sin of a float are glibc's sincosf (the symbol Agent.cpp / CollisionChecker.cu's
 *     host loops resolve to), restated below from its published algorithm so that it does
 *     not depend on the box's libm: oko_sincosf() is bit-identical to glibc 2.39 sincosf
 *     (its FMA ifunc variant) for all 2^32 inputs (oracle/sincosf_exhaustive.c);
 *   - the device part of the reference (CollisionChecker.cu:8-71, compiled by nvcc with
 *     libdevice cosf/sinf and FMA contraction) is restated with the SAME sincosf and
 *     un-contracted binary32 ops.  This is the one place the restatement is a choice rather
 *     than a transcription: libdevice is not available on a CPU.
 *
 * Parity pin: tests/test_oracle_vs_ref.py compares this file with oracle/_ref/libokref.so,
 * which links the unmodified reference Agent.cpp + RaceTrack.cpp.
 */
#define _GNU_SOURCE
#include "ok_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Constants (Typedefs.h:7-10, Agent.h:10-12)                                                  */
/* ------------------------------------------------------------------------------------------ */
static const float kScreenWidth  = 1600.0f; /* Typedefs.h:7 */
static const float kScreenHeight = 1400.0f; /* Typedefs.h:8 */
#define K_DEG2RAD ((float)(M_PI / 180.0))   /* Typedefs.h:10: constexpr float kDeg2Rad{M_PI / 180.0F} */

/* ------------------------------------------------------------------------------------------ */
/* sincosf: third-party arithmetic not under /root/reference.                                  */
/* glibc 2.39 sysdeps/ieee754/flt-32/s_sincosf.{c,h} (ARM optimized-routines, S. Nagy):        */
/* double-precision range reduction by pi/2 and degree-4/3 minimax polynomials in double.      */
/* Call sites in the reference: Agent.cpp:94-96,115-117; CollisionChecker.cu:122-124,157-158.  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double sign[4];
    double hpi_inv, hpi;
    double c0, c1, c2, c3, c4;
    double s1, s2, s3;
} oko_sincos_tab;

static const oko_sincos_tab kSincosTab[2] = {
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     0x1p0,
     -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5,
     -0x1.6c087e89a359dp-10,
     0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     -0x1p0,
     0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5,
     0x1.6c087e89a359dp-10,
     -0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7,
     -0x1.994eb3774cf24p-13}};

/* 4/pi as a 192-bit fixed-point table, for |x| >= 120 */
static const uint32_t kInvPio4[24] = {0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44,
                                      0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
                                      0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62,
                                      0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

static inline uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline uint32_t abstop12(float x)
{
    return (f2u(x) >> 20) & 0x7ff;
}

/* One definition of "a*b + c" for the whole evaluation.  x86-64 glibc dispatches sincosf through
 * an ifunc: __sincosf_fma on CPUs with FMA (every "a + b*c" above contracted, as built with
 * -mfma -mavx2) and __sincosf_sse2 otherwise.  The FMA variant is the canonical definition here:
 * it is bit-identical to glibc 2.39 sincosf on this image's hosts for ALL 2^32 inputs
 * (exhaustive run); the un-fused variant differs from it on 34 of the 2^32 inputs.
 * Define OKO_SINCOS_NOFMA to get the SSE2 variant. */
#ifndef OKO_SINCOS_NOFMA
#define OKO_MADD(a, b, c) __builtin_fma((a), (b), (c))
#else
#define OKO_MADD(a, b, c) ((a) * (b) + (c))
#endif

static inline void sincos_poly(double x, double x2, const oko_sincos_tab *p, int n, float *sinp, float *cosp)
{
    double x3, x4, x5, x6, s, c, c1, c2, s1;
    x4 = x2 * x2;
    x3 = x2 * x;
    c2 = OKO_MADD(x2, p->c4, p->c3);
    s1 = OKO_MADD(x2, p->s3, p->s2);
    if (n & 1) { /* odd quadrant: swap the outputs */
        float *tmp = cosp;
        cosp       = sinp;
        sinp       = tmp;
    }
    c1 = OKO_MADD(x2, p->c1, p->c0);
    x5 = x3 * x2;
    x6 = x4 * x2;
    s  = OKO_MADD(x3, p->s1, x);
    c  = OKO_MADD(x4, p->c2, c1);
    *sinp = (float)OKO_MADD(x5, s1, s);
    *cosp = (float)OKO_MADD(x6, c2, c);
}

void oko_sincosf(float y, float *sinp, float *cosp)
{
    double                x = y;
    double                s;
    int                   n;
    const oko_sincos_tab *p = &kSincosTab[0];

    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) { /* |y| < pi/4 */
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) {
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
        sincos_poly(x, x2, p, 0, sinp, cosp);
    } else if (abstop12(y) < abstop12(120.0f)) {
        double r = x * p->hpi_inv;
        n        = ((int32_t)r + 0x800000) >> 24;
        x        = OKO_MADD(-(double)n, p->hpi, x);
        s        = p->sign[n & 3];
        if (n & 2)
            p = &kSincosTab[1];
        sincos_poly(x * s, x * x, p, n, sinp, cosp);
    } else if (abstop12(y) < abstop12(INFINITY)) {
        uint32_t        xi    = f2u(y);
        int             sign  = (int)(xi >> 31);
        const uint32_t *arr   = &kInvPio4[(xi >> 26) & 15];
        int             shift = (xi >> 23) & 7;
        uint64_t        nn, res0, res1, res2;
        xi = (xi & 0xffffff) | 0x800000;
        xi <<= shift;
        res0 = xi * arr[0];
        res1 = (uint64_t)xi * arr[4];
        res2 = (uint64_t)xi * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        x = (double)(int64_t)res0;
        n = (int)nn;
        x = x * 0x1.921FB54442D18p-62;
        s = p->sign[(n + sign) & 3];
        if ((n + sign) & 2)
            p = &kSincosTab[1];
        sincos_poly(x * s, x * x, p, n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y; /* inf / nan -> nan */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) -- synthetic action stream of SURVEY 8(d)              */
/* ------------------------------------------------------------------------------------------ */
void oko_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

/* ------------------------------------------------------------------------------------------ */
/* Track: RaceTrack.cpp                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int    n;      /* centre-line points */
    int    nseg;   /* 4*(n-1)+4 */
    float *x, *y, *wr, *wl, *heading;
    float *li, *lo, *ri, *ro; /* interleaved xy */
    float *seg;               /* x1,y1,x2,y2 */
} OkoTrack;

struct OkoEnv {
    OkoConfig cfg;
    int       ntracks, cap_tracks;
    OkoTrack *tracks;
    int64_t   n;
    int       rays;
    float    *ray_deg;
    void     *buf[OKO_BUF_COUNT];
};

static float fminf_std(float a, float b) { return (b < a) ? b : a; } /* std::min */
static float fmaxf_std(float a, float b) { return (a < b) ? b : a; } /* std::max */

/* RaceTrack::gradient, RaceTrack.cpp:87-114 */
static void gradient(const float *in, int n, float *out)
{
    for (int i = 0; i < n; ++i) {
        if (i == 0)
            out[i] = in[i + 1] - in[i];
        else if (i == n - 1)
            out[i] = in[i] - in[i - 1];
        else
            out[i] = (in[i + 1] - in[i - 1]) / 2.0f;
    }
}

/* RaceTrack::calculateTrackExtents, RaceTrack.cpp:166-196 */
static void extents(const float *x, const float *y, int n, float e[4])
{
    e[0] = e[1] = 3.402823466e+38f;
    e[2] = e[3] = -3.402823466e+38f;
    for (int i = 0; i < n; ++i) {
        if (x[i] < e[0]) e[0] = x[i];
        if (x[i] > e[2]) e[2] = x[i];
    }
    for (int i = 0; i < n; ++i) {
        if (y[i] < e[1]) e[1] = y[i];
        if (y[i] > e[3]) e[3] = y[i];
    }
}

static void add_polyline(const float *p, int n, float *seg, int *k)
{ /* TrackSegments::addPolylineSegments, TrackSegments.cu:53-67 */
    if (n < 2)
        return;
    for (int i = 0; i < n - 1; ++i) {
        float *s = seg + 4 * (*k)++;
        s[0] = p[2 * i], s[1] = p[2 * i + 1], s[2] = p[2 * i + 2], s[3] = p[2 * i + 3];
    }
}

static void add_closure(const float *p, int n, float *seg, int *k)
{ /* back -> front, TrackSegments.cu:16-38 */
    float *s = seg + 4 * (*k)++;
    s[0] = p[2 * (n - 1)], s[1] = p[2 * (n - 1) + 1], s[2] = p[0], s[3] = p[1];
}

int oko_add_track(OkoEnv *env, const float *x_m, const float *y_m, const float *w_right, const float *w_left, int n)
{
    if (n < 2)
        return -1;
    if (env->ntracks == env->cap_tracks) {
        env->cap_tracks = env->cap_tracks ? 2 * env->cap_tracks : 32;
        env->tracks     = (OkoTrack *)realloc(env->tracks, sizeof(OkoTrack) * (size_t)env->cap_tracks);
    }
    OkoTrack *t = &env->tracks[env->ntracks];
    t->n        = n;
    t->x = (float *)malloc(sizeof(float) * n), t->y = (float *)malloc(sizeof(float) * n);
    t->wr = (float *)malloc(sizeof(float) * n), t->wl = (float *)malloc(sizeof(float) * n);
    t->heading = (float *)malloc(sizeof(float) * n);
    t->li = (float *)malloc(sizeof(float) * 2 * n), t->lo = (float *)malloc(sizeof(float) * 2 * n);
    t->ri = (float *)malloc(sizeof(float) * 2 * n), t->ro = (float *)malloc(sizeof(float) * 2 * n);

    /* width clamp, RaceTrack.cpp:138-160: min(max(4, w) * 3, 17) */
    for (int i = 0; i < n; ++i) {
        t->x[i]  = x_m[i];
        t->y[i]  = y_m[i];
        t->wr[i] = fminf_std(fmaxf_std(4.0f, w_right[i]) * 3.0f, 17.0f);
        t->wl[i] = fminf_std(fmaxf_std(4.0f, w_left[i]) * 3.0f, 17.0f);
    }
    /* centerTrackPointsToWindow, RaceTrack.cpp:198-229 */
    float e[4];
    extents(t->x, t->y, n, e);
    const float track_width  = e[2] - e[0];
    const float track_height = e[3] - e[1];
    float       sx           = kScreenWidth / track_width;
    float       sy           = kScreenHeight / track_height;
    float       scale        = fminf_std(sx, sy);
    scale *= (float)0.9; /* constexpr float kScreenFitScale{0.9} */
    for (int i = 0; i < n; ++i) {
        t->x[i] *= scale;
        t->y[i] *= scale;
        t->wl[i] *= scale;
        t->wr[i] *= scale;
    }
    extents(t->x, t->y, n, e);
    const float cx = (kScreenWidth / 2.0f) - ((e[2] + e[0]) / 2.0f);
    const float cy = (kScreenHeight / 2.0f) - ((e[3] + e[1]) / 2.0f);
    for (int i = 0; i < n; ++i) {
        t->x[i] += cx;
        t->y[i] += cy;
    }
    /* calculateTrackLanes, RaceTrack.cpp:257-307 */
    float *dx = (float *)malloc(sizeof(float) * n), *dy = (float *)malloc(sizeof(float) * n);
    gradient(t->x, n, dx);
    gradient(t->y, n, dy);
    for (int i = 0; i < n; ++i) {
        float mag = sqrtf(dx[i] * dx[i] + dy[i] * dy[i]);
        dx[i] /= mag;
        dy[i] /= mag;
        /* std::atan2(float,float) * 180.0F / M_PI : float product, double division, narrowed */
        t->heading[i] = (float)((double)(atan2f(dy[i], dx[i]) * 180.0f) / M_PI);
    }
    const float kThick = 3.0f;
    for (int i = 0; i < n; ++i) {
        t->ri[2 * i]     = t->x[i] + t->wr[i] * dy[i];
        t->ri[2 * i + 1] = t->y[i] - t->wr[i] * dx[i];
        t->li[2 * i]     = t->x[i] - t->wl[i] * dy[i];
        t->li[2 * i + 1] = t->y[i] + t->wl[i] * dx[i];
        t->ro[2 * i]     = t->x[i] + (t->wr[i] + kThick) * dy[i];
        t->ro[2 * i + 1] = t->y[i] - (t->wr[i] + kThick) * dx[i];
        t->lo[2 * i]     = t->x[i] - (t->wl[i] + kThick) * dy[i];
        t->lo[2 * i + 1] = t->y[i] + (t->wl[i] + kThick) * dx[i];
    }
    free(dx), free(dy);
    /* TrackSegments ctor, TrackSegments.cu:6-42: LI, LO, RI, RO polylines, closures LI, RI, LO, RO */
    t->nseg = 4 * (n - 1) + 4;
    t->seg  = (float *)malloc(sizeof(float) * 4 * (size_t)t->nseg);
    int k   = 0;
    add_polyline(t->li, n, t->seg, &k);
    add_polyline(t->lo, n, t->seg, &k);
    add_polyline(t->ri, n, t->seg, &k);
    add_polyline(t->ro, n, t->seg, &k);
    add_closure(t->li, n, t->seg, &k);
    add_closure(t->ri, n, t->seg, &k);
    add_closure(t->lo, n, t->seg, &k);
    add_closure(t->ro, n, t->seg, &k);
    return env->ntracks++;
}

/* getTrackDataFromCsv, RaceTrack.cpp:127-164: skip the header, 4 comma-separated std::stof */
int oko_load_track_csv(OkoEnv *env, const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f)
        return -1;
    size_t cap = 1024, n = 0;
    float *col[4];
    for (int c = 0; c < 4; ++c)
        col[c] = (float *)malloc(sizeof(float) * cap);
    char line[512];
    if (!fgets(line, sizeof line, f)) {
        fclose(f);
        return -1;
    }
    while (fgets(line, sizeof line, f)) {
        char *p = line;
        if (n == cap) {
            cap *= 2;
            for (int c = 0; c < 4; ++c)
                col[c] = (float *)realloc(col[c], sizeof(float) * cap);
        }
        int ok = 1;
        for (int c = 0; c < 4; ++c) {
            char *end;
            float v = strtof(p, &end); /* std::stof == strtof */
            if (end == p) {
                ok = 0;
                break;
            }
            col[c][n] = v;
            p         = strchr(p, ',');
            if (!p && c < 3) {
                ok = 0;
                break;
            }
            if (p)
                ++p;
        }
        if (!ok)
            break;
        ++n;
    }
    fclose(f);
    int id = oko_add_track(env, col[0], col[1], col[2], col[3], (int)n);
    for (int c = 0; c < 4; ++c)
        free(col[c]);
    return id;
}

int oko_num_tracks(const OkoEnv *env) { return env->ntracks; }
int oko_track_points(const OkoEnv *env, int t) { return env->tracks[t].n; }
int oko_track_segments(const OkoEnv *env, int t) { return env->tracks[t].nseg; }
const float *oko_track_array(const OkoEnv *env, int t, int which)
{
    const OkoTrack *k = &env->tracks[t];
    switch (which) {
    case 0: return k->x;
    case 1: return k->y;
    case 2: return k->wr;
    case 3: return k->wl;
    case 4: return k->heading;
    case 5: return k->li;
    case 6: return k->lo;
    case 7: return k->ri;
    case 8: return k->ro;
    case 9: return k->seg;
    default: return NULL;
    }
}

/* Vec2d::distanceSquared, Typedefs.h:40-43 */
static inline float dist2(float ax, float ay, float bx, float by)
{
    return (ax - bx) * (ax - bx) + (ay - by) * (ay - by);
}

static int32_t nearest_index(const OkoTrack *t, float qx, float qy, float *min_d2_out)
{ /* RaceTrack::findNearestTrackIndexBruteForce, RaceTrack.cpp:16-31 (strict <, lowest index wins) */
    float   min_d = 3.402823466e+38f;
    int32_t idx   = 0;
    for (int i = 0; i < t->n; ++i) {
        float d = dist2(qx, qy, t->x[i], t->y[i]);
        if (d < min_d) {
            min_d = d;
            idx   = i;
        }
    }
    if (min_d2_out)
        *min_d2_out = min_d;
    return idx;
}

int32_t oko_nearest_index(const OkoEnv *env, int track, float x, float y)
{
    return nearest_index(&env->tracks[track], x, y, NULL);
}

float oko_dist_lane_center(const OkoEnv *env, int track, float x, float y)
{ /* RaceTrack.cpp:53-72 */
    const OkoTrack *t = &env->tracks[track];
    float           d2;
    int32_t         i = nearest_index(t, x, y, &d2);
    const float     w = t->wl[i] + t->wr[i];
    return sqrtf(d2) / w;
}

float oko_dist_boundary(const OkoEnv *env, int track, float x, float y)
{ /* RaceTrack.cpp:33-51 */
    const OkoTrack *t     = &env->tracks[track];
    float           min_d = 3.402823466e+38f;
    for (int i = 0; i < t->n; ++i) {
        float d = dist2(x, y, t->li[2 * i], t->li[2 * i + 1]);
        if (d < min_d) min_d = d;
        d = dist2(x, y, t->ri[2 * i], t->ri[2 * i + 1]);
        if (d < min_d) min_d = d;
    }
    return sqrtf(min_d);
}

float oko_normalize_angle_deg(float a)
{ /* Utils.h:3-14 */
    while (a < 360.0f)
        a += 360.0f;
    while (a >= 360.0f)
        a -= 360.0f;
    return a;
}

/* ------------------------------------------------------------------------------------------ */
/* Environment                                                                                 */
/* ------------------------------------------------------------------------------------------ */
void oko_config_default(OkoConfig *c)
{
    memset(c, 0, sizeof *c);
    c->movement_mode        = OKO_MOVE_VELOCITY; /* Agent.h:80 */
    c->reward_mode          = OKO_REWARD_NONE;
    c->auto_reset           = 0;
    c->auto_reset_stride    = 97;
    c->sensor_range         = 200.0f;
    c->speed_limit          = 100.0f;
    c->dt                   = (float)0.016;
    c->collision_dist2      = 2.0f;
    c->sensor_offset        = 0.0f;
    c->standstill_period    = 200u;
    c->standstill_threshold = 20.0f;
}

OkoEnv *oko_create(const OkoConfig *cfg)
{
    OkoEnv *e = (OkoEnv *)calloc(1, sizeof(OkoEnv));
    if (cfg)
        e->cfg = *cfg;
    else
        oko_config_default(&e->cfg);
    return e;
}

/* change the constants of a live env (the product's ok_update_config) */
void oko_update_config(OkoEnv *e, const OkoConfig *cfg)
{
    e->cfg = *cfg;
}

static void free_agents(OkoEnv *e)
{
    for (int i = 0; i < OKO_BUF_COUNT; ++i) {
        free(e->buf[i]);
        e->buf[i] = NULL;
    }
    free(e->ray_deg);
    e->ray_deg = NULL;
    e->n       = 0;
}

void oko_destroy(OkoEnv *e)
{
    if (!e)
        return;
    free_agents(e);
    for (int t = 0; t < e->ntracks; ++t) {
        OkoTrack *k = &e->tracks[t];
        free(k->x), free(k->y), free(k->wr), free(k->wl), free(k->heading);
        free(k->li), free(k->lo), free(k->ri), free(k->ro), free(k->seg);
    }
    free(e->tracks);
    free(e);
}

void oko_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int oko_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

static size_t buf_bytes(const OkoEnv *e, int which)
{
    const size_t n = (size_t)e->n, r = (size_t)e->rays;
    switch (which) {
    case OKO_BUF_CRASHED:
    case OKO_BUF_TIMED_OUT:
    case OKO_BUF_DONE: return n;
    case OKO_BUF_HIT_ABS:
    case OKO_BUF_HIT_REL: return n * r * 8;
    case OKO_BUF_OBS:
    case OKO_BUF_HIT_SEG:
    case OKO_BUF_HIT_T: return n * r * 4;
    default: return n * 4;
    }
}

int oko_alloc_agents(OkoEnv *e, int64_t n, int rays, const float *ray_deg, const int32_t *track_id)
{
    if (n <= 0 || rays <= 0 || e->ntracks == 0)
        return -1;
    free_agents(e);
    e->n       = n;
    e->rays    = rays;
    e->ray_deg = (float *)malloc(sizeof(float) * (size_t)rays);
    memcpy(e->ray_deg, ray_deg, sizeof(float) * (size_t)rays);
    for (int i = 0; i < OKO_BUF_COUNT; ++i)
        e->buf[i] = calloc(1, buf_bytes(e, i) + 16);
    memset(e->buf[OKO_BUF_HIT_SEG], 0xff, buf_bytes(e, OKO_BUF_HIT_SEG)); /* -1: no hit recorded yet */
    int32_t *tid = (int32_t *)e->buf[OKO_BUF_TRACK_ID];
    for (int64_t i = 0; i < n; ++i) {
        tid[i] = track_id ? track_id[i] : 0;
        if (tid[i] < 0 || tid[i] >= e->ntracks)
            return -1;
    }
    /* every agent starts like `new Agent` + resetAgent(pick_random_point=false): RaceTrack.h:18 */
    for (int64_t i = 0; i < n; ++i) {
        int32_t pt = 3 < e->tracks[tid[i]].n ? 3 : 0;
        oko_reset_agents(e, &i, &pt, NULL, NULL, 1);
    }
    return 0;
}

int64_t oko_num_agents(const OkoEnv *e) { return e->n; }
void   *oko_buffer(OkoEnv *e, int which) { return (which >= 0 && which < OKO_BUF_COUNT) ? e->buf[which] : NULL; }

#define F32(e, id) ((float *)(e)->buf[id])
#define I32(e, id) ((int32_t *)(e)->buf[id])
#define U32(e, id) ((uint32_t *)(e)->buf[id])
#define U8(e, id) ((uint8_t *)(e)->buf[id])

/* Environment::resetAgent (Environment.cpp:79-122) + Agent::reset (Agent.cpp:123-135).
 * The standstill counters are NOT touched (Environment.cpp never resets displacement_stats_). */
static void reset_one(OkoEnv *e, int64_t a, int32_t pt, int use_lane, float alpha, float heading_off)
{
    const OkoTrack *t = &e->tracks[I32(e, OKO_BUF_TRACK_ID)[a]];
    float           sx, sy;
    if (use_lane) { /* Environment.cpp:107-114 */
        const float lx = t->li[2 * pt], ly = t->li[2 * pt + 1];
        const float rx = t->ri[2 * pt], ry = t->ri[2 * pt + 1];
        sx = lx * alpha + rx * (1.0f - alpha);
        sy = ly * alpha + ry * (1.0f - alpha);
    } else {
        sx = t->x[pt];
        sy = t->y[pt];
    }
    F32(e, OKO_BUF_POS_X)[a]        = sx;
    F32(e, OKO_BUF_POS_Y)[a]        = sy;
    F32(e, OKO_BUF_ROT)[a]          = t->heading[pt] + heading_off; /* Environment.cpp:121 */
    F32(e, OKO_BUF_ACCEL)[a]        = 0.0f;
    F32(e, OKO_BUF_SPEED)[a]        = 0.0f;
    U8(e, OKO_BUF_CRASHED)[a]       = 0;
    U8(e, OKO_BUF_TIMED_OUT)[a]     = 0;
    U8(e, OKO_BUF_DONE)[a]          = 0;
    F32(e, OKO_BUF_ACT_THROTTLE)[a] = 0.0f;
    F32(e, OKO_BUF_ACT_STEER)[a]    = 0.0f;
    /* episode bookkeeping that the reference keeps app-side (main_eigen.cpp:121-130:
     * prev_track_idx_ = nearest index of the post-reset pose; fitness restarts) */
    I32(e, OKO_BUF_RESET_PT)[a]    = pt;
    F32(e, OKO_BUF_START_X)[a]     = sx;
    F32(e, OKO_BUF_START_Y)[a]     = sy;
    I32(e, OKO_BUF_PREV_IDX)[a]    = nearest_index(t, sx, sy, NULL);
    I32(e, OKO_BUF_NEAREST_IDX)[a] = I32(e, OKO_BUF_PREV_IDX)[a];
    F32(e, OKO_BUF_FITNESS)[a]     = 0.0f;
    F32(e, OKO_BUF_REWARD)[a]      = 0.0f;
}

void oko_reset_agents(OkoEnv *e, const int64_t *agent_idx, const int32_t *pt_idx, const float *lane_alpha,
                      const float *heading_off, int64_t n)
{
    for (int64_t k = 0; k < n; ++k) {
        const int64_t a = agent_idx ? agent_idx[k] : k;
        reset_one(e, a, pt_idx[k], lane_alpha != NULL, lane_alpha ? lane_alpha[k] : 0.0f,
                  heading_off ? heading_off[k] : 0.0f);
    }
}

/* Agent::moveViaVelocity / moveViaAcceleration, Agent.cpp:82-98,108-119 */
static void move_one(OkoEnv *e, int64_t a)
{
    const float kDt   = e->cfg.dt;
    float      *rot   = &F32(e, OKO_BUF_ROT)[a];
    float      *speed = &F32(e, OKO_BUF_SPEED)[a];
    float      *acc   = &F32(e, OKO_BUF_ACCEL)[a];
    const float thr   = F32(e, OKO_BUF_ACT_THROTTLE)[a];
    const float steer = F32(e, OKO_BUF_ACT_STEER)[a];
    *rot += steer;
    if (e->cfg.movement_mode == OKO_MOVE_ACCELERATION) {
        *acc += thr;
        *speed += (*acc * kDt);
        *speed = (*speed < 0.0f) ? 0.0f : *speed;
        *speed = (*speed > e->cfg.speed_limit) ? e->cfg.speed_limit : *speed;
    } else {
        *speed = thr;
    }
    float s, c;
    oko_sincosf(K_DEG2RAD * *rot, &s, &c);
    const float delta_x = c * *speed * kDt;
    F32(e, OKO_BUF_POS_X)[a] += delta_x;
    const float delta_y = s * *speed * kDt;
    F32(e, OKO_BUF_POS_Y)[a] += delta_y;
}

/* checkAndUpdateStandstill, Environment.cpp:16-39; returns displacement_timed_out */
static int standstill_one(OkoEnv *e, int64_t a)
{
    uint32_t   *ctr = &U32(e, OKO_BUF_SS_CTR)[a];
    const float px = F32(e, OKO_BUF_POS_X)[a], py = F32(e, OKO_BUF_POS_Y)[a];
    if (*ctr == 0) {
        F32(e, OKO_BUF_SS_X)[a] = px;
        F32(e, OKO_BUF_SS_Y)[a] = py;
        (*ctr)++;
        return 0;
    }
    if (*ctr >= e->cfg.standstill_period) {
        const float moved = dist2(px, py, F32(e, OKO_BUF_SS_X)[a], F32(e, OKO_BUF_SS_Y)[a]);
        const int   out   = moved < e->cfg.standstill_threshold * e->cfg.standstill_threshold;
        *ctr              = 0;
        return out;
    }
    (*ctr)++;
    return 0;
}

/* raySegmentIntersect + castRaysToSegmentsKernel, CollisionChecker.cu:8-71, and the host
 * pack / unpack around it, CollisionChecker.cu:115-128,144-172.  Extra outputs that the
 * reference does not keep: hit_seg (last index that lowered min_t, -1 if none) and min_t. */
static void cast_one(OkoEnv *e, int64_t a)
{
    const OkoTrack *t    = &e->tracks[I32(e, OKO_BUF_TRACK_ID)[a]];
    const int       R    = e->rays;
    const float     rot  = F32(e, OKO_BUF_ROT)[a];
    const float     px   = F32(e, OKO_BUF_POS_X)[a];
    const float     py   = F32(e, OKO_BUF_POS_Y)[a];
    const int       live = !U8(e, OKO_BUF_CRASHED)[a];
    float          *habs = F32(e, OKO_BUF_HIT_ABS) + (size_t)a * R * 2;
    float          *hrel = F32(e, OKO_BUF_HIT_REL) + (size_t)a * R * 2;
    float          *obs  = F32(e, OKO_BUF_OBS) + (size_t)a * R;
    float          *ht   = F32(e, OKO_BUF_HIT_T) + (size_t)a * R;
    int32_t        *hseg = I32(e, OKO_BUF_HIT_SEG) + (size_t)a * R;
    float           rs, rc;
    oko_sincosf(K_DEG2RAD * rot, &rs, &rc);
    /* pack, CollisionChecker.cu:121-124 */
    const float ox = px + e->cfg.sensor_offset * rc;
    const float oy = py + e->cfg.sensor_offset * rs;
    float       min_dist2 = e->cfg.sensor_range * e->cfg.sensor_range;
    for (int r = 0; r < R; ++r) {
        if (live) { /* kernel early-return for inactive rays, CollisionChecker.cu:44 */
            const float angle = K_DEG2RAD * (rot + e->ray_deg[r]);
            float       dx, dy;
            {
                float s, c;
                oko_sincosf(angle, &s, &c);
                dx = c, dy = s;
            }
            float   min_t = e->cfg.sensor_range;
            int32_t idx   = -1;
            for (int i = 0; i < t->nseg; ++i) {
                const float *sg     = t->seg + 4 * (size_t)i;
                const float  seg_dx = sg[2] - sg[0];
                const float  seg_dy = sg[3] - sg[1];
                const float  denom  = dx * seg_dy - dy * seg_dx;
                if (fabsf(denom) < 1e-8f)
                    continue;
                const float tt = ((sg[0] - ox) * seg_dy - (sg[1] - oy) * seg_dx) / denom;
                const float ss = ((sg[0] - ox) * dy - (sg[1] - oy) * dx) / denom;
                if ((tt >= 0.0f) && (tt <= min_t) && (ss >= 0.0f) && (ss <= 1.0f)) {
                    min_t = tt;
                    idx   = i;
                }
            }
            habs[2 * r]     = ox + min_t * dx;
            habs[2 * r + 1] = oy + min_t * dy;
            ht[r]           = min_t;
            hseg[r]         = idx;
        }
        /* unpack, CollisionChecker.cu:152-165 (runs for crashed agents too, on stale hits) */
        const float xt = habs[2 * r] - ox;
        const float yt = habs[2 * r + 1] - oy;
        const float hx = xt * rc - yt * rs;
        const float hy = xt * rs + yt * rc;
        hrel[2 * r]     = hx;
        hrel[2 * r + 1] = hy;
        const float sq  = hx * hx + hy * hy;
        obs[r]          = sqrtf(sq) / e->cfg.sensor_range; /* PPOAgent.hpp:68-76 */
        if (sq < min_dist2)
            min_dist2 = sq;
    }
    F32(e, OKO_BUF_MIN_DIST2)[a] = min_dist2;
    if (min_dist2 < e->cfg.collision_dist2) /* CollisionChecker.cu:167-171 */
        U8(e, OKO_BUF_CRASHED)[a] = 1;
    U8(e, OKO_BUF_DONE)[a] = U8(e, OKO_BUF_CRASHED)[a]; /* Agent::isDone, Agent.cpp:138-144 */
}

/* App-side progress / reward / done after env.step(); see the mode enum for the sources. */
static void reward_one(OkoEnv *e, int64_t a)
{
    const OkoTrack *t       = &e->tracks[I32(e, OKO_BUF_TRACK_ID)[a]];
    const int       crashed = U8(e, OKO_BUF_CRASHED)[a];
    const int       tout    = U8(e, OKO_BUF_TIMED_OUT)[a];
    const float     px = F32(e, OKO_BUF_POS_X)[a], py = F32(e, OKO_BUF_POS_Y)[a];
    float          *reward  = &F32(e, OKO_BUF_REWARD)[a];
    float          *fitness = &F32(e, OKO_BUF_FITNESS)[a];
    int32_t        *prev    = &I32(e, OKO_BUF_PREV_IDX)[a];
    const int       mode    = e->cfg.reward_mode;
    U8(e, OKO_BUF_DONE)[a]  = (uint8_t)crashed; /* Agent::isDone, Agent.cpp:138-144; completed_ is never set */
    float   d2   = 0.0f;
    int32_t near = I32(e, OKO_BUF_NEAREST_IDX)[a];
    if (mode == OKO_REWARD_Q_PROGRESS || mode == OKO_REWARD_CMAES_PROGRESS || mode == OKO_REWARD_TRACK_INDEX ||
        mode == OKO_REWARD_LANE_CENTER) {
        near                           = nearest_index(t, px, py, &d2);
        I32(e, OKO_BUF_NEAREST_IDX)[a] = near;
    }
    switch (mode) {
    case OKO_REWARD_Q_PROGRESS: {
        if (crashed) {
            *reward = -200.0f;
            break;
        }
        int64_t progression = (int64_t)near - (int64_t)*prev;
        *prev               = near;
        int64_t len         = t->n;
        int64_t ab          = progression < 0 ? -progression : progression;
        *reward             = (float)(ab > (len / 2) ? len - ab : ab);
        break;
    }
    case OKO_REWARD_CMAES_PROGRESS: {
        if (!crashed) {
            int32_t progress = near - *prev;
            *prev            = near;
            float p          = (float)(progress < 0 ? -progress : progress);
            *fitness += p;
            *reward = p;
        } else {
            if (tout)
                *fitness = 0.0f;
            *reward = 0.0f;
        }
        break;
    }
    case OKO_REWARD_CONSTANT: *reward = 1.0f; break;
    case OKO_REWARD_DISPLACEMENT: {
        const float ddx = px - F32(e, OKO_BUF_START_X)[a];
        const float ddy = py - F32(e, OKO_BUF_START_Y)[a];
        *reward         = crashed ? -5.0f : sqrtf(ddx * ddx + ddy * ddy);
        break;
    }
    case OKO_REWARD_MIN_RAY: {
        if (crashed) {
            *reward = -200.0f;
            break;
        }
        const float *hrel = F32(e, OKO_BUF_HIT_REL) + (size_t)a * e->rays * 2;
        float        m    = e->cfg.sensor_range;
        for (int r = 0; r < e->rays; ++r) {
            float nrm = sqrtf(hrel[2 * r] * hrel[2 * r] + hrel[2 * r + 1] * hrel[2 * r + 1]);
            if (m > nrm)
                m = nrm;
        }
        *reward = m;
        break;
    }
    case OKO_REWARD_TRACK_INDEX: *reward = (float)near; break;
    case OKO_REWARD_LANE_CENTER: {
        if (!crashed) {
            const float w       = t->wl[near] + t->wr[near];
            const float closest = sqrtf(d2) / w;
            const float r       = 1.0f - closest;
            *fitness += r;
            *reward = r;
        } else {
            if (tout)
                *fitness = 0.0f;
            *reward = 0.0f;
        }
        break;
    }
    default: *reward = 0.0f; break;
    }
}

void oko_cast_rays(OkoEnv *e)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t a = 0; a < e->n; ++a)
        cast_one(e, a);
}

void oko_step(OkoEnv *e, const float *act_throttle, const float *act_steer)
{
    if (act_throttle)
        memcpy(e->buf[OKO_BUF_ACT_THROTTLE], act_throttle, sizeof(float) * (size_t)e->n);
    if (act_steer)
        memcpy(e->buf[OKO_BUF_ACT_STEER], act_steer, sizeof(float) * (size_t)e->n);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t a = 0; a < e->n; ++a) {
        /* optional app-side reset of crashed agents before the tick (GuidedCostLearning/test.cpp:102-111):
         * Agent::reset zeroes current_action_, so the first tick after a reset runs with a zero action */
        if (e->cfg.auto_reset && U8(e, OKO_BUF_CRASHED)[a]) {
            const OkoTrack *t  = &e->tracks[I32(e, OKO_BUF_TRACK_ID)[a]];
            int32_t         pt = (int32_t)(((int64_t)I32(e, OKO_BUF_RESET_PT)[a] + e->cfg.auto_reset_stride) % t->n);
            reset_one(e, a, pt, 0, 0.0f, 0.0f);
        }
        /* 1) kinematics, Environment.cpp:128-143 */
        if (!U8(e, OKO_BUF_CRASHED)[a]) {
            move_one(e, a);
            if (standstill_one(e, a)) {
                U8(e, OKO_BUF_CRASHED)[a]   = 1;
                U8(e, OKO_BUF_TIMED_OUT)[a] = 1;
            }
        }
        /* 2) collision detection, Environment.cpp:145 */
        cast_one(e, a);
        /* app side: progress / reward / done */
        reward_one(e, a);
    }
}

void oko_fill_random_actions(OkoEnv *e, uint64_t step, uint32_t seed)
{
    static const float kThr[3]   = {-0.3f, 0.0f, 0.3f};            /* GeneticAgent.hpp:22-24 */
    static const float kSteer[5] = {-4.0f, -1.0f, 0.0f, 1.0f, 4.0f};
    float *thr = F32(e, OKO_BUF_ACT_THROTTLE), *st = F32(e, OKO_BUF_ACT_STEER);
    for (int64_t a = 0; a < e->n; ++a) {
        const uint64_t id = e->cfg.agent_id_base + (uint64_t)a;
        uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
        uint32_t key[2] = {seed, 0u};
        uint32_t o[4];
        oko_philox4x32_10(ctr, key, o);
        if (e->cfg.movement_mode == OKO_MOVE_ACCELERATION) {
            thr[a] = kThr[o[0] % 3u];
            st[a]  = kSteer[o[1] % 5u];
        } else {
            const float u0 = (float)(o[0] >> 8) * 0x1p-24f;
            const float u1 = (float)(o[1] >> 8) * 0x1p-24f;
            thr[a]         = 100.0f * u0;       /* ReinforceAgent.hpp:92-93 ranges */
            st[a]          = 10.0f * u1 - 5.0f;
        }
    }
}
