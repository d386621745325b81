// ok_kernels.cuh -- sm_100a kernels of the batched step loop.
//
// One fused kernel per tick (ok::step_kernel) replaces, per agent, the whole of
//   Environment::step stages 1-2            Environment.cpp:125-146
//     Agent::move                           Agent.cpp:82-98,108-119
//     checkAndUpdateStandstill              Environment.cpp:16-39
//     CollisionChecker pack/kernel/unpack   CollisionChecker.cu:8-71,113-172
//   + the app-side progress / reward / done RaceTrack.cpp:16-31,53-72 and SURVEY.md 8a row R
// with no host round trip.
//
// Mapping to the machine
//   * persistent CTAs, one per SM (the staged track fills most of the 227 KB of shared memory);
//     the agent range is cut into tiles of same-track agents, each CTA walks a contiguous run of
//     tiles and re-stages a track only when the tile's track changes;
//   * a track is staged with ONE TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) of its
//     blob: segments as float4, the uniform-grid cell table, the centre line;
//   * LPA = min(32, pow2ceil(R)) lanes serve one agent, one ray per lane (32/LPA agents per warp);
//     the per-agent work (kinematics, standstill) is computed redundantly by those lanes -- a warp
//     instruction costs the same for 1 or 32 active lanes -- and the per-agent reductions
//     (min hit distance, nearest centre-line index) are warp shuffles;
//   * state is structure-of-arrays; per-ray outputs are written coalesced (lane = ray).
// The path is latency/issue bound, not HBM bound, and is not a contraction: no tensor cores.
#pragma once

#include "ok_math.cuh"
#include "ok_track.hpp"

#include <cfloat>

namespace ok
{

struct Tile
{
    int32_t track;
    int32_t count;
    int64_t begin;
};

struct TrackRef
{
    uint64_t offset; // byte offset of the blob in the arena (128-byte aligned)
    uint32_t bytes;  // blob_bytes
    uint32_t pad;
};

struct StepParams
{
    // agent state, structure of arrays
    float    *x, *y, *rot, *speed, *accel, *act_thr, *act_steer;
    uint8_t  *crashed, *timed_out, *done;
    uint32_t *ss_ctr;
    float    *ss_x, *ss_y;
    int32_t  *track_id;
    float    *hit_abs, *hit_rel, *obs, *hit_t, *min_dist2;
    int32_t  *hit_seg, *nearest, *prev;
    float    *reward, *fitness;
    int32_t  *reset_pt;
    float    *start_x, *start_y;
    // static inputs
    const float   *ray_deg;
    const uint8_t *arena;
    const TrackRef *tracks;
    const Tile    *tiles;
    int32_t        n_tiles;
    int32_t        rays;
    int32_t        lanes_per_agent; // LPA
    // this launch
    const float *ext_thr, *ext_steer; // nullable: actions supplied by the caller
    int32_t      action_source;       // 0 stored/ext, 1 philox
    uint32_t     seed;
    uint64_t     step;
    int32_t      do_move;             // 0: ok_cast_rays
    // configuration
    int32_t  movement_mode, reward_mode, raycast_mode, auto_reset, auto_reset_stride;
    float    sensor_range, speed_limit, dt, collision_dist2, sensor_offset, standstill_thr2;
    uint32_t standstill_period;
};

// view of the staged blob
struct TrackView
{
    const float4   *seg;
    const uint2    *words;  // {occupancy bits, occupied-cell rank} per 32 cells
    const uint16_t *starts; // item offsets of the occupied cells
    const uint16_t *items;
    const float2   *pts;
    const float    *widths;
    const float    *headings;
    int32_t         n_pts, n_seg, nx, ny;
    float           gx0, gy0, cell, inv_cell;
};

__device__ __forceinline__ TrackView make_view(const uint8_t *blob)
{
    const TrackHeader *h = reinterpret_cast<const TrackHeader *>(blob);
    TrackView          v;
    v.seg      = reinterpret_cast<const float4 *>(blob + h->off_segments);
    v.words    = reinterpret_cast<const uint2 *>(blob + h->off_words);
    v.starts   = reinterpret_cast<const uint16_t *>(blob + h->off_starts);
    v.items    = reinterpret_cast<const uint16_t *>(blob + h->off_items);
    v.pts      = reinterpret_cast<const float2 *>(blob + h->off_points);
    v.widths   = reinterpret_cast<const float *>(blob + h->off_widths);
    v.headings = reinterpret_cast<const float *>(blob + h->off_headings);
    v.n_pts = h->n_points, v.n_seg = h->n_segments, v.nx = h->grid_nx, v.ny = h->grid_ny;
    v.gx0 = h->grid_x0, v.gy0 = h->grid_y0, v.cell = h->cell, v.inv_cell = h->inv_cell;
    return v;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(phase)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// ray x segment: raySegmentIntersect, CollisionChecker.cu:8-35, with the running-minimum rule of
// castRaysToSegmentsKernel (CollisionChecker.cu:49-66).
//
// The reference walks the segments in ascending index and accepts t <= min_t, so its result is
// (smallest valid t, HIGHEST index among the segments attaining it): an order-independent rule,
// which lets the broadphase visit candidates in any order.  `best` is that index (-1 = none).
//
// Two IEEE divisions per pair is the reference's cost; here a pair is first screened with
// comparisons on the numerators that are SUFFICIENT for the reference's predicate to fail
// (proofs in DESIGN.md "exact early rejection"); only survivors evaluate the literal predicate.
//   a = tn*sign(denom), b = sn*sign(denom), ad = |denom|, so t = a/ad, s = b/ad as real numbers.
//   b > ad              => RN(s) > 1            (RN(b/ad) <= 1 iff b <= ad for floats b, ad)
//   b < -2^-22          => RN(s) <= -2^-149 < 0 (|b/ad| > 2^-150 because ad < 2^128)
//   a < -2^-22          => RN(t) < 0            (same)
//   a > RN(RN(ad*m)*(1+2^-21)), RN(ad*m) >= 2^-100 => RN(t) > m (m = current min_t)
// NaN operands fail every screen and reach the literal predicate, which rejects them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void test_segment(const float4 sg, const int idx, const float ox, const float oy,
                                             const float dx, const float dy, float &min_t, int &best)
{
    const float    ex    = fsub(sg.x, ox);
    const float    ey    = fsub(sg.y, oy);
    const float    denom = fsub(fmul(dx, sg.w), fmul(dy, sg.z));
    const float    sn    = fsub(fmul(ex, dy), fmul(ey, dx));
    const uint32_t db    = __float_as_uint(denom);
    const uint32_t adb   = db & 0x7fffffffu;
    const uint32_t sgn   = db & 0x80000000u;
    const float    ad    = __uint_as_float(adb);
    const float    b     = __uint_as_float(__float_as_uint(sn) ^ sgn);
    // most candidates end here: parallel (fabsf(denom) < 1e-8f), or the ray's line misses the segment
    if ((adb < 0x322BCC77u) | (b > ad) | (b < -0x1p-22f))
        return;
    const float tn = fsub(fmul(ex, sg.w), fmul(ey, sg.z));
    const float a  = __uint_as_float(__float_as_uint(tn) ^ sgn);
    if (a < -0x1p-22f)
        return;
    const float lim = fmul(ad, min_t);
    if ((lim >= 0x1p-100f) & (a > fmul(lim, 1.00000047683715820312f)))
        return;
    // literal predicate of the reference
    const float t = __fdiv_rn(tn, denom);
    const float s = __fdiv_rn(sn, denom);
    if ((t >= 0.0f) && (s >= 0.0f) && (s <= 1.0f) && ((t < min_t) || ((t == min_t) && (idx > best))))
    {
        min_t = t;
        best  = idx;
    }
}

// every segment, ascending: the reference's own loop (OK_RAYCAST_BRUTE)
__device__ __forceinline__ void cast_ray_brute(const TrackView &tv, float ox, float oy, float dx, float dy, float range,
                                               float &min_t, int &best)
{
    min_t = range;
    best  = -1;
    for (int i = 0; i < tv.n_seg; ++i)
        test_segment(tv.seg[i], i, ox, oy, dx, dy, min_t, best);
}

// Uniform-grid broadphase: 2-D DDA over the cells the ray crosses, testing the segments
// registered in each cell, until the next cell starts beyond the current hit (+ slack).
// Exactness w.r.t. the brute-force loop is argued in DESIGN.md: segments are registered with a
// kGridMargin inflation that dominates every rounding error of this walk, and the per-pair
// arithmetic is the same function as above.
#define OK_DDA_SLACK 0.5f
__device__ __forceinline__ void cast_ray_grid(const TrackView &tv, float ox, float oy, float dx, float dy, float range,
                                              float &min_t, int &best)
{
    min_t = range;
    best  = -1;
    // a non-finite ray fails the reference predicate on every segment
    if (!(fabsf(ox) <= FLT_MAX && fabsf(oy) <= FLT_MAX && fabsf(dx) <= FLT_MAX && fabsf(dy) <= FLT_MAX))
        return;
    const float gx1 = tv.gx0 + tv.nx * tv.cell;
    const float gy1 = tv.gy0 + tv.ny * tv.cell;
    float       t0 = 0.0f, t1 = range + OK_DDA_SLACK;
    float       inv_dx = 0.0f, inv_dy = 0.0f;
    if (dx != 0.0f)
    {
        inv_dx         = 1.0f / dx;
        const float ta = (tv.gx0 - ox) * inv_dx, tb = (gx1 - ox) * inv_dx;
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    else if (ox < tv.gx0 || ox > gx1)
        return;
    if (dy != 0.0f)
    {
        inv_dy         = 1.0f / dy;
        const float ta = (tv.gy0 - oy) * inv_dy, tb = (gy1 - oy) * inv_dy;
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    else if (oy < tv.gy0 || oy > gy1)
        return;
    if (!(t0 <= t1))
        return; // the ray does not reach the grid within range

    const float px = ox + t0 * dx, py = oy + t0 * dy; // entry point
    int         ix = static_cast<int>(floorf((px - tv.gx0) * tv.inv_cell));
    int         iy = static_cast<int>(floorf((py - tv.gy0) * tv.inv_cell));
    ix             = min(max(ix, 0), tv.nx - 1);
    iy             = min(max(iy, 0), tv.ny - 1);
    const int   sx = dx > 0.0f ? 1 : -1, sy = dy > 0.0f ? 1 : -1;
    const float inf = __int_as_float(0x7f800000);
    float       tmx = inf, tmy = inf, tdx = inf, tdy = inf;
    if (dx != 0.0f)
    {
        tmx = (tv.gx0 + (ix + (dx > 0.0f ? 1 : 0)) * tv.cell - ox) * inv_dx;
        tdx = tv.cell * fabsf(inv_dx);
    }
    if (dy != 0.0f)
    {
        tmy = (tv.gy0 + (iy + (dy > 0.0f ? 1 : 0)) * tv.cell - oy) * inv_dy;
        tdy = tv.cell * fabsf(inv_dy);
    }
    // One loop, one unit of work per trip: a lane either tests the next segment of its current cell
    // or moves to the next cell.  (A cell loop around a segment loop leaves most lanes idle: the
    // 32 rays of a fan sit in cells with very different populations -- measured 5.7 active lanes.)
    uint32_t k = 0, k_end = 0;
    {
        const int   c = iy * tv.nx + ix;
        const uint2 w = tv.words[c >> 5];
        if ((w.x >> (c & 31)) & 1u)
        {
            const uint32_t r = w.y + __popc(w.x & ((1u << (c & 31)) - 1u));
            k = tv.starts[r], k_end = tv.starts[r + 1];
        }
    }
    for (;;)
    {
        if (k < k_end)
        {
            const int i = tv.items[k++];
            test_segment(tv.seg[i], i, ox, oy, dx, dy, min_t, best);
        }
        else
        {
            const float t_next = fminf(tmx, tmy);
            if (!(t_next <= fminf(min_t + OK_DDA_SLACK, t1)))
                break;
            if (tmx < tmy)
            {
                ix += sx;
                tmx += tdx;
                if (static_cast<unsigned>(ix) >= static_cast<unsigned>(tv.nx))
                    break;
            }
            else
            {
                iy += sy;
                tmy += tdy;
                if (static_cast<unsigned>(iy) >= static_cast<unsigned>(tv.ny))
                    break;
            }
            const int   c = iy * tv.nx + ix;
            const uint2 w = tv.words[c >> 5];
            if ((w.x >> (c & 31)) & 1u)
            {
                const uint32_t r = w.y + __popc(w.x & ((1u << (c & 31)) - 1u));
                k = tv.starts[r], k_end = tv.starts[r + 1];
            }
        }
    }
}

// RaceTrack::findNearestTrackIndexBruteForce (RaceTrack.cpp:16-31), cooperatively by the `lpa`
// lanes of an agent: strict '<' keeps the lowest index on ties, so the combine is the
// lexicographic minimum of (distance, index) -- order independent, hence exact.
__device__ __forceinline__ int nearest_index(const TrackView &tv, float qx, float qy, int sub, int lpa, float &d2_out)
{
    float best = FLT_MAX;
    int   bi   = 0;
    for (int i = sub; i < tv.n_pts; i += lpa)
    {
        const float2 p  = tv.pts[i];
        const float  ddx = fsub(qx, p.x), ddy = fsub(qy, p.y);
        const float  d  = fadd(fmul(ddx, ddx), fmul(ddy, ddy)); // Vec2d::distanceSquared, Typedefs.h:40-43
        if (d < best)
        {
            best = d;
            bi   = i;
        }
    }
    for (int o = lpa >> 1; o > 0; o >>= 1)
    {
        const float od = __shfl_xor_sync(0xffffffffu, best, o);
        const int   oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < best || (od == best && oi < bi))
        {
            best = od;
            bi   = oi;
        }
    }
    d2_out = best;
    return bi;
}

// ---------------------------------------------------------------------------------------------
// the tick
// ---------------------------------------------------------------------------------------------
template <int kBlock> __global__ void __launch_bounds__(kBlock, 1) step_kernel(const StepParams p)
{
    extern __shared__ __align__(128) uint8_t blob[];
    __shared__ __align__(8) uint64_t         bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kBlock / 32;
    const int lpa = p.lanes_per_agent, sub = lane & (lpa - 1), grp = lane / lpa, groups = 32 / lpa;

    if (tid == 0)
        mbar_init(&bar, 1);
    __syncthreads();

    const int tile_begin = static_cast<int>((static_cast<int64_t>(blockIdx.x) * p.n_tiles) / gridDim.x);
    const int tile_end   = static_cast<int>((static_cast<int64_t>(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    int       staged = -1;
    uint32_t  phase  = 0;

    for (int tile = tile_begin; tile < tile_end; ++tile)
    {
        const Tile tl = p.tiles[tile];
        if (tl.track != staged)
        {
            __syncthreads(); // everyone is done with the previous track
            if (tid == 0)
            {
                const TrackRef tr = p.tracks[tl.track];
                fence_proxy_async();
                mbar_expect_tx(&bar, tr.bytes);
                tma_bulk_g2s(blob, p.arena + tr.offset, tr.bytes, &bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            staged = tl.track;
        }
        const TrackView tv = make_view(blob);

        // warp w serves agents (w*groups + grp) + j*(kWarps*groups) of the tile; the trip count is
        // warp-uniform so the shuffles below always see all 32 lanes
        const int per_pass = kWarps * groups;
        for (int base = 0; base < tl.count; base += per_pass)
        {
            const int     local = base + warp * groups + grp;
            const bool    valid = local < tl.count;
            const int64_t a     = tl.begin + (valid ? local : 0);

            // ---- load ----
            float   x = p.x[a], y = p.y[a], rot = p.rot[a], speed = p.speed[a], accel = p.accel[a];
            bool    crashed = p.crashed[a] != 0, timed_out = p.timed_out[a] != 0;
            float   fitness = p.fitness[a];
            int32_t prev    = p.prev[a];

            if (p.do_move)
            {
                float thr, steer;
                if (p.action_source == 1)
                { // synthetic stream: Philox4x32-10, counter (agent, step), key (seed, 0)
                    uint32_t o[4];
                    philox4x32_10(static_cast<uint32_t>(a), static_cast<uint32_t>(static_cast<uint64_t>(a) >> 32),
                                  static_cast<uint32_t>(p.step), static_cast<uint32_t>(p.step >> 32), p.seed, 0u, o);
                    if (p.movement_mode == 1)
                    { // GeneticAgent.hpp:22-24 action set
                        const uint32_t ti = o[0] % 3u, si = o[1] % 5u;
                        thr   = ti == 0 ? -0.3f : (ti == 1 ? 0.0f : 0.3f);
                        steer = si == 0 ? -4.0f : (si == 1 ? -1.0f : (si == 2 ? 0.0f : (si == 3 ? 1.0f : 4.0f)));
                    }
                    else
                    { // ReinforceAgent.hpp:92-93 ranges
                        thr   = fmul(100.0f, fmul(static_cast<float>(o[0] >> 8), 0x1p-24f));
                        steer = fsub(fmul(10.0f, fmul(static_cast<float>(o[1] >> 8), 0x1p-24f)), 5.0f);
                    }
                }
                else if (p.ext_thr)
                {
                    thr   = p.ext_thr[a];
                    steer = p.ext_steer[a];
                }
                else
                {
                    thr   = p.act_thr[a];
                    steer = p.act_steer[a];
                }

                // ---- optional reset of a crashed agent before the tick (GuidedCostLearning/test.cpp:102-111):
                //      Environment::resetAgent + Agent::reset, Environment.cpp:103-121 / Agent.cpp:123-135 ----
                const bool do_reset = p.auto_reset && crashed; // uniform within the agent's lanes
                if (__any_sync(0xffffffffu, do_reset))
                {
                    int32_t pt = 0;
                    float   rx = x, ry = y;
                    if (do_reset)
                    {
                        pt = static_cast<int32_t>((static_cast<int64_t>(p.reset_pt[a]) + p.auto_reset_stride) % tv.n_pts);
                        const float2 c = tv.pts[pt];
                        rx = c.x, ry = c.y;
                    }
                    float     d2;
                    const int near0 = nearest_index(tv, rx, ry, sub, lpa, d2); // all lanes take part
                    if (do_reset)
                    {
                        x = rx, y = ry, rot = tv.headings[pt];
                        accel = 0.0f, speed = 0.0f;
                        crashed = false, timed_out = false;
                        thr = 0.0f, steer = 0.0f; // Agent::reset zeroes current_action_
                        prev    = near0;
                        fitness = 0.0f;
                        if (valid && sub == 0)
                        {
                            p.reset_pt[a] = pt;
                            p.start_x[a]  = rx;
                            p.start_y[a]  = ry;
                            p.nearest[a]  = near0;
                        }
                    }
                }
                if (valid && sub == 0)
                {
                    p.act_thr[a]   = thr;
                    p.act_steer[a] = steer;
                }

                // ---- 1) kinematics + standstill, Environment.cpp:128-143 ----
                if (!crashed)
                {
                    rot = fadd(rot, steer);
                    if (p.movement_mode == 1)
                    { // moveViaAcceleration, Agent.cpp:82-98
                        accel = fadd(accel, thr);
                        speed = fadd(speed, fmul(accel, p.dt));
                        speed = (speed < 0.0f) ? 0.0f : speed;
                        speed = (speed > p.speed_limit) ? p.speed_limit : speed;
                    }
                    else
                    { // moveViaVelocity, Agent.cpp:108-119
                        speed = thr;
                    }
                    float ms, mc;
                    sincosf(fmul(OK_DEG2RAD, rot), ms, mc);
                    x = fadd(x, fmul(fmul(mc, speed), p.dt));
                    y = fadd(y, fmul(fmul(ms, speed), p.dt));

                    // checkAndUpdateStandstill, Environment.cpp:16-39
                    uint32_t ctr = p.ss_ctr[a];
                    bool     out = false;
                    if (ctr == 0)
                    {
                        if (valid && sub == 0)
                        {
                            p.ss_x[a] = x;
                            p.ss_y[a] = y;
                        }
                        ctr = 1;
                    }
                    else if (ctr >= p.standstill_period)
                    {
                        const float mx = fsub(x, p.ss_x[a]), my = fsub(y, p.ss_y[a]);
                        out = fadd(fmul(mx, mx), fmul(my, my)) < p.standstill_thr2;
                        ctr = 0;
                    }
                    else
                        ++ctr;
                    if (valid && sub == 0)
                        p.ss_ctr[a] = ctr;
                    if (out)
                    {
                        crashed   = true;
                        timed_out = true;
                    }
                }
            }

            // ---- 2) lidar: pack (CollisionChecker.cu:115-128), cast (:37-71), unpack (:144-172) ----
            float rs, rc;
            sincosf(fmul(OK_DEG2RAD, rot), rs, rc);
            const float ox   = fadd(x, fmul(p.sensor_offset, rc));
            const float oy   = fadd(y, fmul(p.sensor_offset, rs));
            const bool  live = !crashed;
            float       min_d2 = fmul(p.sensor_range, p.sensor_range);
            for (int r0 = 0; r0 < p.rays; r0 += lpa)
            {
                const int     r      = r0 + sub;
                const bool    has    = valid && r < p.rays;
                const int64_t ri     = a * p.rays + (has ? r : 0);
                float2        hit;
                if (live)
                {
                    float dx, dy;
                    sincosf(fmul(OK_DEG2RAD, fadd(rot, p.ray_deg[has ? r : 0])), dy, dx);
                    float min_t;
                    int   best;
                    if (p.raycast_mode == 0)
                        cast_ray_grid(tv, ox, oy, dx, dy, p.sensor_range, min_t, best);
                    else
                        cast_ray_brute(tv, ox, oy, dx, dy, p.sensor_range, min_t, best);
                    hit.x = fadd(ox, fmul(min_t, dx));
                    hit.y = fadd(oy, fmul(min_t, dy));
                    if (has)
                    {
                        reinterpret_cast<float2 *>(p.hit_abs)[ri] = hit;
                        p.hit_t[ri]                               = min_t;
                        p.hit_seg[ri]                             = best;
                    }
                }
                else
                    hit = reinterpret_cast<const float2 *>(p.hit_abs)[ri]; // stale hits of a crashed agent
                const float xt = fsub(hit.x, ox), yt = fsub(hit.y, oy);
                float2      rel;
                rel.x          = fsub(fmul(xt, rc), fmul(yt, rs));
                rel.y          = fadd(fmul(xt, rs), fmul(yt, rc));
                const float sq = fadd(fmul(rel.x, rel.x), fmul(rel.y, rel.y));
                if (has)
                {
                    reinterpret_cast<float2 *>(p.hit_rel)[ri] = rel;
                    p.obs[ri]                                 = __fdiv_rn(__fsqrt_rn(sq), p.sensor_range);
                    if (sq < min_d2)
                        min_d2 = sq;
                }
            }
            for (int o = lpa >> 1; o > 0; o >>= 1)
                min_d2 = fminf(min_d2, __shfl_xor_sync(0xffffffffu, min_d2, o));
            if (min_d2 < p.collision_dist2)
                crashed = true;

            // ---- app side: progress / reward / done ----
            const int mode      = p.reward_mode;
            const bool need_idx = mode == 1 || mode == 2 || mode == 6 || mode == 7;
            int32_t   near      = 0;
            float     near_d2   = 0.0f;
            if (need_idx && p.do_move)
                near = nearest_index(tv, x, y, sub, lpa, near_d2);
            if (valid && sub == 0)
            {
                if (p.do_move)
                {
                    float reward = 0.0f;
                    switch (mode)
                    {
                    case 1: // QAgent.hpp:150-168
                        if (crashed)
                            reward = -200.0f;
                        else
                        {
                            const int32_t prog = near - prev;
                            prev               = near;
                            const int32_t ab = prog < 0 ? -prog : prog, len = tv.n_pts;
                            reward = static_cast<float>(ab > (len / 2) ? len - ab : ab);
                        }
                        break;
                    case 2: // main_eigen.cpp:147-163
                        if (!crashed)
                        {
                            const int32_t prog = near - prev;
                            prev               = near;
                            reward             = static_cast<float>(prog < 0 ? -prog : prog);
                            fitness            = fadd(fitness, reward);
                        }
                        else if (timed_out)
                            fitness = 0.0f;
                        break;
                    case 3: reward = 1.0f; break; // ppo_sim.cpp:76
                    case 4:                       // ReinforceContinuous/reinforce_sim.cpp:59-73
                    {
                        const float ddx = fsub(x, p.start_x[a]), ddy = fsub(y, p.start_y[a]);
                        reward = crashed ? -5.0f : __fsqrt_rn(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
                        break;
                    }
                    case 5: // DQAgent.hpp:161-180: min over rays of norm(), which is sqrt of the min squared norm
                    {
                        const float m = __fsqrt_rn(min_d2);
                        reward        = crashed ? -200.0f : (p.sensor_range > m ? m : p.sensor_range);
                        break;
                    }
                    case 6: reward = static_cast<float>(near); break; // MiscUtils.hpp:64-71
                    case 7:                                           // WorldModelVaeRnn/main.cpp:336-342
                        if (!crashed)
                        {
                            reward  = fsub(1.0f, __fdiv_rn(__fsqrt_rn(near_d2), tv.widths[near])); // RaceTrack.cpp:53-72
                            fitness = fadd(fitness, reward);
                        }
                        else if (timed_out)
                            fitness = 0.0f;
                        break;
                    default: break;
                    }
                    p.x[a] = x, p.y[a] = y, p.rot[a] = rot, p.speed[a] = speed, p.accel[a] = accel;
                    p.timed_out[a] = timed_out;
                    p.reward[a]    = reward;
                    p.fitness[a]   = fitness;
                    p.prev[a]      = prev;
                    if (need_idx)
                        p.nearest[a] = near;
                }
                p.crashed[a]   = crashed;
                p.done[a]      = crashed; // Agent::isDone, Agent.cpp:138-144 (completed_ is never set)
                p.min_dist2[a] = min_d2;
            }
        }
    }
}

// Environment::resetAgent (Environment.cpp:79-122) + Agent::reset (Agent.cpp:123-135), one thread
// per request, reading the track blobs from global memory.
struct ResetParams
{
    const int64_t *agent_idx; // nullable
    const int32_t *pt_idx;
    const float   *lane_alpha;  // nullable
    const float   *heading_off; // nullable
    int64_t        n, n_agents;
};

__global__ void reset_kernel(const StepParams p, const ResetParams r)
{
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= r.n)
        return;
    const int64_t a = r.agent_idx ? r.agent_idx[k] : k;
    if (a < 0 || a >= r.n_agents)
        return;
    const uint8_t    *blob = p.arena + p.tracks[p.track_id[a]].offset;
    const TrackView   tv   = make_view(blob);
    const TrackHeader *h   = reinterpret_cast<const TrackHeader *>(blob);
    int32_t           pt   = r.pt_idx[k];
    pt                     = min(max(pt, 0), tv.n_pts - 1);
    float sx, sy;
    if (r.lane_alpha)
    { // Environment.cpp:107-114; LI / RI points are the start points of the LI / RI segment chains
        const float   alpha = r.lane_alpha[k];
        const int32_t n     = tv.n_pts;
        float2        l, rr;
        if (pt < n - 1)
        {
            const float4 ls = tv.seg[pt], rs = tv.seg[2 * (n - 1) + pt];
            l = make_float2(ls.x, ls.y), rr = make_float2(rs.x, rs.y);
        }
        else
        { // the last point starts the closure segments
            const float4 ls = tv.seg[4 * (n - 1)], rs = tv.seg[4 * (n - 1) + 1];
            l = make_float2(ls.x, ls.y), rr = make_float2(rs.x, rs.y);
        }
        const float one_minus = fsub(1.0f, alpha);
        sx = fadd(fmul(l.x, alpha), fmul(rr.x, one_minus));
        sy = fadd(fmul(l.y, alpha), fmul(rr.y, one_minus));
    }
    else
    {
        const float2 c = tv.pts[pt];
        sx = c.x, sy = c.y;
    }
    (void)h;
    const float off = r.heading_off ? r.heading_off[k] : 0.0f;
    p.x[a] = sx, p.y[a] = sy;
    p.rot[a]       = fadd(tv.headings[pt], off);
    p.accel[a]     = 0.0f;
    p.speed[a]     = 0.0f;
    p.crashed[a]   = 0;
    p.timed_out[a] = 0;
    p.done[a]      = 0;
    p.act_thr[a]   = 0.0f;
    p.act_steer[a] = 0.0f;
    p.reset_pt[a]  = pt;
    p.start_x[a]   = sx;
    p.start_y[a]   = sy;
    float   d2;
    int32_t near  = nearest_index(tv, sx, sy, 0, 1, d2);
    p.prev[a]     = near;
    p.nearest[a]  = near;
    p.fitness[a]  = 0.0f;
    p.reward[a]   = 0.0f;
}

__global__ void fill_actions_kernel(const StepParams p, int64_t n)
{
    const int64_t a = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (a >= n)
        return;
    uint32_t o[4];
    philox4x32_10(static_cast<uint32_t>(a), static_cast<uint32_t>(static_cast<uint64_t>(a) >> 32),
                  static_cast<uint32_t>(p.step), static_cast<uint32_t>(p.step >> 32), p.seed, 0u, o);
    float thr, steer;
    if (p.movement_mode == 1)
    {
        const uint32_t ti = o[0] % 3u, si = o[1] % 5u;
        thr   = ti == 0 ? -0.3f : (ti == 1 ? 0.0f : 0.3f);
        steer = si == 0 ? -4.0f : (si == 1 ? -1.0f : (si == 2 ? 0.0f : (si == 3 ? 1.0f : 4.0f)));
    }
    else
    {
        thr   = fmul(100.0f, fmul(static_cast<float>(o[0] >> 8), 0x1p-24f));
        steer = fsub(fmul(10.0f, fmul(static_cast<float>(o[1] >> 8), 0x1p-24f)), 5.0f);
    }
    p.act_thr[a]   = thr;
    p.act_steer[a] = steer;
}

} // namespace ok
