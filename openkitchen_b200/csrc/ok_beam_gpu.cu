// ok_beam_gpu.cu -- the beam-table builder of ok_beam.hpp as a kernel: one CTA per covered cell (row).
//
// Same construction as the host builder (ok_beam.cpp, build_row), restated for a CTA:
//   1. candidates  = segments within reach of the cell, with a lower bound `lb` of their distance and a conservative
//                    range of direction bins (threads stride over the segments, list in global scratch);
//   2. completeness distance per bin from 15 sample rays (threads stride over the candidates, each trying the sample
//                    rays of the bins it can be seen in; first hits meet in shared-memory minima);
//   3. membership  = exact distance from the origin to (S (+) -C) clipped to the bin's cone, binary64, for every
//                    (candidate, bin) the lower bound cannot rule out (threads stride over candidates);
//   4. per bin: sort by (distance, segment), pad to chunks of four, append to the item array (one global atomic per
//      list), write the entry.
// A list that does not fit the per-bin scratch is written as "decides nothing" (count 0, d = 0): the kernel then
// walks the grid for those rays -- slower, never wrong.
#include "ok_beam.hpp"
#include "ok_beam_geom.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace ok
{
namespace
{
using namespace beamgeom;

constexpr int kThreads = 256;
constexpr int kListCap = 256; // candidates a (cell, bin) list can hold while it is being built

struct GpuBuild
{
    const float4 *seg; // x1, y1, x2, y2
    int32_t       ns;
    double        x0, y0, h, rb, range, pad, dth;
    int32_t       nx, nb;
    const uint32_t *covered;
    int32_t         n_rows;
    // scratch, per CTA
    float    *cand_lb;
    uint16_t *cand_seg, *cand_b0, *cand_bn, *cand_act;
    float    *list_d;
    uint16_t *list_s;
    // output
    uint4              *entries;
    uint16_t           *items;
    unsigned long long *item_cursor;
    unsigned long long  item_capacity;
    int32_t            *row_cursor;
    int32_t            *overflow; // [0]: lists that did not fit the scratch, [1]: item array full
};

__global__ void __launch_bounds__(kThreads) beam_build_kernel(const GpuBuild g)
{
    extern __shared__ unsigned char smem_raw[];
    double       *dir_xy   = reinterpret_cast<double *>(smem_raw);               // [2 * nb][2] sample directions
    unsigned int *hit_bits = reinterpret_cast<unsigned int *>(dir_xy + 4 * g.nb); // [2 * nb][5] first hit per sample ray
    float        *dcomp    = reinterpret_cast<float *>(hit_bits + 10 * g.nb); // [nb]
    int          *bin_cnt  = reinterpret_cast<int *>(dcomp + g.nb);           // [nb]
    double       *edge     = reinterpret_cast<double *>(bin_cnt + g.nb);      // [nb][4]: the widened cone's two half-plane normals
    __shared__ int s_row, s_ncand, s_nact;
    // candidates are shared by GROUPS of kGroup lanes (the work per candidate -- bins it can be seen in -- varies from 3 to
    // all of them: a thread per candidate left 7 of 32 lanes busy and 69 % of the warp time at the barriers)
    constexpr int kGroup = 8, kGroups = kThreads / kGroup;
    const int     gid = threadIdx.x / kGroup, sub = threadIdx.x % kGroup;

    const int    tid   = threadIdx.x;
    const size_t slot  = blockIdx.x;
    float       *c_lb  = g.cand_lb + slot * g.ns;
    uint16_t    *c_seg = g.cand_seg + slot * g.ns, *c_b0 = g.cand_b0 + slot * g.ns, *c_bn = g.cand_bn + slot * g.ns;
    uint16_t    *c_act = g.cand_act + slot * g.ns;
    float       *l_d   = g.list_d + slot * static_cast<size_t>(g.nb) * kListCap;
    uint16_t    *l_s   = g.list_s + slot * static_cast<size_t>(g.nb) * kListCap;
    const double two_pi = 2.0 * M_PI, dbin = two_pi / g.nb;
    for (int j = tid; j < 2 * g.nb; j += kThreads)
    { // bin edges (even j) and bin centres (odd j)
        dir_xy[2 * j]     = cos(0.5 * j * dbin);
        dir_xy[2 * j + 1] = sin(0.5 * j * dbin);
    }
    for (int b = tid; b < g.nb; b += kThreads)
    { // half planes of bin b's cone, widened by dth on both sides (the same expressions the per-pair code used)
        edge[4 * b]     = cos(b * dbin - g.dth);
        edge[4 * b + 1] = sin(b * dbin - g.dth);
        edge[4 * b + 2] = -cos((b + 1) * dbin + g.dth);
        edge[4 * b + 3] = -sin((b + 1) * dbin + g.dth);
    }
    const float miss = static_cast<float>(2.0 * g.rb);

    for (;;)
    {
        __syncthreads();
        if (tid == 0)
        {
            s_row   = atomicAdd(g.row_cursor, 1);
            s_ncand = 0;
            s_nact  = 0;
        }
        __syncthreads();
        const int row = s_row;
        if (row >= g.n_rows)
            break;
        const uint32_t cell = g.covered[row];
        const int      cx = static_cast<int>(cell % g.nx), cy = static_cast<int>(cell / g.nx);
        const double   m   = kBeamCellMargin;
        const double   bx0 = g.x0 + cx * g.h - m, bx1 = g.x0 + (cx + 1) * g.h + m;
        const double   by0 = g.y0 + cy * g.h - m, by1 = g.y0 + (cy + 1) * g.h + m;
        const double   ccx = 0.5 * (bx0 + bx1), ccy = 0.5 * (by0 + by1);
        const double   half_diag = 0.5 * hypot(bx1 - bx0, by1 - by0);
        for (int i = tid; i < 10 * g.nb; i += kThreads)
            hit_bits[i] = __float_as_uint(miss);
        for (int i = tid; i < g.nb; i += kThreads)
            bin_cnt[i] = 0;

        // ---- 1. candidates ----
        for (int s = tid; s < g.ns; s += kThreads)
        {
            const float4 sg = g.seg[s];
            const double mx = 0.5 * (static_cast<double>(sg.x) + sg.z), my = 0.5 * (static_cast<double>(sg.y) + sg.w);
            const double hl = 0.5 * hypot(static_cast<double>(sg.z) - sg.x, static_cast<double>(sg.w) - sg.y);
            const double vx = mx - ccx, vy = my - ccy;
            const double dist = hypot(vx, vy), rho = half_diag + hl;
            if (dist - rho > g.rb)
                continue;
            const double d = point_segment_distance(ccx, ccy, sg.x, sg.y, sg.z, sg.w) - half_diag;
            if (d > g.rb)
                continue;
            int b0 = 0, bn = g.nb;
            if (dist > rho * 1.0000001)
            { // q - p lies in the disc of radius rho around (midpoint - cell centre)
                const double phi = atan2(vy, vx), alpha = asin(fmin(1.0, rho / dist)) + 2.0 * g.dth;
                const double f_lo = floor((phi - alpha) / dbin), f_hi = floor((phi + alpha) / dbin);
                bn = static_cast<int>(fmin(static_cast<double>(g.nb), f_hi - f_lo + 1.0));
                b0 = static_cast<int>(fmod(fmod(f_lo, static_cast<double>(g.nb)) + g.nb, static_cast<double>(g.nb)));
            }
            const int k = atomicAdd(&s_ncand, 1);
            c_lb[k]     = static_cast<float>(fmax(0.0, d) * (1.0 - 1e-6)); // rounded down: stays a lower bound in binary32
            c_seg[k]    = static_cast<uint16_t>(s);
            c_b0[k]     = static_cast<uint16_t>(b0);
            c_bn[k]     = static_cast<uint16_t>(bn);
        }
        __syncthreads();
        const int ncand = s_ncand;

        // ---- 2. completeness distance: rays from the four corners and the centre, bin edges and centres.  A thread
        // takes a candidate and tries the sample rays of the bins it can be seen in; the first hit of a ray is a
        // shared-memory minimum (bit patterns of non-negative floats order like the floats) ----
        for (int c = gid; c < ncand; c += kGroups)
        { // a group per candidate, its lanes over the (sample direction, origin) pairs
            const float4 sg = g.seg[c_seg[c]];
            const double sx = static_cast<double>(sg.z) - sg.x, sy = static_cast<double>(sg.w) - sg.y;
            const int    b0 = c_b0[c], bn = c_bn[c];
            for (int m = sub; m < 10 * bn; m += kGroup)
            {
                const int    k = m / 5, o = m - 5 * k;
                const int    j  = (2 * b0 + k) & (2 * g.nb - 1);
                const double dx = dir_xy[2 * j], dy = dir_xy[2 * j + 1];
                const double den = dx * sy - dy * sx;
                if (fabs(den) < 1e-12)
                    continue;
                const double ox = o == 4 ? ccx : ((o & 1) ? bx1 : bx0), oy = o == 4 ? ccy : ((o & 2) ? by1 : by0);
                const double ex = sg.x - ox, ey = sg.y - oy;
                const double tt = (ex * sy - ey * sx) / den, ss = (ex * dy - ey * dx) / den;
                if (tt >= 0 && ss >= 0 && ss <= 1 && tt < miss)
                { // rounded UP to binary32: the sampled distance is never understated
                    float bf = static_cast<float>(tt);
                    if (static_cast<double>(bf) < tt)
                        bf = nextafterf(bf, 3.0e38f);
                    atomicMin(&hit_bits[5 * j + o], __float_as_uint(bf));
                }
            }
        }
        __syncthreads();
        for (int b = tid; b < g.nb; b += kThreads)
        {
            float far = 0.0f;
            for (int e = 0; e < 3; ++e)
            {
                const int j = (2 * b + e) & (2 * g.nb - 1);
                for (int o = 0; o < 5; ++o)
                    far = fmaxf(far, __uint_as_float(hit_bits[5 * j + o]));
            }
            double d = static_cast<double>(far) + g.pad;
            if (far >= g.range || d >= g.rb)
                d = g.rb;
            dcomp[b] = static_cast<float>(d);
        }
        __syncthreads();

        // ---- 3. membership ----
        for (int c = tid; c < ncand; c += kThreads)
        { // the candidates that can be in any list at all (most are beyond every completeness distance)
            const float lb = c_lb[c];
            const int   b0 = c_b0[c], bn = c_bn[c];
            bool        any = false;
            for (int k = 0; k < bn && !any; ++k)
                any = lb <= dcomp[(b0 + k) & (g.nb - 1)];
            if (any)
                c_act[atomicAdd(&s_nact, 1)] = static_cast<uint16_t>(c);
        }
        __syncthreads();
        const int nact = s_nact;
        for (int i = gid; i < nact; i += kGroups)
        { // a group per candidate: every lane builds the (same) hull, then the lanes take the bins
            const int    c  = c_act[i];
            const float  lb = c_lb[c];
            const int    b0 = c_b0[c], bn = c_bn[c];
            const float4 sg = g.seg[c_seg[c]];
            V2           pts[8], hull[9];
            int          np = 0;
            for (int e = 0; e < 2; ++e)
                for (int k = 0; k < 4; ++k)
                    pts[np++] = {(e ? sg.z : sg.x) - ((k & 1) ? bx1 : bx0), (e ? sg.w : sg.y) - ((k & 2) ? by1 : by0)};
            const int nh = convex_hull8(pts, np, hull);
            for (int k = sub; k < bn; k += kGroup)
            {
                const int b = (b0 + k) & (g.nb - 1);
                if (lb > dcomp[b])
                    continue;
                const V2 lo = {edge[4 * b], edge[4 * b + 1]};
                const V2 hi = {edge[4 * b + 2], edge[4 * b + 3]};
                V2       c1[12], c2[14];
                const int    n1 = clip_half_plane(hull, nh, lo, c1);
                const int    n2 = clip_half_plane(c1, n1, hi, c2);
                const double d  = origin_distance(c2, n2);
                if (d <= dcomp[b])
                {
                    const int slot_k = atomicAdd(&bin_cnt[b], 1);
                    if (slot_k < kListCap)
                    {
                        l_d[static_cast<size_t>(b) * kListCap + slot_k] = static_cast<float>(d);
                        l_s[static_cast<size_t>(b) * kListCap + slot_k] = c_seg[c];
                    }
                }
            }
        }
        __syncthreads();

        // ---- 4. sort, pad, append, entry ----
        for (int b = tid; b < g.nb; b += kThreads)
        {
            int       cnt = bin_cnt[b];
            float    *dd  = l_d + static_cast<size_t>(b) * kListCap;
            uint16_t *ss  = l_s + static_cast<size_t>(b) * kListCap;
            double    d   = dcomp[b];
            if (cnt > kListCap)
            { // the scratch holds an arbitrary subset: this list decides nothing
                atomicAdd(&g.overflow[0], 1);
                cnt = 0;
                d   = 0.0;
            }
            for (int i = 1; i < cnt; ++i)
            { // insertion sort by (distance, segment)
                const float    vd = dd[i];
                const uint16_t vs = ss[i];
                int            j  = i - 1;
                while (j >= 0 && (dd[j] > vd || (dd[j] == vd && ss[j] > vs)))
                {
                    dd[j + 1] = dd[j];
                    ss[j + 1] = ss[j];
                    --j;
                }
                dd[j + 1] = vd;
                ss[j + 1] = vs;
            }
            // the first four candidates travel in the entry itself; the rest goes to the item array in chunks of four
            const int                rest   = cnt > kBeamInline ? cnt - kBeamInline : 0; // <= kListCap - 4 < 4 * kBeamMaxRest
            const unsigned long long padded = static_cast<unsigned long long>((rest + 3) & ~3);
            unsigned long long       off    = padded ? atomicAdd(g.item_cursor, padded) : 0ull;
            if (off + padded > g.item_capacity)
            {
                atomicAdd(&g.overflow[1], 1);
                cnt = 0;
                d   = 0.0;
                off = 0;
            }
            else
            {
                for (int i = 0; i < static_cast<int>(padded); ++i)
                    g.items[off + i] = i < rest ? ss[kBeamInline + i] : static_cast<uint16_t>(g.ns); // the blob's null segment
            }
            const uint32_t null_seg = static_cast<uint32_t>(g.ns);
            uint32_t       in4[kBeamInline];
            for (int i = 0; i < kBeamInline; ++i)
                in4[i] = i < cnt ? ss[i] : null_seg;
            const int    n_rest = cnt > kBeamInline ? static_cast<int>(padded / 4) : 0;
            const double d1     = cnt > kBeamInline ? fmin(static_cast<double>(dd[kBeamInline]), d) : d;
            g.entries[static_cast<size_t>(row) * g.nb + b] =
                make_uint4(in4[0] | (in4[1] << 16), in4[2] | (in4[3] << 16), static_cast<uint32_t>(off / 4),
                           beam_pack_meta(d, d1, d >= g.rb, static_cast<uint32_t>(n_rest)));
        }
    }
}

#define BEAM_CUDA(expr)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
        {                                                                                                              \
            err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                                 \
            goto done;                                                                                                 \
        }                                                                                                              \
    } while (0)
} // namespace

// The builder's three large scratch areas (candidate / list scratch, entries, items: ~0.6 GB together at the default
// resolution) are kept between tracks: a process builds 23 tables back to back, and cudaMalloc / cudaFree of that much
// memory per track cost about as much as a third of the kernel.  beam_builder_release() (ok_release_caches) frees them.
namespace
{
struct Pool
{
    void  *ptr{nullptr};
    size_t bytes{0};
    int    device{-1};
};
Pool       g_pool[8]; // scratch, entries, items, segments, covered cells, counters, second entries, second items (pipelined builds)
std::mutex g_pool_mu;

cudaError_t pool_get(int which, int device, size_t bytes, void **out)
{ // caller holds g_pool_mu and has made `device` current
    Pool &p = g_pool[which];
    if (p.ptr && (p.device != device || p.bytes < bytes))
    {
        cudaFree(p.ptr);
        p = Pool{};
    }
    if (!p.ptr)
    { // a quarter of headroom: the next, slightly larger track must not trigger a cudaFree + cudaMalloc of its own (on this
      // pool's boxes a cudaFree in the middle of a build sporadically took 200-800 ms: 2.7 of 4.3 s of set-up)
        bytes = bytes + bytes / 4 + 4096;
        const cudaError_t e = cudaMalloc(&p.ptr, bytes);
        if (e != cudaSuccess)
        {
            p = Pool{};
            return e;
        }
        p.bytes  = bytes;
        p.device = device;
    }
    *out = p.ptr;
    return cudaSuccess;
}
} // namespace

void beam_builder_release()
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (auto &p : g_pool)
        if (p.ptr)
        {
            int prev = -1;
            cudaGetDevice(&prev);
            cudaSetDevice(p.device);
            cudaFree(p.ptr);
            if (prev >= 0)
                cudaSetDevice(prev);
            p = Pool{};
        }
}

static bool build_once(const Track &t, const BeamConfig &cfg, int device, int items_per_entry, uint8_t **d_blob_out,
                       size_t *bytes_out, std::string &err, bool &full)
{
    full = false;
    std::lock_guard<std::mutex> pool_lock(g_pool_mu); // one build at a time per process: the scratch pool is shared
    const auto t_begin = std::chrono::steady_clock::now();
    BeamPlan   pl;
    if (!beam_plan(t, cfg, pl, err))
        return false;
    const auto    t_plan = std::chrono::steady_clock::now();
    const int32_t ns = t.n_segments(), nb = pl.nb;
    const size_t  n_rows = pl.covered.size();
    int           prev = -1, sms = 0;
    bool          ok = false;
    void         *d_seg = nullptr, *d_cov = nullptr, *d_scratch = nullptr, *d_entries = nullptr, *d_items = nullptr, *d_ctr = nullptr;
    uint8_t      *d_blob = nullptr;
    BeamHeader    hdr{};
    unsigned long long    used = 0, capacity = 0;
    int32_t               counters[4] = {0, 0, 0, 0};
    if (cudaGetDevice(&prev) != cudaSuccess)
    {
        cudaGetLastError();
        err = "no CUDA device";
        return false;
    }
    {
        BEAM_CUDA(cudaSetDevice(device));
        BEAM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int    grid = static_cast<int>(std::min<size_t>(n_rows, static_cast<size_t>(sms) * 4));
        const size_t per_cta = static_cast<size_t>(ns) * (4 + 2 + 2 + 2 + 2) + static_cast<size_t>(nb) * kListCap * (4 + 2);
        // `items_per_entry` rest candidates per (cell, bin) on average; the caller retries with more when the kernel reports the array full
        capacity = static_cast<unsigned long long>(n_rows) * nb * static_cast<unsigned long long>(items_per_entry) + 1024ull;
        GpuBuild g{};
        BEAM_CUDA(pool_get(3, device, sizeof(float4) * ns, &d_seg));
        BEAM_CUDA(cudaMemcpy(d_seg, t.segments.data(), sizeof(float4) * ns, cudaMemcpyHostToDevice));
        BEAM_CUDA(pool_get(4, device, 4 * std::max<size_t>(n_rows, 1), &d_cov));
        BEAM_CUDA(cudaMemcpy(d_cov, pl.covered.data(), 4 * n_rows, cudaMemcpyHostToDevice));
        BEAM_CUDA(pool_get(0, device, per_cta * grid + 256, &d_scratch));
        BEAM_CUDA(pool_get(1, device, 16 * std::max<size_t>(n_rows * nb, 1), &d_entries));
        BEAM_CUDA(pool_get(2, device, 2 * capacity, &d_items));
        BEAM_CUDA(pool_get(5, device, 64, &d_ctr));
        BEAM_CUDA(cudaMemset(d_ctr, 0, 64));
        g.seg = static_cast<const float4 *>(d_seg);
        g.ns  = ns;
        g.x0 = pl.x0, g.y0 = pl.y0, g.h = pl.h, g.rb = pl.rb, g.range = cfg.range, g.pad = cfg.pad, g.dth = kBeamAngleMargin;
        g.nx = pl.nx, g.nb = nb;
        g.covered = static_cast<const uint32_t *>(d_cov);
        g.n_rows  = static_cast<int32_t>(n_rows);
        {
            uint8_t *p = static_cast<uint8_t *>(d_scratch);
            g.cand_lb  = reinterpret_cast<float *>(p);
            p += static_cast<size_t>(grid) * ns * 4;
            g.list_d = reinterpret_cast<float *>(p);
            p += static_cast<size_t>(grid) * nb * kListCap * 4;
            g.cand_seg = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * ns * 2;
            g.cand_b0 = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * ns * 2;
            g.cand_bn = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * ns * 2;
            g.cand_act = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * ns * 2;
            g.list_s = reinterpret_cast<uint16_t *>(p);
        }
        g.entries       = static_cast<uint4 *>(d_entries);
        g.items         = static_cast<uint16_t *>(d_items);
        g.item_cursor   = static_cast<unsigned long long *>(d_ctr);
        g.item_capacity = capacity;
        g.row_cursor    = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(d_ctr) + 8);
        g.overflow      = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(d_ctr) + 16);
        const auto t_k0 = std::chrono::steady_clock::now();
        if (n_rows)
        {
            const size_t smem = static_cast<size_t>(nb) * (32 + 40 + 8 + 32);
            BEAM_CUDA(cudaFuncSetAttribute(beam_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            beam_build_kernel<<<grid, kThreads, smem>>>(g);
            BEAM_CUDA(cudaGetLastError());
            BEAM_CUDA(cudaDeviceSynchronize());
        }
        const auto t_k1 = std::chrono::steady_clock::now();
        BEAM_CUDA(cudaMemcpy(&used, d_ctr, 8, cudaMemcpyDeviceToHost));
        BEAM_CUDA(cudaMemcpy(counters, static_cast<uint8_t *>(d_ctr) + 16, 8, cudaMemcpyDeviceToHost));
        if (counters[1] != 0 || used > capacity)
        {
            err  = "beam table: item array full";
            full = true;
            goto done;
        }
        // the table in its final layout, assembled on the device: header | rows | entries | items
        if (!beam_layout(pl, cfg, used, ns, hdr, err))
            goto done;
        BEAM_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_blob), hdr.bytes));
        // (only the padding behind the rows and behind the items is not overwritten below)
        BEAM_CUDA(cudaMemset(d_blob + hdr.off_entries - 16, 0, 16));
        BEAM_CUDA(cudaMemset(d_blob + hdr.bytes - 16, 0, 16));
        BEAM_CUDA(cudaMemcpy(d_blob, &hdr, sizeof hdr, cudaMemcpyHostToDevice));
        BEAM_CUDA(cudaMemcpy(d_blob + hdr.off_rows, pl.rows.data(), pl.rows.size() * 4, cudaMemcpyHostToDevice));
        if (n_rows)
            BEAM_CUDA(cudaMemcpy(d_blob + hdr.off_entries, d_entries, 16 * n_rows * nb, cudaMemcpyDeviceToDevice));
        if (used)
            BEAM_CUDA(cudaMemcpy(d_blob + hdr.off_items, d_items, 2 * used, cudaMemcpyDeviceToDevice));
        BEAM_CUDA(cudaDeviceSynchronize());
        ok = true;
        if (std::getenv("OK_BEAM_VERBOSE"))
        {
            const auto t_a1 = std::chrono::steady_clock::now();
            auto       ms   = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            std::fprintf(stderr, "[ok_beam_gpu] rows %zu bins %d: plan %.1f ms, setup %.1f ms, kernel %.1f ms, assemble %.1f ms, free %.1f ms, %.1f MB, %d lists overflowed\n",
                         n_rows, nb, ms(t_begin, t_plan), ms(t_plan, t_k0), ms(t_k0, t_k1), ms(t_k1, t_a1), ms(t_a1, std::chrono::steady_clock::now()), hdr.bytes / 1e6, counters[0]);
        }
    }
done: // (every scratch buffer stays in the pool: beam_builder_release)
    if (ok)
    {
        *d_blob_out = d_blob;
        *bytes_out  = hdr.bytes;
    }
    else if (d_blob)
        cudaFree(d_blob);
    if (prev >= 0)
        cudaSetDevice(prev);
    return ok;
}

bool build_beam_table_device(const Track &t, const BeamConfig &cfg, int device, uint8_t **d_blob, size_t *bytes, std::string &err)
{
    for (int per = 16; per <= 1024; per *= 4) // rest candidates per entry the item array is sized for (3-10 on the 23 tracks)
    {
        bool full = false;
        if (build_once(t, cfg, device, per, d_blob, bytes, err, full))
            return true;
        if (!full)
            return false;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// Several tables back to back, pipelined: while the kernel of track i runs, the host sizes, allocates and assembles the
// table of track i - 1 on a second stream (the cudaMalloc of a 120-430 MB table sporadically takes 50-100 ms on this
// pool's boxes; hidden behind the kernel it costs nothing).  Entries and items are double-buffered for that.  A track
// whose item array turns out too small (or anything else that fails) is left to the caller: blobs[i] stays null.
// ---------------------------------------------------------------------------------------------------------------------
void build_beam_tables_device(const std::vector<const Track *> &tracks, const BeamConfig &cfg, int device, std::vector<uint8_t *> &blobs,
                              std::vector<size_t> &bytes)
{
    const size_t n = tracks.size();
    blobs.assign(n, nullptr);
    bytes.assign(n, 0);
    if (!n)
        return;
    std::lock_guard<std::mutex> pool_lock(g_pool_mu);
    int                         prev = -1, sms = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess)
    {
        cudaGetLastError();
        return;
    }
    struct Job
    {
        BeamPlan            pl;
        int32_t             ns{0};
        size_t              n_rows{0};
        unsigned long long  capacity{0};
        void               *d_entries{nullptr}, *d_items{nullptr};
        bool                launched{false};
    };
    std::vector<Job>    jobs(n);
    cudaStream_t        sk = nullptr, sc = nullptr;
    cudaEvent_t         ev_kernel[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    unsigned long long *h_ctr = nullptr; // pinned: 2 x 8 words {items used, -, overflow[0..1]}
    std::string         err;
    const auto          t_begin = std::chrono::steady_clock::now();
    bool                ready = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess &&
                 cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking) == cudaSuccess &&
                 cudaMallocHost(reinterpret_cast<void **>(&h_ctr), 2 * 64) == cudaSuccess;
    for (int k = 0; ready && k < 2; ++k)
        ready = cudaEventCreateWithFlags(&ev_kernel[k], cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&ev_copy[k], cudaEventDisableTiming) == cudaSuccess;
    bool copy_pending[2] = {false, false};

    // stage 1 of track i: plan, upload, kernel, counters back -- all on stream sk, nothing waits
    auto launch = [&](size_t i) -> bool {
        Job &j = jobs[i];
        const Track &t = *tracks[i];
        const int    set = static_cast<int>(i & 1);
        if (!beam_plan(t, cfg, j.pl, err))
            return false;
        j.ns = t.n_segments(), j.n_rows = j.pl.covered.size();
        const int32_t nb = j.pl.nb;
        if (!j.n_rows)
            return false;
        const int    grid    = static_cast<int>(std::min<size_t>(j.n_rows, static_cast<size_t>(sms) * 4));
        const size_t per_cta = static_cast<size_t>(j.ns) * (4 + 2 + 2 + 2 + 2) + static_cast<size_t>(nb) * kListCap * (4 + 2);
        j.capacity           = static_cast<unsigned long long>(j.n_rows) * nb * 16ull + 1024ull;
        void *d_seg = nullptr, *d_cov = nullptr, *d_scratch = nullptr, *d_ctr = nullptr;
        if (copy_pending[set])
        { // the copies of track i - 2 read this set's entries and items
            if (cudaEventSynchronize(ev_copy[set]) != cudaSuccess)
                return false;
            copy_pending[set] = false;
        }
        // (the shared scratch, segments, cells and counters are only touched by work on sk, which is ordered)
        if (pool_get(0, device, per_cta * grid + 256, &d_scratch) != cudaSuccess || pool_get(3, device, sizeof(float4) * j.ns, &d_seg) != cudaSuccess ||
            pool_get(4, device, 4 * j.n_rows, &d_cov) != cudaSuccess || pool_get(5, device, 64, &d_ctr) != cudaSuccess ||
            pool_get(set ? 6 : 1, device, 16 * j.n_rows * nb, &j.d_entries) != cudaSuccess || pool_get(set ? 7 : 2, device, 2 * j.capacity, &j.d_items) != cudaSuccess)
            return false;
        GpuBuild g{};
        if (cudaMemcpyAsync(d_seg, t.segments.data(), sizeof(float4) * j.ns, cudaMemcpyHostToDevice, sk) != cudaSuccess ||
            cudaMemcpyAsync(d_cov, j.pl.covered.data(), 4 * j.n_rows, cudaMemcpyHostToDevice, sk) != cudaSuccess || cudaMemsetAsync(d_ctr, 0, 64, sk) != cudaSuccess)
            return false;
        g.seg = static_cast<const float4 *>(d_seg);
        g.ns  = j.ns;
        g.x0 = j.pl.x0, g.y0 = j.pl.y0, g.h = j.pl.h, g.rb = j.pl.rb, g.range = cfg.range, g.pad = cfg.pad, g.dth = kBeamAngleMargin;
        g.nx = j.pl.nx, g.nb = nb;
        g.covered = static_cast<const uint32_t *>(d_cov);
        g.n_rows  = static_cast<int32_t>(j.n_rows);
        {
            uint8_t *p = static_cast<uint8_t *>(d_scratch);
            g.cand_lb  = reinterpret_cast<float *>(p);
            p += static_cast<size_t>(grid) * j.ns * 4;
            g.list_d = reinterpret_cast<float *>(p);
            p += static_cast<size_t>(grid) * nb * kListCap * 4;
            g.cand_seg = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * j.ns * 2;
            g.cand_b0 = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * j.ns * 2;
            g.cand_bn = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * j.ns * 2;
            g.cand_act = reinterpret_cast<uint16_t *>(p);
            p += static_cast<size_t>(grid) * j.ns * 2;
            g.list_s = reinterpret_cast<uint16_t *>(p);
        }
        g.entries       = static_cast<uint4 *>(j.d_entries);
        g.items         = static_cast<uint16_t *>(j.d_items);
        g.item_cursor   = static_cast<unsigned long long *>(d_ctr);
        g.item_capacity = j.capacity;
        g.row_cursor    = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(d_ctr) + 8);
        g.overflow      = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(d_ctr) + 16);
        const size_t smem = static_cast<size_t>(nb) * (32 + 40 + 8 + 32);
        if (cudaFuncSetAttribute(beam_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return false;
        beam_build_kernel<<<grid, kThreads, smem, sk>>>(g);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(h_ctr + 8 * set, d_ctr, 64, cudaMemcpyDeviceToHost, sk) != cudaSuccess ||
            cudaEventRecord(ev_kernel[set], sk) != cudaSuccess)
            return false;
        j.launched = true;
        return true;
    };
    // stage 2 of track i (while the kernel of track i + 1 runs): size, allocate, assemble on stream sc
    auto finish = [&](size_t i) {
        Job &j = jobs[i];
        if (!j.launched)
            return;
        const int set = static_cast<int>(i & 1);
        if (cudaEventSynchronize(ev_kernel[set]) != cudaSuccess)
            return;
        const unsigned long long used     = h_ctr[8 * set];
        const int32_t           *counters = reinterpret_cast<const int32_t *>(h_ctr + 8 * set + 2);
        if (counters[1] != 0 || used > j.capacity)
            return; // item array full: the caller's one-track build retries with a larger one
        BeamHeader hdr{};
        if (!beam_layout(j.pl, cfg, used, j.ns, hdr, err))
            return;
        uint8_t *d_blob = nullptr;
        if (cudaMalloc(reinterpret_cast<void **>(&d_blob), hdr.bytes) != cudaSuccess)
        {
            cudaGetLastError();
            return;
        }
        // the header and the rows go through a staging copy the driver makes at the call (pageable source): safe to let go of
        bool ok = cudaMemsetAsync(d_blob + hdr.off_entries - 16, 0, 16, sc) == cudaSuccess && cudaMemsetAsync(d_blob + hdr.bytes - 16, 0, 16, sc) == cudaSuccess &&
                  cudaMemcpyAsync(d_blob, &hdr, sizeof hdr, cudaMemcpyHostToDevice, sc) == cudaSuccess &&
                  cudaMemcpyAsync(d_blob + hdr.off_rows, j.pl.rows.data(), j.pl.rows.size() * 4, cudaMemcpyHostToDevice, sc) == cudaSuccess &&
                  cudaMemcpyAsync(d_blob + hdr.off_entries, j.d_entries, 16 * j.n_rows * j.pl.nb, cudaMemcpyDeviceToDevice, sc) == cudaSuccess &&
                  (!used || cudaMemcpyAsync(d_blob + hdr.off_items, j.d_items, 2 * used, cudaMemcpyDeviceToDevice, sc) == cudaSuccess) &&
                  cudaEventRecord(ev_copy[set], sc) == cudaSuccess;
        if (!ok)
        {
            cudaGetLastError();
            cudaStreamSynchronize(sc);
            cudaFree(d_blob);
            return;
        }
        copy_pending[set] = true;
        blobs[i] = d_blob, bytes[i] = hdr.bytes;
    };
    if (ready)
    {
        for (size_t i = 0; i < n; ++i)
        {
            launch(i);
            if (i > 0)
                finish(i - 1);
        }
        finish(n - 1);
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(sk);
    }
    if (std::getenv("OK_BEAM_VERBOSE"))
    {
        size_t built = 0, total = 0;
        for (size_t i = 0; i < n; ++i)
            built += blobs[i] != nullptr, total += bytes[i];
        std::fprintf(stderr, "[ok_beam_gpu] pipelined build: %zu of %zu tables, %.1f MB, %.1f ms\n", built, n, total / 1e6,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    }
    for (int k = 0; k < 2; ++k)
    {
        if (ev_kernel[k])
            cudaEventDestroy(ev_kernel[k]);
        if (ev_copy[k])
            cudaEventDestroy(ev_copy[k]);
    }
    if (sk)
        cudaStreamDestroy(sk);
    if (sc)
        cudaStreamDestroy(sc);
    if (h_ctr)
        cudaFreeHost(h_ctr);
    cudaGetLastError();
    if (prev >= 0)
        cudaSetDevice(prev);
}

} // namespace ok
