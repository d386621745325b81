// ok_track.cpp -- see ok_track.hpp.  Host code only (compiled without -ffast-math, -ffp-contract=off).
#include "ok_track.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace ok
{
namespace
{
constexpr float kWindowW = 1600.f; // Typedefs.h:7
constexpr float kWindowH = 1400.f; // Typedefs.h:8

struct Box
{
    float lo_x, lo_y, hi_x, hi_y;
};

// RaceTrack::calculateTrackExtents, RaceTrack.cpp:166-196
Box bounds_of(const std::vector<float> &xs, const std::vector<float> &ys)
{
    Box b{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(),
          -std::numeric_limits<float>::max()};
    for (float v : xs)
    {
        if (v < b.lo_x)
            b.lo_x = v;
        if (v > b.hi_x)
            b.hi_x = v;
    }
    for (float v : ys)
    {
        if (v < b.lo_y)
            b.lo_y = v;
        if (v > b.hi_y)
            b.hi_y = v;
    }
    return b;
}

// RaceTrack::gradient, RaceTrack.cpp:87-114: one-sided at the ends, central elsewhere
std::vector<float> finite_difference(const std::vector<float> &v)
{
    const size_t       n = v.size();
    std::vector<float> g(n);
    g[0]     = v[1] - v[0];
    g[n - 1] = v[n - 1] - v[n - 2];
    for (size_t i = 1; i + 1 < n; ++i)
        g[i] = (v[i + 1] - v[i - 1]) / 2.0f;
    return g;
}

// width rule of RaceTrack.cpp:138-160 with std::max / std::min comparison order
inline float lane_width(float raw)
{
    float w = (4.0f < raw) ? raw : 4.0f; // std::max(4, raw)
    w       = w * 3.0f;
    return (17.0f < w) ? 17.0f : w;      // std::min(w, 17)
}

template <typename T> size_t append_section(std::vector<uint8_t> &blob, const T *data, size_t count)
{
    const size_t off = blob.size();
    blob.resize(off + ((count * sizeof(T) + 15) / 16) * 16, 0);
    if (count)
        std::memcpy(blob.data() + off, data, count * sizeof(T));
    return off;
}

// Does segment p->q come within `m` of the axis-aligned box?  Conservative (never false when
// it does): slab-clips the segment against the box grown by m, in double.
bool segment_touches_box(double px, double py, double qx, double qy, double bx0, double by0, double bx1, double by1, double m)
{
    bx0 -= m, by0 -= m, bx1 += m, by1 += m;
    double       t0 = 0.0, t1 = 1.0;
    const double d[2]  = {qx - px, qy - py};
    const double o[2]  = {px, py};
    const double lo[2] = {bx0, by0};
    const double hi[2] = {bx1, by1};
    for (int a = 0; a < 2; ++a)
    {
        if (d[a] == 0.0)
        {
            if (o[a] < lo[a] || o[a] > hi[a])
                return false;
            continue;
        }
        double ta = (lo[a] - o[a]) / d[a], tb = (hi[a] - o[a]) / d[a];
        if (ta > tb)
            std::swap(ta, tb);
        // widen by a few ulps so that rounding in the divisions can only add cells
        ta -= 1e-12, tb += 1e-12;
        t0 = std::max(t0, ta);
        t1 = std::min(t1, tb);
        if (t0 > t1)
            return false;
    }
    return true;
}

bool build_grid(Track &t, float cell, std::string &err)
{
    const int32_t ns = t.n_segments();
    double        lo_x = 1e300, lo_y = 1e300, hi_x = -1e300, hi_y = -1e300;
    for (int32_t s = 0; s < ns; ++s)
        for (int e = 0; e < 2; ++e)
        {
            const double px = t.segments[4 * s + 2 * e], py = t.segments[4 * s + 2 * e + 1];
            if (!std::isfinite(px) || !std::isfinite(py))
            {
                err = "track geometry is not finite";
                return false;
            }
            lo_x = std::min(lo_x, px), hi_x = std::max(hi_x, px);
            lo_y = std::min(lo_y, py), hi_y = std::max(hi_y, py);
        }
    t.cell    = cell;
    // the segments' box grown by 1 px is the device's clip box; one more ring of (empty) cells around
    // it lets the DDA run without bounds checks
    const float in_x0 = std::floor(static_cast<float>(lo_x) - 1.0f);
    const float in_y0 = std::floor(static_cast<float>(lo_y) - 1.0f);
    const int32_t in_nx = std::max<int32_t>(1, static_cast<int32_t>(std::ceil((hi_x + 1.0 - in_x0) / cell)));
    const int32_t in_ny = std::max<int32_t>(1, static_cast<int32_t>(std::ceil((hi_y + 1.0 - in_y0) / cell)));
    t.grid_x0 = in_x0 - cell;
    t.grid_y0 = in_y0 - cell;
    t.grid_nx = in_nx + 2;
    t.grid_ny = in_ny + 2;
    const int64_t n_cells = static_cast<int64_t>(t.grid_nx) * t.grid_ny;
    if (n_cells > (1 << 22))
    {
        err = "grid has too many cells";
        return false;
    }
    std::vector<std::vector<uint16_t>> lists(static_cast<size_t>(n_cells));
    const double                       m = kGridMargin;
    for (int32_t s = 0; s < ns; ++s)
    {
        const double px = t.segments[4 * s], py = t.segments[4 * s + 1];
        const double qx = t.segments[4 * s + 2], qy = t.segments[4 * s + 3];
        auto cell_of = [&](double v, double origin, int32_t n) {
            int64_t c = static_cast<int64_t>(std::floor((v - origin) / cell));
            return static_cast<int32_t>(std::min<int64_t>(std::max<int64_t>(c, 0), n - 1));
        };
        const int32_t ix0 = cell_of(std::min(px, qx) - m, t.grid_x0, t.grid_nx);
        const int32_t ix1 = cell_of(std::max(px, qx) + m, t.grid_x0, t.grid_nx);
        const int32_t iy0 = cell_of(std::min(py, qy) - m, t.grid_y0, t.grid_ny);
        const int32_t iy1 = cell_of(std::max(py, qy) + m, t.grid_y0, t.grid_ny);
        for (int32_t iy = iy0; iy <= iy1; ++iy)
            for (int32_t ix = ix0; ix <= ix1; ++ix)
            {
                const double bx0 = t.grid_x0 + static_cast<double>(ix) * cell;
                const double by0 = t.grid_y0 + static_cast<double>(iy) * cell;
                if (segment_touches_box(px, py, qx, qy, bx0, by0, bx0 + cell, by0 + cell, m))
                    lists[static_cast<size_t>(iy) * t.grid_nx + ix].push_back(static_cast<uint16_t>(s));
            }
    }
    size_t total = 0, occupied = 0;
    for (auto &l : lists)
    {
        total += l.size();
        occupied += l.empty() ? 0 : 1;
    }
    if (total > 65535 || occupied > 65535)
    {
        err = "too many (cell, segment) registrations for 16-bit offsets";
        return false;
    }
    // Sparse cell table: one occupancy bit per cell, 32 cells per word, each word paired with the
    // number of occupied cells before it; item offsets exist only for occupied cells.  An empty
    // cell -- most of what a ray crosses -- costs one shared-memory word and a bit test.
    const int64_t n_words = (n_cells + 31) / 32;
    t.cell_words.assign(static_cast<size_t>(n_words) * 2, 0);
    t.cell_starts.clear();
    t.items.clear();
    t.items.reserve(total);
    uint32_t rank = 0;
    for (int64_t c = 0; c < n_cells; ++c)
    {
        if ((c & 31) == 0)
            t.cell_words[2 * (c >> 5) + 1] = rank;
        if (lists[c].empty())
            continue;
        t.cell_words[2 * (c >> 5)] |= 1u << (c & 31);
        const size_t first_item = t.items.size();
        // inner boundaries first: from inside the lane they are hit before the outer ones 3 px behind
        // them, which then fail the cheap "beyond the current hit" screen (any order is exact)
        std::stable_sort(lists[c].begin(), lists[c].end(), [&](uint16_t p, uint16_t q) {
            auto outer = [&](uint16_t s) {
                const int32_t chain = t.n_points() - 1;
                if (s < 4 * chain)
                    return (s / chain) & 1; // chains LI, LO, RI, RO
                return (s - 4 * chain) >= 2 ? 1 : 0; // closures LI, RI, LO, RO
            };
            return outer(p) < outer(q);
        });
        t.items.insert(t.items.end(), lists[c].begin(), lists[c].end());
        t.cell_starts.push_back(static_cast<uint32_t>(first_item) | (static_cast<uint32_t>(t.items.size()) << 16));
        ++rank;
    }
    return true;
}

void pack_blob(Track &t)
{
    const int32_t      n = t.n_points(), ns = t.n_segments();
    // one extra, all-zero "null segment" at index ns: zero length makes denom = 0, which fails the parallel test
    // of CollisionChecker.cu:23 for every ray -- the beam lists pad their chunks with it
    std::vector<float> seg4(static_cast<size_t>(ns + 1) * 4, 0.0f), pts(static_cast<size_t>(n) * 2), widths(n);
    for (int32_t s = 0; s < ns; ++s)
    {
        const float *g  = &t.segments[4 * s];
        seg4[4 * s]     = g[0];
        seg4[4 * s + 1] = g[1];
        seg4[4 * s + 2] = g[2] - g[0]; // seg_dx of CollisionChecker.cu:19, same single rounding
        seg4[4 * s + 3] = g[3] - g[1]; // seg_dy
    }
    for (int32_t i = 0; i < n; ++i)
    {
        pts[2 * i]     = t.x[i];
        pts[2 * i + 1] = t.y[i];
        widths[i]      = t.w_left[i] + t.w_right[i]; // lane_width of RaceTrack.cpp:67-69
    }
    t.blob.assign(sizeof(TrackHeader), 0);
    TrackHeader h{};
    h.n_points = n, h.n_segments = ns, h.grid_nx = t.grid_nx, h.grid_ny = t.grid_ny;
    h.grid_x0 = t.grid_x0, h.grid_y0 = t.grid_y0, h.cell = t.cell, h.inv_cell = 1.0f / t.cell;
    h.off_segments = static_cast<uint32_t>(append_section(t.blob, seg4.data(), seg4.size()));
    h.off_words    = static_cast<uint32_t>(append_section(t.blob, t.cell_words.data(), t.cell_words.size()));
    h.off_starts   = static_cast<uint32_t>(append_section(t.blob, t.cell_starts.data(), t.cell_starts.size()));
    h.off_items    = static_cast<uint32_t>(append_section(t.blob, t.items.data(), t.items.size()));
    h.off_points   = static_cast<uint32_t>(append_section(t.blob, pts.data(), pts.size()));
    h.off_widths   = static_cast<uint32_t>(append_section(t.blob, widths.data(), widths.size()));
    h.off_headings = static_cast<uint32_t>(append_section(t.blob, t.heading.data(), t.heading.size()));
    // safe radius of the windowed nearest-point search (ok_track.hpp, kNearestWindow)
    t.safe_radius.assign(n, std::numeric_limits<float>::infinity());
    if (n > kNearestWindow)
        for (int32_t i = 0; i < n; ++i)
        {
            double best = 1e300;
            for (int32_t j = 0; j < n; ++j)
            {
                const int32_t fwd = ((j - i) % n + n) % n; // cyclic offset of j from i
                if (fwd < kNearestWindow / 2 || fwd >= n - kNearestWindow / 2)
                    continue; // inside the window [i - 16, i + 16)
                best = std::min(best, std::hypot(static_cast<double>(t.x[j]) - t.x[i], static_cast<double>(t.y[j]) - t.y[i]));
            }
            t.safe_radius[i] = static_cast<float>(best * (1.0 - 1e-6));
        }
    h.off_safe   = static_cast<uint32_t>(append_section(t.blob, t.safe_radius.data(), t.safe_radius.size()));
    h.blob_bytes = static_cast<uint32_t>(t.blob.size());
    std::memcpy(t.blob.data(), &h, sizeof h);
}
} // namespace

bool read_track_csv(const std::string &path, std::vector<float> cols[4], std::string &err)
{
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f)
    {
        err = "cannot open track csv: " + path;
        return false;
    }
    for (int c = 0; c < 4; ++c)
        cols[c].clear();
    char line[1024];
    bool header = true;
    while (std::fgets(line, sizeof line, f))
    {
        if (header)
        { // RaceTrack.cpp:136-137 discards the first line unconditionally
            header = false;
            continue;
        }
        const char *p = line;
        float       v[4];
        int         got = 0;
        for (; got < 4; ++got)
        {
            char *end = nullptr;
            v[got]    = std::strtof(p, &end); // std::stof
            if (end == p)
                break;
            p = std::strchr(end, ',');
            if (!p)
            {
                ++got;
                break;
            }
            ++p;
        }
        if (got < 4)
        {
            // the reference would throw out of std::stof here; report instead of continuing to UB
            bool blank = true;
            for (const char *q = line; *q; ++q)
                if (!std::isspace(static_cast<unsigned char>(*q)))
                    blank = false;
            if (blank)
                continue;
            std::fclose(f);
            err = "malformed row in track csv: " + path;
            return false;
        }
        for (int c = 0; c < 4; ++c)
            cols[c].push_back(v[c]);
    }
    std::fclose(f);
    if (cols[0].size() < 2)
    {
        err = "track csv has fewer than 2 points: " + path;
        return false;
    }
    return true;
}

bool build_track(const float *x_m, const float *y_m, const float *w_right, const float *w_left, int32_t n, float cell,
                 size_t max_blob_bytes, Track &t, std::string &err)
{
    if (n < 2 || !x_m || !y_m || !w_right || !w_left)
    {
        err = "a track needs at least 2 centre-line points";
        return false;
    }
    if (4 * (static_cast<int64_t>(n) - 1) + 4 > 65535)
    {
        err = "track has too many points for 16-bit segment indices";
        return false;
    }
    t = Track{};
    t.x.assign(x_m, x_m + n);
    t.y.assign(y_m, y_m + n);
    t.w_right.resize(n);
    t.w_left.resize(n);
    for (int32_t i = 0; i < n; ++i)
    {
        t.w_right[i] = lane_width(w_right[i]);
        t.w_left[i]  = lane_width(w_left[i]);
    }

    // fit to the 1600x1400 window with 10 % padding and recentre: RaceTrack.cpp:198-229
    {
        const Box   b  = bounds_of(t.x, t.y);
        const float sx = kWindowW / (b.hi_x - b.lo_x);
        const float sy = kWindowH / (b.hi_y - b.lo_y);
        float       k  = (sy < sx) ? sy : sx; // std::min(sx, sy)
        k *= static_cast<float>(0.9);
        for (int32_t i = 0; i < n; ++i)
        {
            t.x[i] *= k;
            t.y[i] *= k;
            t.w_left[i] *= k;
            t.w_right[i] *= k;
        }
        const Box   s  = bounds_of(t.x, t.y);
        const float mx = (kWindowW / 2.f) - ((s.hi_x + s.lo_x) / 2.f);
        const float my = (kWindowH / 2.f) - ((s.hi_y + s.lo_y) / 2.f);
        for (int32_t i = 0; i < n; ++i)
        {
            t.x[i] += mx;
            t.y[i] += my;
        }
    }

    // unit tangents, headings and the four offset polylines: RaceTrack.cpp:257-307
    {
        std::vector<float> tx = finite_difference(t.x), ty = finite_difference(t.y);
        t.heading.resize(n);
        for (int32_t i = 0; i < n; ++i)
        {
            const float len = std::sqrt(tx[i] * tx[i] + ty[i] * ty[i]);
            tx[i] /= len;
            ty[i] /= len;
            // float atan2, float * 180, then a DOUBLE division by M_PI narrowed back (RaceTrack.cpp:278)
            t.heading[i] = static_cast<float>(static_cast<double>(std::atan2(ty[i], tx[i]) * 180.0f) / M_PI);
        }
        constexpr float kEdge = 3.f; // kBoundaryThickness
        t.left_inner.resize(2 * n), t.left_outer.resize(2 * n), t.right_inner.resize(2 * n), t.right_outer.resize(2 * n);
        for (int32_t i = 0; i < n; ++i)
        {
            const float cx = t.x[i], cy = t.y[i], wr = t.w_right[i], wl = t.w_left[i];
            t.right_inner[2 * i]     = cx + wr * ty[i];
            t.right_inner[2 * i + 1] = cy - wr * tx[i];
            t.left_inner[2 * i]      = cx - wl * ty[i];
            t.left_inner[2 * i + 1]  = cy + wl * tx[i];
            t.right_outer[2 * i]     = cx + (wr + kEdge) * ty[i];
            t.right_outer[2 * i + 1] = cy - (wr + kEdge) * tx[i];
            t.left_outer[2 * i]      = cx - (wl + kEdge) * ty[i];
            t.left_outer[2 * i + 1]  = cy + (wl + kEdge) * tx[i];
        }
    }

    // segment list in TrackSegments order (TrackSegments.cu:11-38): LI, LO, RI, RO chains, then the
    // loop closures LI, RI, LO, RO (last point -> first point)
    {
        const std::vector<float> *chains[4]  = {&t.left_inner, &t.left_outer, &t.right_inner, &t.right_outer};
        const std::vector<float> *closing[4] = {&t.left_inner, &t.right_inner, &t.left_outer, &t.right_outer};
        t.segments.reserve(static_cast<size_t>(4 * (n - 1) + 4) * 4);
        for (auto *c : chains)
            for (int32_t i = 0; i + 1 < n; ++i)
                t.segments.insert(t.segments.end(), {(*c)[2 * i], (*c)[2 * i + 1], (*c)[2 * i + 2], (*c)[2 * i + 3]});
        for (auto *c : closing)
            t.segments.insert(t.segments.end(), {(*c)[2 * (n - 1)], (*c)[2 * (n - 1) + 1], (*c)[0], (*c)[1]});
    }

    // broadphase: grow the cell until the staged blob fits
    for (float c = cell;; c *= 2.f)
    {
        std::string gerr;
        if (build_grid(t, c, gerr))
        {
            pack_blob(t);
            if (t.blob.size() <= max_blob_bytes)
                return true;
            gerr = "track blob does not fit shared memory";
        }
        if (c > 512.f)
        {
            err = gerr;
            return false;
        }
    }
}

} // namespace ok
