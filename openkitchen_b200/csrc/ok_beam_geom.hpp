// ok_beam_geom.hpp -- the binary64 geometry of the beam-table builders (host: ok_beam.cpp, device: ok_beam_gpu.cu).
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define OK_HD __host__ __device__ inline
#else
#define OK_HD inline
#endif

namespace ok
{
namespace beamgeom
{
struct V2
{
    double x, y;
};
OK_HD double cross(const V2 &a, const V2 &b)
{
    return a.x * b.y - a.y * b.x;
}

// convex hull (counter-clockwise, no collinear points) of at most 8 points; returns the vertex count
OK_HD int convex_hull8(V2 *pts, int n, V2 *out)
{
    for (int i = 1; i < n; ++i)
    { // insertion sort by (x, y): n <= 8
        const V2 v = pts[i];
        int      j = i - 1;
        while (j >= 0 && (pts[j].x > v.x || (pts[j].x == v.x && pts[j].y > v.y)))
        {
            pts[j + 1] = pts[j];
            --j;
        }
        pts[j + 1] = v;
    }
    int k = 0;
    for (int i = 0; i < n; ++i)
    {
        while (k >= 2 && cross({out[k - 1].x - out[k - 2].x, out[k - 1].y - out[k - 2].y},
                               {pts[i].x - out[k - 2].x, pts[i].y - out[k - 2].y}) <= 0)
            --k;
        out[k++] = pts[i];
    }
    for (int i = n - 2, lo = k + 1; i >= 0; --i)
    {
        while (k >= lo && cross({out[k - 1].x - out[k - 2].x, out[k - 1].y - out[k - 2].y},
                                {pts[i].x - out[k - 2].x, pts[i].y - out[k - 2].y}) <= 0)
            --k;
        out[k++] = pts[i];
    }
    return k > 1 ? k - 1 : k;
}

// keeps the part of the convex polygon with cross(n, v) >= 0
OK_HD int clip_half_plane(const V2 *poly, int m, const V2 &n, V2 *out)
{
    int k = 0;
    for (int i = 0; i < m; ++i)
    {
        const V2     a = poly[i], b = poly[(i + 1) % m];
        const double fa = cross(n, a), fb = cross(n, b);
        if (fa >= 0)
            out[k++] = a;
        if ((fa >= 0) != (fb >= 0))
        {
            const double t = fa / (fa - fb);
            out[k++]       = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)};
        }
    }
    return k;
}

// distance from the origin to a convex polygon (counter-clockwise), 0 when the origin is inside
OK_HD double origin_distance(const V2 *poly, int m)
{
    if (m == 0)
        return 1e300;
    if (m == 1)
        return hypot(poly[0].x, poly[0].y);
    bool   inside = m >= 3;
    double best   = 1e300;
    for (int i = 0; i < m; ++i)
    {
        const V2 a = poly[i], b = poly[(i + 1) % m];
        const V2 e = {b.x - a.x, b.y - a.y};
        if (cross(e, {-a.x, -a.y}) < 0)
            inside = false;
        const double l2 = e.x * e.x + e.y * e.y;
        double       t  = l2 > 0 ? (-(a.x * e.x) - a.y * e.y) / l2 : 0.0;
        t               = fmin(1.0, fmax(0.0, t));
        best            = fmin(best, hypot(a.x + t * e.x, a.y + t * e.y));
    }
    return inside ? 0.0 : best;
}

OK_HD double point_segment_distance(double px, double py, double ax, double ay, double bx, double by)
{
    const double ex = bx - ax, ey = by - ay, l2 = ex * ex + ey * ey;
    double       t  = l2 > 0 ? ((px - ax) * ex + (py - ay) * ey) / l2 : 0.0;
    t               = fmin(1.0, fmax(0.0, t));
    return hypot(ax + t * ex - px, ay + t * ey - py);
}

} // namespace beamgeom
} // namespace ok
