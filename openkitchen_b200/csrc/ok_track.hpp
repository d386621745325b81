// ok_track.hpp -- host-side track builder: CSV -> centre line -> boundary polylines -> segments
// -> uniform-grid broadphase -> one contiguous "track blob" that a CTA stages into shared memory
// with a single bulk copy.
//
// Replaces (reference file:line): RaceTrack ctor chain RaceTrack.cpp:3-14,127-164,166-229,257-307
// and TrackSegments ctor TrackSegments.cu:6-42,53-67.  Runs once per track on the host; the
// arithmetic is binary32 with the same operation order as the reference so the arrays are
// bit-identical to RaceTrack's public vectors (tests/test_tracks.py).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace ok
{

// Header at the start of every track blob (device + shared memory).  Offsets are in bytes from
// the start of the blob; every section is 16-byte aligned.
struct TrackHeader
{
    int32_t  n_points;
    int32_t  n_segments;
    int32_t  grid_nx;
    int32_t  grid_ny;
    float    grid_x0;
    float    grid_y0;
    float    cell;
    float    inv_cell;
    uint32_t off_segments; // float4 {x1, y1, x2-x1, y2-y1}[n_segments], TrackSegments order
    uint32_t off_words;    // uint2 {occupancy bits of 32 cells, #occupied cells before this word}[ceil(nx*ny/32)]
    uint32_t off_starts;   // uint32 per OCCUPIED cell, in cell order: first item | (one past the last) << 16
    uint32_t off_items;    // uint16 segment index [n_items]
    uint32_t off_points;   // float2 centre line [n_points]
    uint32_t off_widths;   // float  w_left + w_right [n_points]
    uint32_t off_headings; // float  heading degrees [n_points]
    uint32_t blob_bytes;   // multiple of 16
    uint32_t off_safe;     // float  [n_points]: see kNearestWindow
    uint32_t pad[3];
};
static_assert(sizeof(TrackHeader) == 80, "TrackHeader must be 80 bytes");

// Nearest-centre-line-point search (RaceTrack.cpp:16-31) with a hint h: the kernel looks at the cyclic window
// [h - kNearestWindow/2, h + kNearestWindow/2) first.  safe[h] = min distance from point h to any point OUTSIDE
// that window (rounded down): if safe[h] - |q - p_h| exceeds the best distance found in the window (plus
// slack), no outside point can be nearer -- |q - p_j| >= |p_j - p_h| - |q - p_h| -- and the windowed argmin
// is the global one; otherwise the kernel falls back to the full search.
constexpr int kNearestWindow = 32;

struct Track
{
    // RaceTrack public state (RaceTrack.h:72-79), same layout as the reference's vectors
    std::vector<float> x, y, w_right, w_left, heading;
    std::vector<float> left_inner, left_outer, right_inner, right_outer; // interleaved xy
    std::vector<float> segments;                                         // x1,y1,x2,y2 per segment
    // broadphase
    int32_t               grid_nx{0}, grid_ny{0};
    float                 grid_x0{0}, grid_y0{0}, cell{8.f};
    std::vector<uint32_t> cell_words; // pairs {occupancy bits, occupied-cell rank} per 32 cells
    std::vector<uint32_t> cell_starts; // per occupied cell: first item | (one past the last) << 16
    std::vector<uint16_t> items;
    std::vector<float>    safe_radius; // per centre-line point, see kNearestWindow
    // staged form
    std::vector<uint8_t> blob;

    int32_t n_points() const
    {
        return static_cast<int32_t>(x.size());
    }
    int32_t n_segments() const
    {
        return static_cast<int32_t>(segments.size() / 4);
    }
};

// margin (px) by which segments are inflated when registered into grid cells; covers every
// rounding error of the device-side traversal (DESIGN.md "broadphase exactness")
constexpr double kGridMargin = 0.0625;

// getTrackDataFromCsv, RaceTrack.cpp:127-164.  Fills the four raw columns; false on I/O error.
bool read_track_csv(const std::string &path, std::vector<float> cols[4], std::string &err);

// Full pipeline on the raw columns.  max_blob_bytes bounds the staged size (shared memory).
bool build_track(const float *x_m, const float *y_m, const float *w_right, const float *w_left, int32_t n, float cell,
                 size_t max_blob_bytes, Track &out, std::string &err);

} // namespace ok
