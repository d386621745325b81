// ok_beam.hpp -- "beam lists": the narrow phase's candidate table.
//
// For every cell of a fine uniform grid over the lane (`cell` px) and every direction bin (2*pi / `bins`)
// the table holds the segments a ray that STARTS in that cell with a direction in that bin can reach within
// a completeness distance d, nearest first:
//
//     list(C, b) = { S : dist(C, b, S) <= d(C, b) },
//     dist(C, b, S) = min { t >= 0 : p + t u(theta) in S, p in C, theta in b }.
//
// dist is computed exactly (up to binary64 rounding, which the inflation margins dominate): with
// H = S (+) (-C) (Minkowski sum, a convex hexagon), dist = Euclidean distance from the origin to
// H intersected with the bin's cone.  d(C, b) is chosen from a few sample rays as "a little beyond the
// farthest first hit"; it only decides how long the list is, never correctness: the kernel accepts
// a list-based result only if the hit parameter is <= d - slack, and otherwise continues with the
// uniform-grid walk from t = d - 1 (ok_kernels.cuh, "beam phase").  Hence
//     every segment whose true crossing parameter is <= d is in the list      (by construction)
//  => a ray whose nearest listed hit is at t <= d - slack has no nearer hit   (rounding << slack)
// which is the same exactness class as the grid walk (DESIGN.md "broadphase exactness").
//
// The reference has no counterpart: CollisionChecker.cu:51-66 tests every ray against every segment.
#pragma once

#include "ok_track.hpp"

#include <cstdint>
#include <string>
#include <vector>

namespace ok
{

// Header of a track's beam blob (device global memory).  Offsets in bytes from the blob start.
struct BeamHeader
{
    float    x0, y0;     // origin of the beam-cell grid
    float    inv_h, h;   // 1 / cell, cell
    int32_t  nx, ny;     // cells
    int32_t  nb;         // direction bins, a power of two
    float    bin_scale;  // nb / (2 pi): bin = floor(angle * bin_scale) & (nb - 1)
    float    rb;         // completeness distance of entries flagged complete (dq = 0xffff)
    uint32_t off_rows;   // uint32[nx * ny]: row of the cell in `entries`, 0xffffffff = not covered
    uint32_t off_entries; // uint2[n_rows * nb]: {first chunk, count | dq << 16}; dq = floor(d * 256), 0xffff = rb
    uint32_t off_items;  // uint16[]: segment indices in chunks of 4, padded with n_segments (the blob's null segment)
    uint32_t n_rows;
    uint32_t n_chunks;
    uint32_t bytes;
    uint32_t pad;
};
static_assert(sizeof(BeamHeader) == 64, "BeamHeader must be 64 bytes");

struct BeamConfig
{
    float   cell{8.f};
    int32_t bins{64};
    float   range{200.f}; // Agent::kSensorRange; lists flagged complete cover range + 1
    float   pad{1.0f};    // px added to the farthest sampled first hit to get the completeness distance
    int32_t threads{0};   // 0 = std::thread::hardware_concurrency()
};

// bump when build_beam_table changes what it produces (the on-disk cache is keyed on it)
constexpr int    kBeamBuildVersion = 1;
constexpr double kBeamCellMargin  = 0.0625; // px: the cell is grown by this much on every side
constexpr double kBeamAngleMargin = 1e-4;   // rad: the bin's cone is widened by this much on both sides
constexpr float  kBeamSlack       = 0.25f;  // px: a list result is trusted up to d - slack
constexpr float  kBeamMaxAngle    = 100.0f; // rad: rays with a larger |angle| skip the table (bin rounding)

// The grid over the lane and the covered cells (shared by the host and the device builder).
struct BeamPlan
{
    double                x0{0}, y0{0}, h{0}, rb{0};
    int32_t               nx{0}, ny{0}, nb{0};
    std::vector<uint32_t> rows;    // per cell: row, 0xffffffff = not covered
    std::vector<uint32_t> covered; // per row: cell
};
bool beam_plan(const Track &t, const BeamConfig &cfg, BeamPlan &plan, std::string &err);
// section offsets of a table with n_items uint16 items
bool beam_layout(const BeamPlan &plan, const BeamConfig &cfg, size_t n_items, BeamHeader &h, std::string &err);
// entries: 2 x uint32 per (row, bin) = {first chunk, count | dq << 16}; items: uint16 chunks of 4
bool beam_assemble(const BeamPlan &plan, const BeamConfig &cfg, const uint32_t *entries, const uint16_t *items, size_t n_items,
                   std::vector<uint8_t> &blob, std::string &err);

// Builds the blob on the host.  False (with err) if the configuration is unusable; never throws.
bool build_beam_table(const Track &t, const BeamConfig &cfg, std::vector<uint8_t> &blob, std::string &err);

// The same table built by a kernel on CUDA device `device` (ok_beam_gpu.cu): one CTA per covered cell.  Same
// construction, binary64 on the device; libm differences may move a list boundary by an ulp, never its validity.
// The blob is left in device memory (*d_blob, cudaMalloc'ed on `device`, owned by the caller).
bool build_beam_table_device(const Track &t, const BeamConfig &cfg, int device, uint8_t **d_blob, size_t *bytes, std::string &err);

// Host-side lookup used by the CPU tests: the list of (x, y, angle[rad]); returns false when the cell is not
// covered or the angle is out of range.  d_out = completeness distance.
bool beam_lookup(const std::vector<uint8_t> &blob, float x, float y, float angle, std::vector<uint16_t> &items,
                 float &d_out);

} // namespace ok
