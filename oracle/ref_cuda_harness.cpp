// ref_cuda_harness.cpp -- runs the reference's REAL CollisionChecker / TrackSegments objects
// (Environment/CollisionChecker.cu, TrackSegments.cu compiled unchanged for sm_100a) on a GPU.
// TEST INFRASTRUCTURE ONLY: second, independent check of the lidar (tests/test_gpu_ref_kernel.py)
// and the "existing GPU kernel" baseline row of bench.py.  Needs a GPU at run time.
#include <chrono>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "Agent.h"            // reference
#include "CollisionChecker.h" // reference
#include "RaceTrack.h"        // reference
#include "TrackSegments.h"    // reference

namespace
{
class FakeAgent : public Agent
{
  public:
    FakeAgent() : Agent(Vec2d{0.F, 0.F}, 0.F, 0)
    {
    }
    void updateAction() override
    {
    }
};
} // namespace

struct OkcEnv
{
    std::unique_ptr<RaceTrack>              track;
    std::unique_ptr<TrackSegments>          segments;
    std::vector<std::unique_ptr<FakeAgent>> agents;
    std::unique_ptr<CollisionChecker>       checker;
    int                                     rays{0};
};

extern "C"
{
OkcEnv *okc_create(const char *csv_path, int64_t n_agents, int rays, const float *ray_deg)
{
    auto *e     = new OkcEnv();
    e->track    = std::make_unique<RaceTrack>(std::string(csv_path));
    e->segments = std::make_unique<TrackSegments>(*e->track);
    e->rays     = rays;
    for (int64_t i = 0; i < n_agents; ++i)
    {
        auto a = std::make_unique<FakeAgent>();
        a->sensor_ray_angles_.assign(ray_deg, ray_deg + rays);
        e->agents.push_back(std::move(a));
    }
    e->checker = std::make_unique<CollisionChecker>(
        e->segments->getDeviceSegments(), e->segments->getNumSegments(), createBaseAgentPtrs(e->agents));
    return e;
}

void okc_destroy(OkcEnv *e)
{
    delete e;
}

int okc_num_segments(const OkcEnv *e)
{
    return static_cast<int>(e->segments->getNumSegments());
}

void okc_set_poses(OkcEnv *e, const float *x, const float *y, const float *rot, const uint8_t *crashed)
{
    for (size_t i = 0; i < e->agents.size(); ++i)
    {
        e->agents[i]->pos_     = Vec2d{x[i], y[i]};
        e->agents[i]->rot_     = rot[i];
        e->agents[i]->crashed_ = crashed ? crashed[i] != 0 : false;
    }
}

void okc_check(OkcEnv *e)
{
    e->checker->checkCollision();
}

void okc_get(const OkcEnv *e, float *hit_rel, float *hit_abs, uint8_t *crashed)
{
    const Ray_ *rays = e->checker->getHostRays();
    for (size_t i = 0; i < e->agents.size(); ++i)
    {
        for (int r = 0; r < e->rays; ++r)
        {
            const size_t k = i * e->rays + r;
            if (hit_rel)
            {
                hit_rel[2 * k]     = e->agents[i]->sensor_hits_[r].x;
                hit_rel[2 * k + 1] = e->agents[i]->sensor_hits_[r].y;
            }
            if (hit_abs)
            {
                hit_abs[2 * k]     = rays[k].hit_x;
                hit_abs[2 * k + 1] = rays[k].hit_y;
            }
        }
        if (crashed)
            crashed[i] = e->agents[i]->crashed_;
    }
}

// seconds for `iters` back-to-back checkCollision() calls (pack + H2D + kernel + D2H + unpack)
double okc_time_checks(OkcEnv *e, int iters)
{
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i)
    {
        for (auto &a : e->agents)
            a->crashed_ = false;
        e->checker->checkCollision();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
}
