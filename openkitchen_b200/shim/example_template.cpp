// example_template.cpp -- a Template/main.cpp-shaped application written against the reference's class
// surface only (Agent subclass, Environment, createBaseAgentPtrs, resetAgent, step, public fields).  It
// builds against openkitchen_b200/shim unchanged from how it would build against the reference's
// Environment/ (minus raylib), and doubles as the shim's end-to-end test: with --trace it prints every
// action and the resulting state so tests/test_gpu_shim.py can replay the run through the oracle.
#include "Environment/Environment.h"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace
{
class RandomDriver : public Agent
{
  public:
    RandomDriver(Vec2d pos, float rot, int16_t id, unsigned seed) : Agent(pos, rot, id), rng_(seed)
    {
        sensor_ray_angles_ = {-70.F, -30.F, 0.F, 30.F, 70.F}; // the Q-learning fan, QAgent.hpp:56-62
        setMovementMode(MovementMode::VELOCITY);
    }
    void updateAction() override
    {
        // steer towards the freer side, drive faster when the front ray is long
        const float left = sensor_hits_.empty() ? 0.F : sensor_hits_.front().norm();
        const float right = sensor_hits_.empty() ? 0.F : sensor_hits_.back().norm();
        const float front = sensor_hits_.empty() ? 0.F : sensor_hits_[2].norm();
        std::uniform_real_distribution<float> jitter(-1.F, 1.F);
        current_action_.throttle_delta = 20.F + 0.3F * front;
        current_action_.steering_delta = (right > left ? 3.F : -3.F) + jitter(rng_);
    }
    int episodes_{0};

  private:
    std::mt19937 rng_;
};
} // namespace

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s <track.csv> [steps] [agents] [--trace]\n", argv[0]);
        return 2;
    }
    const int  steps    = argc > 2 ? std::atoi(argv[2]) : 300;
    const int  n_agents = argc > 3 ? std::atoi(argv[3]) : 4;
    const bool trace    = argc > 4 && std::string(argv[4]) == "--trace";

    std::vector<std::unique_ptr<RandomDriver>> agents;
    for (int i = 0; i < n_agents; ++i)
        agents.push_back(std::make_unique<RandomDriver>(Vec2d{0.F, 0.F}, 0.F, static_cast<int16_t>(i), 100u + i));

    Environment env(argv[1], createBaseAgentPtrs(agents), /*draw_rays=*/false, /*hidden_window=*/true);
    env.seed(42);
    env.visualizer_->setAgentToFollow(agents[0].get());
    for (auto &a : agents)
        env.resetAgent(a.get(), /*pick_random_point=*/true);
    env.step(); // initial observation after reset (Template/main.cpp:102-103)

    for (int s = 0; s < steps; ++s)
    {
        for (auto &a : agents)
        {
            if (a->crashed_)
            {
                env.resetAgent(a.get(), true, /*randomize_lane=*/true, /*randomize_heading=*/false);
                a->episodes_++;
            }
            else
                a->updateAction();
        }
        if (trace)
            for (auto &a : agents)
                std::printf("A %d %d %.9g %.9g %.9g %.9g %.9g %d\n", s, a->id_, a->current_action_.throttle_delta,
                            a->current_action_.steering_delta, a->pos_.x, a->pos_.y, a->rot_, a->crashed_ ? 1 : 0);
        env.step();
        if (trace)
            for (auto &a : agents)
                std::printf("S %d %d %.9g %.9g %.9g %d %d %.9g %.9g\n", s, a->id_, a->pos_.x, a->pos_.y, a->rot_,
                            a->crashed_ ? 1 : 0, a->timed_out_ ? 1 : 0, a->sensor_hits_[2].x, a->sensor_hits_[2].y);
    }
    // the lidar alone, the way Environment.cpp:49-54 wires it: RaceTrack -> TrackSegments -> CollisionChecker(device
    // segments, count, agents) -> checkCollision() on the agents' current poses (reference signatures throughout)
    if (trace)
    {
        RaceTrack        track(argv[1]);
        TrackSegments    segments(track);
        CollisionChecker checker(segments.getDeviceSegments(), segments.getNumSegments(), createBaseAgentPtrs(agents));
        for (auto &a : agents)
            a->crashed_ = false; // cast for everyone
        checker.checkCollision();
        const Ray_ *rays = checker.getHostRays();
        for (size_t k = 0; k < checker.getNumRays(); ++k)
            std::printf("R %zu %.9g %.9g %.9g %.9g %.9g %d\n", k, rays[k].x, rays[k].y, rays[k].angle, rays[k].hit_x, rays[k].hit_y,
                        rays[k].active ? 1 : 0);
        for (auto &a : agents)
            std::printf("C %d %d %.9g %.9g\n", a->id_, a->crashed_ ? 1 : 0, a->sensor_hits_[0].x, a->sensor_hits_[0].y);
    }
    int         episodes = 0;
    const auto &rt       = *env.race_track_;
    for (auto &a : agents)
        episodes += a->episodes_;
    std::printf("done: %d agents x %d steps on %s (%zu points, %zu segments), %d episodes, agent0 nearest idx %zu\n", n_agents, steps,
                rt.track_name_.c_str(), rt.track_data_points_.x_m.size(), env.track_segments_->getNumSegments(), episodes,
                rt.findNearestTrackIndexBruteForce(agents[0]->pos_));
    return 0;
}
