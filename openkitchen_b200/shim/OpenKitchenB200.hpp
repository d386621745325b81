// OpenKitchenB200.hpp -- source-compatible C++ surface of OpenKitchen's `Environment/` library on top of
// the B200 C ABI (include/openkitchen_b200.h).
//
// An application written against the reference (e.g. Template/main.cpp, RLRacers/PPO/ppo_sim.cpp) keeps
// its `#include "Environment/Environment.h"` / `"Environment/Agent.h"` lines (forwarding headers live next
// to this file), its `Agent` subclasses with the public fields it reads and writes, and its
// `Environment env(path, createBaseAgentPtrs(agents)); env.step(); env.resetAgent(a.get());` loop.  What
// changes underneath:
//   Environment::step()        Environment.cpp:125-149  -> gather actions/state, ONE ok_launch_step, scatter
//   CollisionChecker           CollisionChecker.cu      -> part of that kernel (no per-step Ray_ memcpy pair)
//   TrackSegments / RaceTrack  TrackSegments.cu, RaceTrack.cpp -> ok_load_track_csv; arrays mirrored on host
//   Visualizer / ScreenGrabber Visualizer.cpp, ScreenGrabber.cu -> inert stubs (rendering is out of scope)
// The object model costs a host round trip per step; it is the compatibility path.  Batch
// users go through the C ABI or openkitchen_b200.BatchEnv and never touch host memory.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

struct OkEnv;
namespace okshim
{
struct Engine; // one OkEnv + the pinned staging blocks of the packed tick (OpenKitchenB200.cpp)
}

// ---- Typedefs.h ------------------------------------------------------------------------------------
constexpr int   kScreenWidth  = 1600;
constexpr int   kScreenHeight = 1400;
constexpr float kDeg2Rad      = static_cast<float>(M_PI / 180.0);

#ifndef GOX_ASSERT
#define GOX_ASSERT(cond)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
            std::terminate();                                                                                          \
    } while (0)
#endif

struct Pixel
{
    int x{0};
    int y{0};
};

struct Vec2d
{
    float x{0.F};
    float y{0.F};

    float squaredNorm() const { return x * x + y * y; }
    float norm() const { return std::sqrt(squaredNorm()); }
    float length() const { return norm(); }
    float distanceSquared(const Vec2d &o) const { return (x - o.x) * (x - o.x) + (y - o.y) * (y - o.y); }
    Vec2d operator+(const Vec2d &o) const { return {x + o.x, y + o.y}; }
    Vec2d operator-(const Vec2d &o) const { return {x - o.x, y - o.y}; }
    Vec2d operator*(float k) const { return {x * k, y * k}; }
    Vec2d operator/(float k) const { return {x / k, y / k}; }
    Vec2d operator/(const Vec2d &o) const { return {x / o.x, y / o.y}; }
};

struct Extent2d
{
    float min_x, min_y, max_x, max_y;
    bool  isPointInside(const Vec2d &p) const { return p.x > min_x && p.y > min_y && p.x < max_x && p.y < max_y; }
};

struct Ray_ // Typedefs.h:91-99
{
    float x, y, angle, hit_x, hit_y;
    bool  active{true};
};

struct Segment2d // Typedefs.h:101-105
{
    float x1, y1, x2, y2;
};

inline float normalizeAngleDeg(float angle) // Utils.h
{
    while (angle < 360.F)
        angle += 360.F;
    while (angle >= 360.F)
        angle -= 360.F;
    return angle;
}

// ---- Agent.h ---------------------------------------------------------------------------------------
// Host-side mirror of one agent.  Environment::step uploads the fields the device needs (pose, speed,
// acceleration, flags, current_action_) and writes the results back (pose, flags, sensor_hits_).
class Agent
{
  public:
    static constexpr float kSensorRange{200.F};
    static constexpr float kSpeedLimit{100.F};
    static constexpr float kRotationLimit{360.F};

    struct Action
    {
        float throttle_delta{0.F};
        float steering_delta{0.F};
    };
    enum class MovementMode
    {
        VELOCITY     = 0,
        ACCELERATION = 1,
        MANUAL       = 2
    };

    Agent() = default;
    Agent(Vec2d start_pos, float start_rot, int16_t id);
    virtual ~Agent() = default;

    virtual void reset(const Vec2d &reset_pos, const float reset_rot);
    virtual void updateAction() = 0;

    void setPose(const Vec2d pos, const float rot)
    {
        pos_ = pos;
        rot_ = rot;
    }
    bool isDone() const { return crashed_ || completed_; }
    void setMovementMode(const MovementMode mode) { movement_mode_ = mode; }
    void setHeadingDrawing(const bool on) { draw_agent_heading_ = on; }

    // the reference's host-side integrators are gone: movement happens on the device inside
    // Environment::step.  move() is kept so that code calling it directly still links; it throws.
    void move();

  public:
    Vec2d   pos_{};
    float   speed_{0.F};
    float   acceleration_{0.F};
    float   rot_{0.F};
    float   radius_{9.0F};
    float   sensor_offset_{0.0F};
    int16_t id_{};
    int     color_[4]{80, 80, 80, 255};

    bool has_raycast_sensor_{true};
    bool manual_control_enabled_{true};
    bool draw_agent_heading_{true};

    std::vector<float> sensor_ray_angles_;
    float              sensor_range_{kSensorRange};

    bool crashed_{false};
    bool completed_{false};
    bool timed_out_{false};

    std::vector<Vec2d> sensor_hits_;
    std::vector<Pixel> pixels_until_hit_;

    Action       current_action_{0.F, 0.F};
    MovementMode movement_mode_{MovementMode::VELOCITY};
};

template <typename TDerivedAgent>
inline std::vector<Agent *> createBaseAgentPtrs(const std::vector<std::unique_ptr<TDerivedAgent>> &derived)
{
    std::vector<Agent *> out;
    out.reserve(derived.size());
    for (const auto &a : derived)
        out.push_back(a.get());
    return out;
}

// ---- RaceTrack.h -----------------------------------------------------------------------------------
// Host mirror of the geometry the device was given (filled from ok_track_copy); the three queries are the
// reference's brute-force loops, kept for the apps that call them between steps.
class RaceTrack
{
  public:
    static constexpr size_t kStartingIdx{3};
    struct TrackData
    {
        std::vector<float> x_m, y_m, w_tr_right_m, w_tr_left_m;
    };

    RaceTrack() = delete;
    explicit RaceTrack(const std::string &track_csv_path); // stand-alone: builds through a host-only OkEnv
    RaceTrack(const OkEnv *env, int track_id, const std::string &track_csv_path);

    size_t findNearestTrackIndexBruteForce(const Vec2d &query_pt) const;
    float  getNearestDistanceToTrackBoundary(const Vec2d &query_pt) const;
    float  getDistanceToLaneCenter(const Vec2d &query_pt) const;

  public:
    std::string        track_name_{};
    TrackData          track_data_points_{};
    std::vector<Vec2d> left_bound_inner_, left_bound_outer_, right_bound_inner_, right_bound_outer_;
    std::vector<Vec2d> start_line_, finish_line_;
    std::vector<float> headings_{};
    std::vector<Segment2d> segments_; // TrackSegments order (not in the reference's RaceTrack; used by the shim)
    std::string            source_path_; // the CSV this track was built from (not in the reference; TrackSegments carries it on)

  private:
    void fill(const OkEnv *env, int track_id);
};

// ---- TrackSegments.h -------------------------------------------------------------------------------
// Same interface as the reference (TrackSegments.h:11-26).  The device copy of the segments lives inside an OkEnv's track
// arena, so getDeviceSegments() hands out an opaque, non-null HANDLE with the reference's type: like the reference's device
// pointer it must not be dereferenced on the host; its one consumer, CollisionChecker's constructor, resolves it.
class TrackSegments
{
  public:
    explicit TrackSegments(const RaceTrack &race_track);
    ~TrackSegments();
    TrackSegments(const TrackSegments &)            = delete;
    TrackSegments &operator=(const TrackSegments &) = delete;

    const Segment2d *getDeviceSegments() const { return reinterpret_cast<const Segment2d *>(&handle_); }
    size_t           getNumSegments() const { return host_.size(); }
    const Segment2d *getHostSegments() const { return host_.data(); } // not in the reference
    const std::string &sourcePath() const { return source_path_; }   // not in the reference

  private:
    std::vector<Segment2d> host_;
    std::string            source_path_;
    char                   handle_{0}; // its address is the handle
};

// ---- CollisionChecker.h ----------------------------------------------------------------------------
// Same interface as the reference (CollisionChecker.h:8-24): built from the segment handle and the agents, it casts
// every agent's rays on checkCollision() (CollisionChecker.cu:96-100) -- lidar + crash flags, no movement.
class Environment;
class CollisionChecker
{
  public:
    CollisionChecker(const Segment2d *d_segments, size_t num_segments, const std::vector<Agent *> &agents);
    ~CollisionChecker();
    void        checkCollision();
    const Ray_ *getHostRays() const;
    size_t      getNumRays() const;

  private:
    friend class Environment;
    explicit CollisionChecker(std::shared_ptr<okshim::Engine> engine); // the Environment's own checker shares its engine
    std::shared_ptr<okshim::Engine> engine_;
};

// ---- Visualizer.h / ScreenGrabber.h: inert -----------------------------------------------------------
namespace env
{
class Visualizer
{
  public:
    explicit Visualizer(bool hidden_window = false) : hidden_(hidden_window) {}
    void setAgentToFollow(const Agent *agent) { followed_ = agent; }
    void activateDrawing(bool) {}
    void disableDrawing() {}
    void enableDrawing() {}
    std::function<void()> user_draw_callback_{};

  private:
    bool         hidden_{true};
    const Agent *followed_{nullptr};
};
} // namespace env

class ScreenGrabber
{
  public:
    struct RenderTargetInfo
    {
        int    width{kScreenWidth};
        int    height{kScreenHeight};
        int    channels{4};
        size_t row_bytes() const { return static_cast<size_t>(width) * channels; }
    };
    ScreenGrabber(int w, int h) : info_{w, h, 4} {}
    std::vector<uint8_t> getRenderTargetHost() const { return std::vector<uint8_t>(info_.row_bytes() * info_.height, 0); }
    RenderTargetInfo     getRenderTargetInfo() const { return info_; }
    void                 saveRenderTargetToFile(const std::string &) const {}

  private:
    RenderTargetInfo info_;
};

// ---- Environment.h ---------------------------------------------------------------------------------
struct DisplacementStats
{
    static constexpr uint32_t kPeriod{200};
    static constexpr float    kDisplamentThreshold{20.0F};
    bool                      displacement_timed_out{false};
    uint32_t                  displacement_ctr{0U};
    Vec2d                     init_pos{0.F, 0.F};
};

class Environment
{
  public:
    Environment(const std::string &race_track_path, const std::vector<Agent *> &agents, const bool draw_rays = true,
                const bool hidden_window = false);
    ~Environment();
    Environment(const Environment &)            = delete;
    Environment &operator=(const Environment &) = delete;

    void    step();
    int32_t pickRandomResetTrackIdx() const;
    void    resetAgent(Agent *agent, const bool pick_random_point = true, const bool randomize_lane = false,
                       const bool randomize_heading = false);
    void    drawSensorRanges(const std::vector<Vec2d> &) {}
    bool    isEnterPressed() const { return false; }
    void    saveImage(const std::string &) const {}

    std::vector<uint8_t>            getRenderTargetHost() const { return screen_grabber_->getRenderTargetHost(); }
    ScreenGrabber::RenderTargetInfo getRenderTargetInfo() const { return screen_grabber_->getRenderTargetInfo(); }

    // not in the reference: its resets draw from raylib's never-seeded global RNG (Environment.cpp:76,92,111)
    void   seed(uint64_t s) { rng_.seed(s); }
    OkEnv *handle() const;

  public:
    std::unique_ptr<RaceTrack>        race_track_;
    std::unique_ptr<TrackSegments>    track_segments_;
    std::unique_ptr<env::Visualizer>  visualizer_;
    std::vector<Agent *>              agents_;
    std::vector<DisplacementStats>    displacement_stats_;
    std::unique_ptr<CollisionChecker> collision_checker_{nullptr};
    std::unique_ptr<ScreenGrabber>    screen_grabber_{nullptr};
    bool                              draw_rays_{false};

  private:
    std::shared_ptr<okshim::Engine> engine_; // OkEnv + packed staging: ONE upload, one launch, two downloads per step
    mutable std::mt19937            rng_{0x0C17C4E2u};
    size_t                          heading_ctr_{0};
};
