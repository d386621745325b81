// OpenKitchenB200.cpp -- see OpenKitchenB200.hpp.  Everything here is a thin host adapter over the C ABI.
#include "OpenKitchenB200.hpp"

#include "../../include/openkitchen_b200.h"

#include <algorithm>
#include <cctype>
#include <limits>
#include <map>
#include <mutex>

namespace
{
[[noreturn]] void raise(const char *what)
{
    throw std::runtime_error(std::string(what) + ": " + ok_last_error());
}

void must(int rc, const char *what)
{
    if (rc < 0)
        raise(what);
}

std::string upper_stem(const std::string &path)
{ // RaceTrack::getTrackName, RaceTrack.cpp:116-125
    const size_t slash = path.rfind('/'), dot = path.rfind('.');
    std::string  s     = path.substr(slash == std::string::npos ? 0 : slash + 1,
                                     (dot == std::string::npos ? path.size() : dot) - (slash == std::string::npos ? 0 : slash + 1));
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::toupper(c); });
    return s;
}

std::vector<float> track_array(const OkEnv *env, int id, int which)
{
    const int64_t n = ok_track_copy(env, id, which, nullptr, 0);
    if (n < 0)
        raise("ok_track_copy");
    std::vector<float> v(static_cast<size_t>(n));
    ok_track_copy(env, id, which, v.data(), n);
    return v;
}

std::vector<Vec2d> as_points(const std::vector<float> &flat)
{
    std::vector<Vec2d> out(flat.size() / 2);
    for (size_t i = 0; i < out.size(); ++i)
        out[i] = Vec2d{flat[2 * i], flat[2 * i + 1]};
    return out;
}
} // namespace

// ---- Agent -------------------------------------------------------------------------------------------
Agent::Agent(Vec2d start_pos, float start_rot, int16_t id) : pos_{start_pos}, rot_{start_rot}, id_{id}
{
    // default fan: -70..70 in steps of 10 degrees (Agent.cpp:8-19)
    for (int deg = -70; deg <= 70; deg += 10)
        sensor_ray_angles_.push_back(static_cast<float>(deg));
}

void Agent::reset(const Vec2d &reset_pos, const float reset_rot)
{ // Agent.cpp:123-135
    pos_            = reset_pos;
    rot_            = reset_rot;
    acceleration_   = 0.F;
    speed_          = 0.F;
    crashed_        = false;
    timed_out_      = false;
    completed_      = false;
    current_action_ = Action{0.F, 0.F};
}

void Agent::move()
{
    throw std::logic_error("Agent::move(): kinematics run on the device inside Environment::step()");
}

// ---- RaceTrack ---------------------------------------------------------------------------------------
RaceTrack::RaceTrack(const OkEnv *env, int track_id, const std::string &track_csv_path)
    : track_name_{upper_stem(track_csv_path)}, source_path_{track_csv_path}
{
    fill(env, track_id);
}

RaceTrack::RaceTrack(const std::string &track_csv_path) : track_name_{upper_stem(track_csv_path)}, source_path_{track_csv_path}
{
    OkConfig cfg;
    ok_config_default(&cfg);
    cfg.device = -1; // host-only: geometry without a GPU
    OkEnv *env = nullptr;
    must(ok_create(&cfg, &env), "ok_create");
    int32_t id = -1;
    if (ok_load_track_csv(env, track_csv_path.c_str(), &id) < 0)
    {
        ok_destroy(env);
        raise("ok_load_track_csv");
    }
    fill(env, id);
    ok_destroy(env);
}

void RaceTrack::fill(const OkEnv *env, int id)
{
    track_data_points_.x_m          = track_array(env, id, OK_TRACK_X);
    track_data_points_.y_m          = track_array(env, id, OK_TRACK_Y);
    track_data_points_.w_tr_right_m = track_array(env, id, OK_TRACK_W_RIGHT);
    track_data_points_.w_tr_left_m  = track_array(env, id, OK_TRACK_W_LEFT);
    headings_                       = track_array(env, id, OK_TRACK_HEADING);
    left_bound_inner_               = as_points(track_array(env, id, OK_TRACK_LEFT_INNER));
    left_bound_outer_               = as_points(track_array(env, id, OK_TRACK_LEFT_OUTER));
    right_bound_inner_              = as_points(track_array(env, id, OK_TRACK_RIGHT_INNER));
    right_bound_outer_              = as_points(track_array(env, id, OK_TRACK_RIGHT_OUTER));
    start_line_                     = {right_bound_outer_.front(), right_bound_outer_.back()};
    finish_line_                    = {left_bound_outer_.front(), left_bound_outer_.back()};
    const std::vector<float> s      = track_array(env, id, OK_TRACK_SEGMENTS);
    segments_.resize(s.size() / 4);
    for (size_t i = 0; i < segments_.size(); ++i)
        segments_[i] = Segment2d{s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]};
}

size_t RaceTrack::findNearestTrackIndexBruteForce(const Vec2d &q) const
{ // RaceTrack.cpp:16-31: strict '<' keeps the lowest index among ties
    float  best = std::numeric_limits<float>::max();
    size_t at   = 0;
    for (size_t i = 0; i < track_data_points_.x_m.size(); ++i)
    {
        const float d = q.distanceSquared({track_data_points_.x_m[i], track_data_points_.y_m[i]});
        if (d < best)
        {
            best = d;
            at   = i;
        }
    }
    return at;
}

float RaceTrack::getNearestDistanceToTrackBoundary(const Vec2d &q) const
{ // RaceTrack.cpp:33-51
    float best = std::numeric_limits<float>::max();
    for (size_t i = 0; i < left_bound_inner_.size(); ++i)
    {
        best = std::min(best, q.distanceSquared(left_bound_inner_[i]));
        best = std::min(best, q.distanceSquared(right_bound_inner_[i]));
    }
    return std::sqrt(best);
}

float RaceTrack::getDistanceToLaneCenter(const Vec2d &q) const
{ // RaceTrack.cpp:53-72
    const size_t i = findNearestTrackIndexBruteForce(q);
    const float  d = q.distanceSquared({track_data_points_.x_m[i], track_data_points_.y_m[i]});
    return std::sqrt(d) / (track_data_points_.w_tr_left_m[i] + track_data_points_.w_tr_right_m[i]);
}

// ---- the engine: one OkEnv + the packed tick --------------------------------------------------------------
namespace okshim
{
struct Engine
{
    OkEnv               *ok{nullptr};
    std::vector<Agent *> agents;
    std::vector<DisplacementStats> *stats{nullptr}; // the Environment's standstill records (nullptr: stand-alone checker)
    std::vector<DisplacementStats>  own_stats;
    OkPackedLayout       lay{};
    uint8_t             *h_state{nullptr}, *h_hits{nullptr}; // pinned images of the device blocks
    std::vector<Ray_>    rays;
    size_t               n_rays{0};

    Engine(const std::string &track_csv_path, const std::vector<Agent *> &agents_in) : agents(agents_in)
    {
        if (agents.empty())
            throw std::invalid_argument("at least one agent is needed");
        OkConfig cfg;
        ok_config_default(&cfg);
        cfg.movement_mode = agents[0]->movement_mode_ == Agent::MovementMode::ACCELERATION ? OK_MOVE_ACCELERATION : OK_MOVE_VELOCITY;
        cfg.sensor_offset = agents[0]->sensor_offset_;
        cfg.sensor_range  = agents[0]->sensor_range_;
        must(ok_create(&cfg, &ok), "ok_create");
        try
        {
            int32_t track = -1;
            must(ok_load_track_csv(ok, track_csv_path.c_str(), &track), "ok_load_track_csv");
            // the reference sizes everything from agents[0]'s fan (CollisionChecker.cu:82) and silently assumes the
            // others match; here a mismatch is an error
            const auto &fan = agents[0]->sensor_ray_angles_;
            for (const Agent *a : agents)
                if (a->sensor_ray_angles_ != fan)
                    throw std::invalid_argument("all agents of one Environment must share one sensor_ray_angles_ fan");
            n_rays = fan.size();
            must(ok_alloc_agents(ok, static_cast<int64_t>(agents.size()), static_cast<int32_t>(fan.size()), fan.data(), nullptr),
                 "ok_alloc_agents");
            must(ok_packed_layout(ok, &lay), "ok_packed_layout");
            must(ok_host_alloc(reinterpret_cast<void **>(&h_state), lay.state_bytes), "ok_host_alloc");
            must(ok_host_alloc(reinterpret_cast<void **>(&h_hits), lay.hits_bytes), "ok_host_alloc");
            std::fill(h_state, h_state + lay.state_bytes, 0);
        }
        catch (...)
        {
            release();
            throw;
        }
        rays.assign(agents.size() * n_rays, Ray_{});
        own_stats.resize(agents.size());
        stats = &own_stats;
    }
    Engine(const Engine &)            = delete;
    Engine &operator=(const Engine &) = delete;
    ~Engine() { release(); }

    void release()
    {
        if (h_state)
            ok_host_free(h_state);
        if (h_hits)
            ok_host_free(h_hits);
        if (ok)
            ok_destroy(ok);
        h_state = h_hits = nullptr;
        ok               = nullptr;
    }

    template <typename T> T *col(int buffer) { return reinterpret_cast<T *>(h_state + lay.state_offset[buffer]); }

    // Environment::step (Environment.cpp:125-149 minus the render) or CollisionChecker::checkCollision: the Agent objects'
    // fields go up in one block, one kernel runs, pose / flags / standstill record and the hits come back in two blocks
    void run(bool move)
    {
        // Agent::setMovementMode may be called at any time (Agent.h:46-49); the reference reads it per move()
        const int mode = agents[0]->movement_mode_ == Agent::MovementMode::ACCELERATION ? OK_MOVE_ACCELERATION : OK_MOVE_VELOCITY;
        for (const Agent *a : agents)
        {
            if (a->movement_mode_ == Agent::MovementMode::MANUAL)
                throw std::logic_error("MovementMode::MANUAL (keyboard) is not supported");
            if ((a->movement_mode_ == Agent::MovementMode::ACCELERATION ? OK_MOVE_ACCELERATION : OK_MOVE_VELOCITY) != mode)
                throw std::logic_error("all agents of one Environment must use the same MovementMode");
        }
        OkConfig cfg;
        must(ok_get_config(ok, &cfg), "ok_get_config");
        if (cfg.movement_mode != mode)
        {
            cfg.movement_mode = mode;
            must(ok_update_config(ok, &cfg), "ok_update_config");
        }
        const size_t n = agents.size();
        float *x = col<float>(OK_BUF_POS_X), *y = col<float>(OK_BUF_POS_Y), *rot = col<float>(OK_BUF_ROT);
        float *speed = col<float>(OK_BUF_SPEED), *accel = col<float>(OK_BUF_ACCEL);
        float *thr = col<float>(OK_BUF_ACT_THROTTLE), *steer = col<float>(OK_BUF_ACT_STEER);
        uint8_t  *crashed = col<uint8_t>(OK_BUF_CRASHED), *timed_out = col<uint8_t>(OK_BUF_TIMED_OUT), *done = col<uint8_t>(OK_BUF_DONE);
        uint32_t *ctr = col<uint32_t>(OK_BUF_SS_CTR);
        float    *ssx = col<float>(OK_BUF_SS_X), *ssy = col<float>(OK_BUF_SS_Y);
        for (size_t i = 0; i < n; ++i)
        {
            const Agent *a = agents[i];
            x[i] = a->pos_.x, y[i] = a->pos_.y, rot[i] = a->rot_, speed[i] = a->speed_, accel[i] = a->acceleration_;
            thr[i] = a->current_action_.throttle_delta, steer[i] = a->current_action_.steering_delta;
            crashed[i] = a->crashed_, timed_out[i] = a->timed_out_, done[i] = a->crashed_;
            const DisplacementStats &ds = (*stats)[i];
            ctr[i] = ds.displacement_ctr, ssx[i] = ds.init_pos.x, ssy[i] = ds.init_pos.y;
        }
        must(ok_step_packed(ok, h_state, h_hits, move ? 1 : 0, nullptr), "ok_step_packed");
        const float *hit_abs = reinterpret_cast<const float *>(h_hits + lay.hit_abs_offset);
        const float *hit_rel = reinterpret_cast<const float *>(h_hits + lay.hit_rel_offset);
        const size_t R = n_rays;
        for (size_t i = 0; i < n; ++i)
        {
            Agent *a = agents[i];
            if (move)
            {
                a->pos_ = Vec2d{x[i], y[i]};
                a->rot_ = rot[i], a->speed_ = speed[i], a->acceleration_ = accel[i];
                DisplacementStats &ds     = (*stats)[i];
                ds.displacement_ctr       = ctr[i];
                ds.init_pos               = Vec2d{ssx[i], ssy[i]};
                ds.displacement_timed_out = timed_out[i] != 0;
            }
            a->crashed_   = crashed[i] != 0;
            a->timed_out_ = timed_out[i] != 0;
            a->sensor_hits_.resize(R);
            for (size_t r = 0; r < R; ++r)
            {
                const size_t k     = i * R + r;
                a->sensor_hits_[r] = Vec2d{hit_rel[2 * k], hit_rel[2 * k + 1]};
                Ray_ &ray          = rays[k];
                ray.x              = a->pos_.x + a->sensor_offset_ * std::cos(kDeg2Rad * a->rot_);
                ray.y              = a->pos_.y + a->sensor_offset_ * std::sin(kDeg2Rad * a->rot_);
                ray.angle          = kDeg2Rad * (a->rot_ + a->sensor_ray_angles_[r]);
                ray.hit_x          = hit_abs[2 * k];
                ray.hit_y          = hit_abs[2 * k + 1];
                ray.active         = !a->crashed_;
            }
        }
    }
};
} // namespace okshim

// ---- TrackSegments -----------------------------------------------------------------------------------
namespace
{
std::mutex                                  g_seg_mu;
std::map<const void *, const TrackSegments *> g_segments; // handle -> owner
} // namespace

TrackSegments::TrackSegments(const RaceTrack &race_track) : host_(race_track.segments_), source_path_(race_track.source_path_)
{
    GOX_ASSERT(!host_.empty()); // TrackSegments.cu:72
    std::lock_guard<std::mutex> lock(g_seg_mu);
    g_segments[&handle_] = this;
}

TrackSegments::~TrackSegments()
{
    std::lock_guard<std::mutex> lock(g_seg_mu);
    g_segments.erase(&handle_);
}

// ---- CollisionChecker --------------------------------------------------------------------------------
CollisionChecker::CollisionChecker(const Segment2d *d_segments, size_t num_segments, const std::vector<Agent *> &agents)
{
    std::string path;
    {
        std::lock_guard<std::mutex> lock(g_seg_mu);
        const auto                  it = g_segments.find(d_segments);
        if (it == g_segments.end() || it->second->getNumSegments() != num_segments)
            throw std::invalid_argument("CollisionChecker: d_segments must be what a live TrackSegments::getDeviceSegments() returned");
        path = it->second->sourcePath();
    }
    engine_ = std::make_shared<okshim::Engine>(path, agents);
}

CollisionChecker::CollisionChecker(std::shared_ptr<okshim::Engine> engine) : engine_(std::move(engine)) {}
CollisionChecker::~CollisionChecker() = default;

void CollisionChecker::checkCollision()
{
    engine_->run(false);
}
const Ray_ *CollisionChecker::getHostRays() const
{
    return engine_->rays.data();
}
size_t CollisionChecker::getNumRays() const
{
    return engine_->rays.size();
}

// ---- Environment -------------------------------------------------------------------------------------
Environment::Environment(const std::string &race_track_path, const std::vector<Agent *> &agents, const bool draw_rays,
                         const bool hidden_window)
    : draw_rays_(draw_rays)
{
    if (agents.empty())
        throw std::invalid_argument("Environment needs at least one agent");
    engine_         = std::make_shared<okshim::Engine>(race_track_path, agents);
    race_track_     = std::make_unique<RaceTrack>(engine_->ok, 0, race_track_path);
    track_segments_ = std::make_unique<TrackSegments>(*race_track_);
    visualizer_     = std::make_unique<env::Visualizer>(hidden_window);
    screen_grabber_ = std::make_unique<ScreenGrabber>(kScreenWidth, kScreenHeight);
    agents_         = agents;
    displacement_stats_.resize(agents.size());
    engine_->stats = &displacement_stats_;
    collision_checker_.reset(new CollisionChecker(engine_));
}

Environment::~Environment() = default;

OkEnv *Environment::handle() const
{
    return engine_->ok;
}
int32_t Environment::pickRandomResetTrackIdx() const
{ // Environment.cpp:74-77 (GetRandomValue(0, n-1) -> a seedable generator)
    std::uniform_int_distribution<int32_t> d(0, static_cast<int32_t>(race_track_->track_data_points_.x_m.size()) - 1);
    return d(rng_);
}

void Environment::resetAgent(Agent *agent, const bool pick_random_point, const bool randomize_lane, const bool randomize_heading)
{ // Environment.cpp:79-122.  The virtual Agent::reset is what derived agents hook, so the reset happens on
  // the host object; the new pose reaches the device with the next step()'s upload.
    const int32_t idx = pick_random_point ? pickRandomResetTrackIdx() : static_cast<int32_t>(RaceTrack::kStartingIdx);
    float         heading_offset = 0.F;
    if (pick_random_point && randomize_heading)
    {
        std::uniform_int_distribution<int> d(0, 45);
        const float                        mag = static_cast<float>(d(rng_)) + 45.F;
        heading_offset                         = (heading_ctr_++ % 2 == 0) ? -mag : mag;
    }
    Vec2d start;
    if (pick_random_point && randomize_lane)
    {
        std::uniform_int_distribution<int> d(10, 90);
        const float                        alpha = static_cast<float>(d(rng_)) / 100.F;
        const Vec2d                        l = race_track_->left_bound_inner_[idx], r = race_track_->right_bound_inner_[idx];
        start = Vec2d{l.x * alpha + r.x * (1.F - alpha), l.y * alpha + r.y * (1.F - alpha)};
    }
    else
        start = Vec2d{race_track_->track_data_points_.x_m[idx], race_track_->track_data_points_.y_m[idx]};
    agent->reset(start, race_track_->headings_[idx] + heading_offset);
}

void Environment::step()
{ // Environment.cpp:125-149 minus the render
    engine_->run(true);
}
