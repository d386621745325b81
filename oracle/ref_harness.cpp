// ref_harness.cpp -- drives the UNMODIFIED reference objects (TEST INFRASTRUCTURE ONLY).
//
// Built by oracle/Makefile into oracle/_ref/libokref.so together with the reference's own
// Environment/Agent.cpp and Environment/RaceTrack.cpp, compiled from where they lie under
// /root/reference (never copied into this repo).  The reference classes do the work they can:
//   * RaceTrack            -> CSV parse, scaling, headings, boundary polylines, nearest-index queries
//   * Agent::move / reset  -> kinematics and reset
// Environment.cpp, Visualizer.cpp and CollisionChecker.cu cannot be built for a CPU (raylib /
// GLEW / a CUDA kernel), so this file restates ONLY those pieces, against the real objects:
//   * Environment::step stages 1-2 and checkAndUpdateStandstill   (Environment.cpp:16-39,125-146)
//   * Environment::resetAgent with explicit inputs                 (Environment.cpp:79-122)
//   * TrackSegments ordering                                       (TrackSegments.cu:6-42,53-67)
//   * the raycast kernel and its host pack/unpack, with libm cosf/sinf standing in for
//     libdevice's (CollisionChecker.cu:8-71,113-172)
// It exports the same C API as ok_oracle.h under the okr_ prefix so one ctypes wrapper
// serves both libraries.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unistd.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "Agent.h"     // reference header
#include "RaceTrack.h" // reference header
#include "Utils.h"     // reference header

#include "ok_oracle.h" // buffer ids / config layout only

namespace
{
class HarnessAgent : public Agent
{
  public:
    HarnessAgent() : Agent(Vec2d{0.F, 0.F}, 0.F, 0)
    {
    }
    void updateAction() override
    {
    }
};

// Environment.h:16-26
struct DisplacementStats
{
    static constexpr uint32_t kPeriod{200};
    static constexpr float    kDisplamentThreshold{20.0F};
    bool                      displacement_timed_out{false};
    uint32_t                  displacement_ctr{0U};
    Vec2d                     init_pos{0.F, 0.F};
};

struct Track
{
    std::unique_ptr<RaceTrack> rt;
    std::vector<Segment2d>     segments;
    std::vector<float>         li, lo, ri, ro;
};
} // namespace

struct OkrEnv
{
    OkoConfig                                  cfg;
    std::vector<Track>                         tracks;
    std::vector<std::unique_ptr<HarnessAgent>> agents;
    std::vector<DisplacementStats>             stats;
    std::vector<Ray_>                          rays; // the pinned h_rays_ of CollisionChecker.cu:85 (keeps stale hits)
    int                                        nrays{0};
    std::vector<std::vector<uint8_t>>          buf;
};

namespace
{
void addPolyline(const std::vector<Vec2d> &pl, std::vector<Segment2d> &out)
{
    if (pl.size() < 2)
        return;
    for (size_t i = 0; i + 1 < pl.size(); ++i)
        out.push_back(Segment2d{pl[i].x, pl[i].y, pl[i + 1].x, pl[i + 1].y});
}

void flatten(const std::vector<Vec2d> &v, std::vector<float> &out)
{
    out.clear();
    for (auto &p : v)
    {
        out.push_back(p.x);
        out.push_back(p.y);
    }
}

int finishTrack(OkrEnv *e, std::unique_ptr<RaceTrack> rt)
{
    Track t;
    t.rt = std::move(rt);
    const RaceTrack &r = *t.rt;
    addPolyline(r.left_bound_inner_, t.segments);
    addPolyline(r.left_bound_outer_, t.segments);
    addPolyline(r.right_bound_inner_, t.segments);
    addPolyline(r.right_bound_outer_, t.segments);
    if (r.left_bound_inner_.size() > 1)
    {
        t.segments.push_back({r.left_bound_inner_.back().x,
                              r.left_bound_inner_.back().y,
                              r.left_bound_inner_.front().x,
                              r.left_bound_inner_.front().y});
        t.segments.push_back({r.right_bound_inner_.back().x,
                              r.right_bound_inner_.back().y,
                              r.right_bound_inner_.front().x,
                              r.right_bound_inner_.front().y});
    }
    if (r.left_bound_outer_.size() > 1)
    {
        t.segments.push_back({r.left_bound_outer_.back().x,
                              r.left_bound_outer_.back().y,
                              r.left_bound_outer_.front().x,
                              r.left_bound_outer_.front().y});
        t.segments.push_back({r.right_bound_outer_.back().x,
                              r.right_bound_outer_.back().y,
                              r.right_bound_outer_.front().x,
                              r.right_bound_outer_.front().y});
    }
    flatten(r.left_bound_inner_, t.li);
    flatten(r.left_bound_outer_, t.lo);
    flatten(r.right_bound_inner_, t.ri);
    flatten(r.right_bound_outer_, t.ro);
    e->tracks.push_back(std::move(t));
    return static_cast<int>(e->tracks.size()) - 1;
}

template <typename T> T *B(OkrEnv *e, int id)
{
    return reinterpret_cast<T *>(e->buf[id].data());
}

// checkAndUpdateStandstill, Environment.cpp:16-39
void checkAndUpdateStandstill(const Agent &agent, DisplacementStats &ds)
{
    if (ds.displacement_ctr == 0)
    {
        ds.init_pos               = agent.pos_;
        ds.displacement_timed_out = false;
        ds.displacement_ctr++;
        return;
    }
    if (ds.displacement_ctr >= DisplacementStats::kPeriod)
    {
        const float dist_moved = agent.pos_.distanceSquared(ds.init_pos);
        if (dist_moved < DisplacementStats::kDisplamentThreshold * DisplacementStats::kDisplamentThreshold)
            ds.displacement_timed_out = true;
        ds.displacement_ctr = 0;
    }
    else
    {
        ds.displacement_timed_out = false;
        ds.displacement_ctr++;
    }
}

// raySegmentIntersect, CollisionChecker.cu:8-35
bool raySegmentIntersect(float ox, float oy, float dx, float dy, float x1, float y1, float x2, float y2, float range, float *out_t)
{
    const float seg_dx = x2 - x1;
    const float seg_dy = y2 - y1;
    const float denom  = dx * seg_dy - dy * seg_dx;
    if (fabsf(denom) < 1e-8F)
        return false;
    const float t = ((x1 - ox) * seg_dy - (y1 - oy) * seg_dx) / denom;
    const float s = ((x1 - ox) * dy - (y1 - oy) * dx) / denom;
    if ((t >= 0.0F) && (t <= range) && (s >= 0.0F) && (s <= 1.0F))
    {
        *out_t = t;
        return true;
    }
    return false;
}

void castAgent(OkrEnv *e, size_t a)
{
    Agent       *agent = e->agents[a].get();
    const Track &trk   = e->tracks[B<int32_t>(e, OKO_BUF_TRACK_ID)[a]];
    const size_t R     = agent->sensor_ray_angles_.size();
    Ray_        *h     = e->rays.data() + a * R;
    // pack, CollisionChecker.cu:115-128
    for (size_t i = 0; i < R; ++i)
    {
        h[i].x      = agent->pos_.x + agent->sensor_offset_ * cos(kDeg2Rad * agent->rot_);
        h[i].y      = agent->pos_.y + agent->sensor_offset_ * sin(kDeg2Rad * agent->rot_);
        h[i].angle  = kDeg2Rad * (agent->rot_ + agent->sensor_ray_angles_[i]);
        h[i].active = (!agent->crashed_);
    }
    // kernel, CollisionChecker.cu:37-71
    float   *ht   = B<float>(e, OKO_BUF_HIT_T) + a * R;
    int32_t *hseg = B<int32_t>(e, OKO_BUF_HIT_SEG) + a * R;
    for (size_t i = 0; i < R; ++i)
    {
        if (!h[i].active)
            continue;
        const float ray_dx = cosf(h[i].angle);
        const float ray_dy = sinf(h[i].angle);
        float       min_t  = e->cfg.sensor_range;
        int32_t     idx    = -1;
        for (size_t s = 0; s < trk.segments.size(); ++s)
        {
            float       t;
            const auto &sg = trk.segments[s];
            if (raySegmentIntersect(h[i].x, h[i].y, ray_dx, ray_dy, sg.x1, sg.y1, sg.x2, sg.y2, min_t, &t))
            {
                min_t = t;
                idx   = static_cast<int32_t>(s);
            }
        }
        h[i].hit_x = h[i].x + min_t * ray_dx;
        h[i].hit_y = h[i].y + min_t * ray_dy;
        ht[i]      = min_t;
        hseg[i]    = idx;
    }
    // unpack, CollisionChecker.cu:144-172
    float agent_rot_rad = agent->rot_ * kDeg2Rad;
    agent->sensor_hits_.clear();
    float min_dist2{e->cfg.sensor_range * e->cfg.sensor_range};
    for (size_t i = 0; i < R; ++i)
    {
        Vec2d hit_pt_relative;
        float xTranslated = h[i].hit_x - h[i].x;
        float yTranslated = h[i].hit_y - h[i].y;
        hit_pt_relative.x = xTranslated * cos(agent_rot_rad) - yTranslated * sin(agent_rot_rad);
        hit_pt_relative.y = xTranslated * sin(agent_rot_rad) + yTranslated * cos(agent_rot_rad);
        agent->sensor_hits_.push_back(hit_pt_relative);
        if (hit_pt_relative.squaredNorm() < min_dist2)
            min_dist2 = hit_pt_relative.squaredNorm();
    }
    B<float>(e, OKO_BUF_MIN_DIST2)[a] = min_dist2;
    if (min_dist2 < e->cfg.collision_dist2)
        agent->crashed_ = true;
}

void resetOne(OkrEnv *e, size_t a, int32_t reset_idx, bool lane, float alpha, float heading_offset)
{
    // Environment::resetAgent, Environment.cpp:103-121, explicit inputs instead of GetRandomValue
    const Track &trk = e->tracks[B<int32_t>(e, OKO_BUF_TRACK_ID)[a]];
    const auto  &rt  = *trk.rt;
    float        start_pos_x, start_pos_y;
    if (lane)
    {
        auto const nearest_lane_boundary_l = rt.left_bound_inner_[reset_idx];
        auto const nearest_lane_boundary_r = rt.right_bound_inner_[reset_idx];
        start_pos_x = nearest_lane_boundary_l.x * alpha + nearest_lane_boundary_r.x * (1.F - alpha);
        start_pos_y = nearest_lane_boundary_l.y * alpha + nearest_lane_boundary_r.y * (1.F - alpha);
    }
    else
    {
        start_pos_x = rt.track_data_points_.x_m[reset_idx];
        start_pos_y = rt.track_data_points_.y_m[reset_idx];
    }
    e->agents[a]->reset({start_pos_x, start_pos_y}, rt.headings_[reset_idx] + heading_offset);
    B<int32_t>(e, OKO_BUF_RESET_PT)[a]    = reset_idx;
    B<float>(e, OKO_BUF_START_X)[a]       = start_pos_x;
    B<float>(e, OKO_BUF_START_Y)[a]       = start_pos_y;
    B<int32_t>(e, OKO_BUF_PREV_IDX)[a]    = static_cast<int32_t>(rt.findNearestTrackIndexBruteForce({start_pos_x, start_pos_y}));
    B<int32_t>(e, OKO_BUF_NEAREST_IDX)[a] = B<int32_t>(e, OKO_BUF_PREV_IDX)[a];
    B<float>(e, OKO_BUF_FITNESS)[a]       = 0.F;
    B<float>(e, OKO_BUF_REWARD)[a]        = 0.F;
}

// progress / reward / done the apps compute after env.step(), from the real RaceTrack queries
void rewardOne(OkrEnv *e, size_t a)
{
    Agent      *agent   = e->agents[a].get();
    const auto &rt      = *e->tracks[B<int32_t>(e, OKO_BUF_TRACK_ID)[a]].rt;
    float      &reward  = B<float>(e, OKO_BUF_REWARD)[a];
    float      &fitness = B<float>(e, OKO_BUF_FITNESS)[a];
    int32_t    &prev    = B<int32_t>(e, OKO_BUF_PREV_IDX)[a];
    int32_t    &nearest = B<int32_t>(e, OKO_BUF_NEAREST_IDX)[a];
    const int   mode    = e->cfg.reward_mode;
    if (mode == OKO_REWARD_Q_PROGRESS || mode == OKO_REWARD_CMAES_PROGRESS || mode == OKO_REWARD_TRACK_INDEX ||
        mode == OKO_REWARD_LANE_CENTER)
        nearest = static_cast<int32_t>(rt.findNearestTrackIndexBruteForce(agent->pos_));
    switch (mode)
    {
    case OKO_REWARD_Q_PROGRESS: // QAgent.hpp:150-168
    {
        if (agent->crashed_)
        {
            reward = -200.F;
            break;
        }
        int64_t track_idx_len = static_cast<int64_t>(rt.track_data_points_.x_m.size());
        int64_t progression   = static_cast<int64_t>(nearest) - static_cast<int64_t>(prev);
        prev                  = nearest;
        reward = std::abs(progression) > (track_idx_len / 2) ? track_idx_len - std::abs(progression) : std::abs(progression);
        break;
    }
    case OKO_REWARD_CMAES_PROGRESS: // main_eigen.cpp:147-163
    {
        if (!agent->crashed_)
        {
            const int32_t progress{nearest - prev};
            prev = nearest;
            fitness += static_cast<float>(std::abs(progress));
            reward = static_cast<float>(std::abs(progress));
        }
        else
        {
            if (agent->timed_out_)
                fitness = 0.F;
            reward = 0.F;
        }
        break;
    }
    case OKO_REWARD_CONSTANT: reward = 1.0F; break; // ppo_sim.cpp:76
    case OKO_REWARD_DISPLACEMENT:                    // ReinforceContinuous/reinforce_sim.cpp:59-73
    {
        Vec2d prev_pos{B<float>(e, OKO_BUF_START_X)[a], B<float>(e, OKO_BUF_START_Y)[a]};
        reward = (agent->pos_ - prev_pos).norm();
        if (agent->crashed_)
            reward = -5.F;
        break;
    }
    case OKO_REWARD_MIN_RAY: // DQAgent.hpp:161-180
    {
        if (agent->crashed_)
        {
            reward = -200.F;
            break;
        }
        float min_distance = e->cfg.sensor_range;
        for (size_t i{0}; i < agent->sensor_hits_.size(); i++)
            if (min_distance > agent->sensor_hits_[i].norm())
                min_distance = agent->sensor_hits_[i].norm();
        reward = min_distance;
        break;
    }
    case OKO_REWARD_TRACK_INDEX: reward = static_cast<float>(nearest); break; // MiscUtils.hpp:64-71
    case OKO_REWARD_LANE_CENTER:                                                // WorldModelVaeRnn/main.cpp:336-342
    {
        if (!agent->crashed_)
        {
            const float closest_dist = rt.getDistanceToLaneCenter(agent->pos_);
            fitness += (1.F - closest_dist);
            reward = (1.F - closest_dist);
        }
        else
        {
            if (agent->timed_out_)
                fitness = 0.F;
            reward = 0.F;
        }
        break;
    }
    default: reward = 0.F; break;
    }
}

// mirror the object state into the flat buffers the tests read
void publish(OkrEnv *e, size_t a)
{
    Agent *ag                          = e->agents[a].get();
    B<float>(e, OKO_BUF_POS_X)[a]      = ag->pos_.x;
    B<float>(e, OKO_BUF_POS_Y)[a]      = ag->pos_.y;
    B<float>(e, OKO_BUF_ROT)[a]        = ag->rot_;
    B<float>(e, OKO_BUF_SPEED)[a]      = ag->speed_;
    B<float>(e, OKO_BUF_ACCEL)[a]      = ag->acceleration_;
    B<float>(e, OKO_BUF_ACT_THROTTLE)[a] = ag->current_action_.throttle_delta;
    B<float>(e, OKO_BUF_ACT_STEER)[a]    = ag->current_action_.steering_delta;
    B<uint8_t>(e, OKO_BUF_CRASHED)[a]  = ag->crashed_;
    B<uint8_t>(e, OKO_BUF_TIMED_OUT)[a] = ag->timed_out_;
    B<uint8_t>(e, OKO_BUF_DONE)[a]     = ag->isDone();
    B<uint32_t>(e, OKO_BUF_SS_CTR)[a]  = e->stats[a].displacement_ctr;
    B<float>(e, OKO_BUF_SS_X)[a]       = e->stats[a].init_pos.x;
    B<float>(e, OKO_BUF_SS_Y)[a]       = e->stats[a].init_pos.y;
    const size_t R                     = static_cast<size_t>(e->nrays);
    for (size_t i = 0; i < R && i < ag->sensor_hits_.size(); ++i)
    {
        B<float>(e, OKO_BUF_HIT_ABS)[(a * R + i) * 2]     = e->rays[a * R + i].hit_x;
        B<float>(e, OKO_BUF_HIT_ABS)[(a * R + i) * 2 + 1] = e->rays[a * R + i].hit_y;
        B<float>(e, OKO_BUF_HIT_REL)[(a * R + i) * 2]     = ag->sensor_hits_[i].x;
        B<float>(e, OKO_BUF_HIT_REL)[(a * R + i) * 2 + 1] = ag->sensor_hits_[i].y;
        B<float>(e, OKO_BUF_OBS)[a * R + i]               = ag->sensor_hits_[i].norm() / e->cfg.sensor_range;
    }
}
} // namespace

extern "C"
{
void okr_config_default(OkoConfig *c)
{
    oko_config_default(c);
}

OkrEnv *okr_create(const OkoConfig *cfg)
{
    auto *e = new OkrEnv();
    if (cfg)
        e->cfg = *cfg;
    else
        oko_config_default(&e->cfg);
    return e;
}

// constants of a live env; the reference's Agent objects keep dt / speed limit / sensor range as compile-time
// constants (Agent.h:10-12, Agent.cpp:84,110), so only the offset and the movement mode reach them
void okr_update_config(OkrEnv *e, const OkoConfig *cfg)
{
    e->cfg = *cfg;
    for (auto &ag : e->agents)
    {
        ag->sensor_offset_ = e->cfg.sensor_offset;
        ag->setMovementMode(e->cfg.movement_mode == OKO_MOVE_ACCELERATION ? Agent::MovementMode::ACCELERATION
                                                                           : Agent::MovementMode::VELOCITY);
    }
}

void okr_destroy(OkrEnv *e)
{
    delete e;
}

void okr_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

int okr_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int okr_load_track_csv(OkrEnv *e, const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f)
        return -1;
    fclose(f);
    return finishTrack(e, std::make_unique<RaceTrack>(std::string(path)));
}

// raw columns -> temporary CSV (%.9g round-trips binary32 through std::stof) -> the real RaceTrack ctor
int okr_add_track(OkrEnv *e, const float *x, const float *y, const float *wr, const float *wl, int n)
{
    char tmpl[] = "/tmp/okref_track_XXXXXX";
    int  fd     = mkstemp(tmpl);
    if (fd < 0)
        return -1;
    FILE *f = fdopen(fd, "w");
    fprintf(f, "# x_m,y_m,w_tr_right_m,w_tr_left_m\n");
    for (int i = 0; i < n; ++i)
        fprintf(f, "%.9g,%.9g,%.9g,%.9g\n", x[i], y[i], wr[i], wl[i]);
    fclose(f);
    int id = finishTrack(e, std::make_unique<RaceTrack>(std::string(tmpl)));
    unlink(tmpl);
    return id;
}

int okr_num_tracks(const OkrEnv *e)
{
    return static_cast<int>(e->tracks.size());
}
int okr_track_points(const OkrEnv *e, int t)
{
    return static_cast<int>(e->tracks[t].rt->track_data_points_.x_m.size());
}
int okr_track_segments(const OkrEnv *e, int t)
{
    return static_cast<int>(e->tracks[t].segments.size());
}
const float *okr_track_array(const OkrEnv *e, int t, int which)
{
    const Track &k = e->tracks[t];
    switch (which)
    {
    case 0: return k.rt->track_data_points_.x_m.data();
    case 1: return k.rt->track_data_points_.y_m.data();
    case 2: return k.rt->track_data_points_.w_tr_right_m.data();
    case 3: return k.rt->track_data_points_.w_tr_left_m.data();
    case 4: return k.rt->headings_.data();
    case 5: return k.li.data();
    case 6: return k.lo.data();
    case 7: return k.ri.data();
    case 8: return k.ro.data();
    case 9: return reinterpret_cast<const float *>(k.segments.data());
    default: return nullptr;
    }
}

void okr_reset_agents(OkrEnv *e, const int64_t *agent_idx, const int32_t *pt_idx, const float *lane_alpha,
                      const float *heading_off, int64_t n)
{
    for (int64_t k = 0; k < n; ++k)
    {
        const size_t a = static_cast<size_t>(agent_idx ? agent_idx[k] : k);
        resetOne(e, a, pt_idx[k], lane_alpha != nullptr, lane_alpha ? lane_alpha[k] : 0.F, heading_off ? heading_off[k] : 0.F);
        publish(e, a);
    }
}

int okr_alloc_agents(OkrEnv *e, int64_t n, int rays, const float *ray_deg, const int32_t *track_id)
{
    if (n <= 0 || rays <= 0 || e->tracks.empty())
        return -1;
    e->agents.clear();
    e->stats.assign(static_cast<size_t>(n), DisplacementStats{});
    e->nrays = rays;
    e->rays.assign(static_cast<size_t>(n) * rays, Ray_{});
    e->buf.assign(OKO_BUF_COUNT, {});
    for (int i = 0; i < OKO_BUF_COUNT; ++i)
    {
        size_t bytes = static_cast<size_t>(n) * 4;
        if (i == OKO_BUF_CRASHED || i == OKO_BUF_TIMED_OUT || i == OKO_BUF_DONE)
            bytes = static_cast<size_t>(n);
        if (i == OKO_BUF_HIT_ABS || i == OKO_BUF_HIT_REL)
            bytes = static_cast<size_t>(n) * rays * 8;
        if (i == OKO_BUF_OBS || i == OKO_BUF_HIT_SEG || i == OKO_BUF_HIT_T)
            bytes = static_cast<size_t>(n) * rays * 4;
        e->buf[i].assign(bytes + 16, i == OKO_BUF_HIT_SEG ? 0xff : 0);
    }
    for (int64_t i = 0; i < n; ++i)
    {
        auto ag = std::make_unique<HarnessAgent>();
        ag->id_ = static_cast<int16_t>(i);
        ag->sensor_ray_angles_.assign(ray_deg, ray_deg + rays);
        ag->sensor_offset_ = e->cfg.sensor_offset;
        ag->setMovementMode(e->cfg.movement_mode == OKO_MOVE_ACCELERATION ? Agent::MovementMode::ACCELERATION
                                                                           : Agent::MovementMode::VELOCITY);
        ag->sensor_hits_.assign(static_cast<size_t>(rays), Vec2d{});
        e->agents.push_back(std::move(ag));
        B<int32_t>(e, OKO_BUF_TRACK_ID)[i] = track_id ? track_id[i] : 0;
        if (B<int32_t>(e, OKO_BUF_TRACK_ID)[i] < 0 || B<int32_t>(e, OKO_BUF_TRACK_ID)[i] >= static_cast<int>(e->tracks.size()))
            return -1;
    }
    for (int64_t i = 0; i < n; ++i)
    {
        int32_t pt = okr_track_points(e, B<int32_t>(e, OKO_BUF_TRACK_ID)[i]) > 3 ? static_cast<int32_t>(RaceTrack::kStartingIdx) : 0;
        okr_reset_agents(e, &i, &pt, nullptr, nullptr, 1);
    }
    return 0;
}

int64_t okr_num_agents(const OkrEnv *e)
{
    return static_cast<int64_t>(e->agents.size());
}

void *okr_buffer(OkrEnv *e, int which)
{
    return (which >= 0 && which < OKO_BUF_COUNT) ? e->buf[which].data() : nullptr;
}

// the tests may edit the flat state buffers (poses, flags, counters); pull them into the objects
static void absorb(OkrEnv *e, size_t a)
{
    Agent *ag                          = e->agents[a].get();
    ag->pos_.x                         = B<float>(e, OKO_BUF_POS_X)[a];
    ag->pos_.y                         = B<float>(e, OKO_BUF_POS_Y)[a];
    ag->rot_                           = B<float>(e, OKO_BUF_ROT)[a];
    ag->speed_                         = B<float>(e, OKO_BUF_SPEED)[a];
    ag->acceleration_                  = B<float>(e, OKO_BUF_ACCEL)[a];
    ag->current_action_.throttle_delta = B<float>(e, OKO_BUF_ACT_THROTTLE)[a];
    ag->current_action_.steering_delta = B<float>(e, OKO_BUF_ACT_STEER)[a];
    ag->crashed_                       = B<uint8_t>(e, OKO_BUF_CRASHED)[a] != 0;
    ag->timed_out_                     = B<uint8_t>(e, OKO_BUF_TIMED_OUT)[a] != 0;
    e->stats[a].displacement_ctr       = B<uint32_t>(e, OKO_BUF_SS_CTR)[a];
    e->stats[a].init_pos.x             = B<float>(e, OKO_BUF_SS_X)[a];
    e->stats[a].init_pos.y             = B<float>(e, OKO_BUF_SS_Y)[a];
}

void okr_cast_rays(OkrEnv *e)
{
    const int64_t n = static_cast<int64_t>(e->agents.size());
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t a = 0; a < n; ++a)
    {
        absorb(e, a);
        castAgent(e, a);
        publish(e, a);
    }
}

void okr_step(OkrEnv *e, const float *act_throttle, const float *act_steer)
{
    const int64_t n = static_cast<int64_t>(e->agents.size());
    if (act_throttle)
        memcpy(B<float>(e, OKO_BUF_ACT_THROTTLE), act_throttle, sizeof(float) * n);
    if (act_steer)
        memcpy(B<float>(e, OKO_BUF_ACT_STEER), act_steer, sizeof(float) * n);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t a = 0; a < n; ++a)
    {
        absorb(e, a);
        Agent *agent = e->agents[a].get();
        // app-side reset of crashed agents before the tick, GuidedCostLearning/test.cpp:102-111
        if (e->cfg.auto_reset && agent->crashed_)
        {
            const int npts = okr_track_points(e, B<int32_t>(e, OKO_BUF_TRACK_ID)[a]);
            int32_t   pt   = static_cast<int32_t>((static_cast<int64_t>(B<int32_t>(e, OKO_BUF_RESET_PT)[a]) + e->cfg.auto_reset_stride) % npts);
            resetOne(e, a, pt, false, 0.F, 0.F);
        }
        // Environment::step stage 1, Environment.cpp:128-143
        auto &displacement_stats = e->stats[a];
        if (!agent->crashed_)
        {
            agent->move();
            checkAndUpdateStandstill(*agent, displacement_stats);
            if (displacement_stats.displacement_timed_out)
            {
                agent->crashed_   = true;
                agent->timed_out_ = true;
            }
        }
        // stage 2, Environment.cpp:145
        castAgent(e, a);
        rewardOne(e, a);
        publish(e, a);
    }
}

void okr_fill_random_actions(OkrEnv *e, uint64_t step, uint32_t seed)
{
    // same synthetic stream as the oracle (it is an input generator, not reference behaviour)
    static const float kThr[3]   = {-0.3f, 0.0f, 0.3f};
    static const float kSteer[5] = {-4.0f, -1.0f, 0.0f, 1.0f, 4.0f};
    const int64_t      n         = static_cast<int64_t>(e->agents.size());
    for (int64_t a = 0; a < n; ++a)
    {
        const uint64_t id = e->cfg.agent_id_base + (uint64_t)a;
        uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
        uint32_t key[2] = {seed, 0u};
        uint32_t o[4];
        oko_philox4x32_10(ctr, key, o);
        float thr, st;
        if (e->cfg.movement_mode == OKO_MOVE_ACCELERATION)
        {
            thr = kThr[o[0] % 3u];
            st  = kSteer[o[1] % 5u];
        }
        else
        {
            thr = 100.0f * ((float)(o[0] >> 8) * 0x1p-24f);
            st  = 10.0f * ((float)(o[1] >> 8) * 0x1p-24f) - 5.0f;
        }
        B<float>(e, OKO_BUF_ACT_THROTTLE)[a] = thr;
        B<float>(e, OKO_BUF_ACT_STEER)[a]    = st;
    }
}

void okr_sincosf(float x, float *s, float *c)
{
    sincosf(x, s, c); // the libm the reference links
}

int32_t okr_nearest_index(const OkrEnv *e, int track, float x, float y)
{
    return static_cast<int32_t>(e->tracks[track].rt->findNearestTrackIndexBruteForce({x, y}));
}
float okr_dist_lane_center(const OkrEnv *e, int track, float x, float y)
{
    return e->tracks[track].rt->getDistanceToLaneCenter({x, y});
}
float okr_dist_boundary(const OkrEnv *e, int track, float x, float y)
{
    return e->tracks[track].rt->getNearestDistanceToTrackBoundary({x, y});
}
float okr_normalize_angle_deg(float a)
{
    return normalizeAngleDeg(a);
}
void okr_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    oko_philox4x32_10(ctr, key, out);
}
}
