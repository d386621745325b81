/*
 * openkitchen_b200.h -- C ABI of the B200-native batched step loop for OpenKitchen's racing
 * environment.  This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 *
 * The reference has no FFI layer; its seam is the C++ class surface of Environment/ and the
 * pybind module.  Each entry point below names the reference interface it replaces (file:line
 * relative to the reference repo root).  The C++ shim (openkitchen_b200/shim/) and the pybind
 * module (openkitchen_b200/pybind/) are written on top of exactly these functions; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (OK_SUCCESS) or a negative OkStatus; nothing throws or aborts
 *     across the ABI; ok_last_error() returns a thread-local message for the last failure;
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = the legacy default stream);
 *     launches are asynchronous on that stream and never synchronise the device;
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - agent buffers are owned by the OkEnv and stay valid until ok_alloc_agents is called again
 *     or the env is destroyed (borrowed by callers -- this is what DLPack exports);
 *   - one OkEnv per GPU; an env is thread-compatible (external synchronisation per env).
 */
#ifndef OPENKITCHEN_B200_H
#define OPENKITCHEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OK_ABI_VERSION 2

typedef enum OkStatus {
    OK_SUCCESS            = 0,
    OK_ERR_INVALID_ARG    = -1,
    OK_ERR_CUDA           = -2,
    OK_ERR_IO             = -3,
    OK_ERR_STATE          = -4, /* e.g. stepping before ok_alloc_agents */
    OK_ERR_CAPACITY       = -5, /* a track does not fit the shared-memory staging area */
    OK_ERR_NO_DEVICE      = -6
} OkStatus;

/* Agent::MovementMode, Environment/Agent.h:20-25 (MANUAL is keyboard input: not supported) */
typedef enum OkMovementMode { OK_MOVE_VELOCITY = 0, OK_MOVE_ACCELERATION = 1 } OkMovementMode;

/* App-side progress/reward definitions */
typedef enum OkRewardMode {
    OK_REWARD_NONE           = 0,
    OK_REWARD_Q_PROGRESS     = 1, /* RLRacers/Q_Learning/QAgent.hpp:150-168 */
    OK_REWARD_CMAES_PROGRESS = 2, /* CovarianceMatrixAdaptationEvolution/main_eigen.cpp:147-163 */
    OK_REWARD_CONSTANT       = 3, /* RLRacers/PPO/ppo_sim.cpp:76 */
    OK_REWARD_DISPLACEMENT   = 4, /* RLRacers/ReinforceContinuous/reinforce_sim.cpp:59-73 */
    OK_REWARD_MIN_RAY        = 5, /* RLRacers/Deep_Q_Learning/DQAgent.hpp:161-180 */
    OK_REWARD_TRACK_INDEX    = 6, /* EvolutionaryRacer/MiscUtils.hpp:64-71 */
    OK_REWARD_LANE_CENTER    = 7  /* WorldModelVaeRnn/main.cpp:336-342 */
} OkRewardMode;

typedef enum OkRaycastMode {
    OK_RAYCAST_GRID  = 0, /* uniform-grid broadphase: DDA walk over 8 px cells */
    OK_RAYCAST_BRUTE = 1, /* every ray x every segment, the reference's loop (CollisionChecker.cu:51-66) */
    OK_RAYCAST_BEAM  = 2  /* precomputed (start cell, direction bin) candidate lists, grid walk only for the rays a
                             list cannot decide; same results as the two modes above */
} OkRaycastMode;

/* Device buffers exported by ok_get_buffer.  N = agents, R = rays per agent. */
typedef enum OkBuffer {
    OK_BUF_POS_X = 0,    /* f32[N]     Agent::pos_.x                         Agent.h:56 */
    OK_BUF_POS_Y,        /* f32[N]     Agent::pos_.y */
    OK_BUF_ROT,          /* f32[N]     Agent::rot_ (degrees, never wrapped)  Agent.h:59 */
    OK_BUF_SPEED,        /* f32[N]     Agent::speed_                         Agent.h:57 */
    OK_BUF_ACCEL,        /* f32[N]     Agent::acceleration_                  Agent.h:58 */
    OK_BUF_ACT_THROTTLE, /* f32[N]     Agent::current_action_.throttle_delta Agent.h:78 */
    OK_BUF_ACT_STEER,    /* f32[N]     Agent::current_action_.steering_delta */
    OK_BUF_CRASHED,      /* u8[N]      Agent::crashed_                       Agent.h:72 */
    OK_BUF_TIMED_OUT,    /* u8[N]      Agent::timed_out_                     Agent.h:74 */
    OK_BUF_DONE,         /* u8[N]      Agent::isDone()                       Agent.cpp:138-144 */
    OK_BUF_SS_CTR,       /* u32[N]     DisplacementStats::displacement_ctr   Environment.h:23 */
    OK_BUF_SS_X,         /* f32[N]     DisplacementStats::init_pos.x         Environment.h:25 */
    OK_BUF_SS_Y,         /* f32[N]     DisplacementStats::init_pos.y */
    OK_BUF_TRACK_ID,     /* i32[N]     which track the agent drives on */
    OK_BUF_HIT_ABS,      /* f32[N,R,2] Ray_::hit_x/hit_y (window coords)     Typedefs.h:96-97 */
    OK_BUF_HIT_REL,      /* f32[N,R,2] Agent::sensor_hits_                   Agent.h:76 */
    OK_BUF_OBS,          /* f32[N,R]   sensor_hits_[i].norm()/sensor_range   PPOAgent.hpp:68-76 */
    OK_BUF_HIT_SEG,      /* i32[N,R]   index of the hit segment in TrackSegments order, -1 = none */
    OK_BUF_HIT_T,        /* f32[N,R]   ray parameter of the hit (min_t of CollisionChecker.cu:49) */
    OK_BUF_MIN_DIST2,    /* f32[N]     min_dist2 of CollisionChecker.cu:150 */
    OK_BUF_NEAREST_IDX,  /* i32[N]     RaceTrack::findNearestTrackIndexBruteForce(pos_) */
    OK_BUF_PREV_IDX,     /* i32[N]     prev_track_idx_ of the progress rewards */
    OK_BUF_REWARD,       /* f32[N]     per-tick reward of OkConfig::reward_mode */
    OK_BUF_FITNESS,      /* f32[N]     accumulated fitness (CMA-ES / lane-centre modes) */
    OK_BUF_RESET_PT,     /* i32[N]     centre-line index of the agent's last reset */
    OK_BUF_START_X,      /* f32[N]     pose at the start of the episode */
    OK_BUF_START_Y,      /* f32[N] */
    OK_BUF_COUNT
} OkBuffer;

typedef enum OkDType { OK_DTYPE_F32 = 0, OK_DTYPE_I32 = 1, OK_DTYPE_U32 = 2, OK_DTYPE_U8 = 3 } OkDType;

/* Track geometry arrays readable through ok_track_copy (host side, reference layout). */
typedef enum OkTrackArray {
    OK_TRACK_X = 0,     /* f32[P]   RaceTrack::track_data_points_.x_m (after fit-to-window) */
    OK_TRACK_Y,         /* f32[P] */
    OK_TRACK_W_RIGHT,   /* f32[P]   w_tr_right_m (clamped, scaled) */
    OK_TRACK_W_LEFT,    /* f32[P] */
    OK_TRACK_HEADING,   /* f32[P]   RaceTrack::headings_ (degrees) */
    OK_TRACK_LEFT_INNER,  /* f32[P,2] RaceTrack::left_bound_inner_ */
    OK_TRACK_LEFT_OUTER,  /* f32[P,2] */
    OK_TRACK_RIGHT_INNER, /* f32[P,2] */
    OK_TRACK_RIGHT_OUTER, /* f32[P,2] */
    OK_TRACK_SEGMENTS     /* f32[S,4] Segment2d in TrackSegments order, S = 4(P-1)+4 */
} OkTrackArray;

/* Constants the reference bakes in at compile time; defaults = reference values. */
typedef struct OkConfig {
    int32_t  device;               /* CUDA device ordinal */
    int32_t  movement_mode;        /* OkMovementMode; Agent.h:80 default VELOCITY */
    int32_t  reward_mode;          /* OkRewardMode */
    int32_t  raycast_mode;         /* OkRaycastMode */
    int32_t  auto_reset;           /* !=0: a crashed agent is reset at the start of the next step
                                      (GuidedCostLearning/test.cpp:102-111) to centre-line index
                                      (last_reset + auto_reset_stride) mod P, with a zero action */
    int32_t  auto_reset_stride;    /* 97 */
    float    sensor_range;         /* Agent::kSensorRange 200        Agent.h:10 */
    float    speed_limit;          /* Agent::kSpeedLimit 100         Agent.h:11 */
    float    dt;                   /* kDt 0.016                      Agent.cpp:84,110 */
    float    collision_dist2;      /* 2.0 (strict <)                 CollisionChecker.cu:167 */
    float    sensor_offset;        /* Agent::sensor_offset_ 0        Agent.h:61 */
    uint32_t standstill_period;    /* DisplacementStats::kPeriod 200 Environment.h:19 */
    float    standstill_threshold; /* kDisplamentThreshold 20        Environment.h:20 */
    float    grid_cell;            /* broadphase cell size in px (8) */
    float    beam_cell;            /* OK_RAYCAST_BEAM: start-cell size in px (2); 0 = default */
    int32_t  beam_bins;            /* OK_RAYCAST_BEAM: direction bins, a power of two (256); 0 = default */
    uint64_t agent_id_base;        /* global id of this env's agent 0 (multi-GPU: the first agent of the rank's slice).
                                      The synthetic Philox action stream is keyed by (agent_id_base + a, step), so a
                                      shard reproduces, bit for bit, its slice of the unsharded population */
} OkConfig;

typedef struct OkTrackInfo {
    int32_t n_points;
    int32_t n_segments;
    int32_t grid_nx, grid_ny;
    int32_t grid_items;   /* total (cell, segment) registrations */
    int32_t blob_bytes;   /* bytes staged into shared memory for this track */
    float   grid_x0, grid_y0, grid_cell;
} OkTrackInfo;

typedef struct OkEnv OkEnv;

/* ---- lifetime ------------------------------------------------------------------------------ */
int         ok_abi_version(void);
const char *ok_last_error(void);
void        ok_config_default(OkConfig *cfg);
/* Environment::Environment (Environment.cpp:42-62) minus raylib/visualizer/screen grabber. */
int  ok_create(const OkConfig *cfg, OkEnv **out);
void ok_destroy(OkEnv *env);
/* change the per-tick constants of a live env (everything except `device`, `grid_cell` and `beam_*`, which are fixed
 * at creation): e.g. Agent::setMovementMode (Agent.h:46-49), the reward definition, auto-reset */
int  ok_update_config(OkEnv *env, const OkConfig *cfg);
int  ok_get_config(const OkEnv *env, OkConfig *out);

/* ---- tracks: RaceTrack ctor chain (RaceTrack.cpp:3-14,127-307) + TrackSegments (TrackSegments.cu:6-76)
 *      + broadphase grid build + device upload. Must precede ok_alloc_agents. -------------------- */
/* the four raw CSV columns x_m,y_m,w_tr_right_m,w_tr_left_m (before clamp/scale) */
int ok_add_track(OkEnv *env, const float *h_x_m, const float *h_y_m, const float *h_w_right, const float *h_w_left,
                 int32_t n_points, int32_t *track_id_out);
/* RaceTrack::getTrackDataFromCsv, RaceTrack.cpp:127-164 */
int ok_load_track_csv(OkEnv *env, const char *path, int32_t *track_id_out);
int ok_num_tracks(const OkEnv *env);
int ok_track_info(const OkEnv *env, int32_t track_id, OkTrackInfo *out);
/* copies min(count, available) elements of the named array to h_out; returns elements available */
int64_t ok_track_copy(const OkEnv *env, int32_t track_id, int32_t which /*OkTrackArray*/, float *h_out, int64_t count);

/* ---- agents -------------------------------------------------------------------------------- */
/* replaces `std::vector<Agent*>` + CollisionChecker ctor (CollisionChecker.cu:76-86).  All agents
 * have `rays` rays (the reference assumes agents[0]'s count, CollisionChecker.cu:82) with angles
 * h_ray_deg[rays] relative to heading (Agent::sensor_ray_angles_).  h_track_id may be NULL (all on
 * track 0).  Every agent starts reset at RaceTrack::kStartingIdx (RaceTrack.h:18). Agents on the
 * same track should be contiguous for best performance. */
int     ok_alloc_agents(OkEnv *env, int64_t n_agents, int32_t rays, const float *h_ray_deg, const int32_t *h_track_id);
int64_t ok_num_agents(const OkEnv *env);
int32_t ok_num_rays(const OkEnv *env);

/* Environment::resetAgent (Environment.cpp:79-122) + Agent::reset (Agent.cpp:123-135) with the
 * random draws made explicit: agent d_agent_idx[k] (NULL = k) goes to centre-line index d_pt_idx[k];
 * d_lane_alpha (nullable) = the lane lerp alpha of Environment.cpp:107-114; d_heading_off (nullable)
 * = the heading offset of Environment.cpp:88-101.  Device pointers, asynchronous on `stream`. */
int ok_reset_agents(OkEnv *env, const int64_t *d_agent_idx, const int32_t *d_pt_idx, const float *d_lane_alpha,
                    const float *d_heading_off, int64_t n, void *stream);
/* same with HOST arrays (staged through an internal buffer; synchronises `stream`) */
int ok_reset_agents_host(OkEnv *env, const int64_t *h_agent_idx, const int32_t *h_pt_idx, const float *h_lane_alpha,
                         const float *h_heading_off, int64_t n, void *stream);

/* ---- the tick ------------------------------------------------------------------------------ */
/* CollisionChecker::checkCollision (CollisionChecker.cu:96-100,113-174): lidar + crash flag on the
 * current poses, no movement. */
int ok_cast_rays(OkEnv *env, void *stream);
/* Environment::step stages 1-2 (Environment.cpp:125-146) + the app-side progress/reward/done:
 * one fused kernel.  d_act_* (nullable) are copied into the ACT_* buffers first; NULL = step with
 * the actions already stored there (what Agent::updateAction wrote). */
int ok_launch_step(OkEnv *env, const float *d_act_throttle, const float *d_act_steer, void *stream);
/* k consecutive ticks whose actions come from the counter-based Philox4x32-10 stream of
 * SURVEY.md 8(d) (key = seed, counter = (agent, first_step + i)), generated inside the step kernel. */
int ok_launch_steps_random(OkEnv *env, uint64_t first_step, int32_t k, uint32_t seed, void *stream);
/* write the Philox actions for `step` into the ACT_* buffers without stepping */
int ok_fill_random_actions(OkEnv *env, uint64_t step, uint32_t seed, void *stream);

/* Track queries for arbitrary points (the app-side reward code of WorldModelVaeRnn/main.cpp:336-337 and every
 * progress reward): RaceTrack::findNearestTrackIndexBruteForce (RaceTrack.cpp:16-31), getDistanceToLaneCenter
 * (RaceTrack.cpp:53-72) and getNearestDistanceToTrackBoundary (RaceTrack.cpp:33-51) for n points.  d_x, d_y f32[n];
 * d_track_id i32[n] or NULL (track 0); each output is nullable.  One warp per query; asynchronous on `stream`. */
int ok_track_query(OkEnv *env, const float *d_x, const float *d_y, const int32_t *d_track_id, int64_t n,
                   int32_t *d_nearest_idx, float *d_dist_lane_center, float *d_dist_boundary, void *stream);
/* same with HOST arrays (staged; synchronises `stream`) */
int ok_track_query_host(OkEnv *env, const float *h_x, const float *h_y, const int32_t *h_track_id, int64_t n,
                        int32_t *h_nearest_idx, float *h_dist_lane_center, float *h_dist_boundary, void *stream);

/* The EvolutionaryRacer policy for the whole population: GeneticAgent::updateAction +
 * Network::infer (EvolutionaryRacer/GeneticAgent.hpp:37-50, Network.hpp:119-155).  Agent i has its own weights
 * d_w1[i] (f32 [R+2][hidden], inputs speed/100, normalizeAngleDeg(rot)/360, hits/200) and d_w2[i] (f32 [hidden][6]);
 * hidden = relu(x W1), out = sigmoid(hidden W2), actions decoded with the 0.5 activation limit into the ACT_*
 * buffers.  hidden <= 32.  One warp per agent; the kernel streams the weights once (HBM bound: 4*(R+2+6)*hidden
 * bytes per agent). */
int ok_genetic_policy(OkEnv *env, const float *d_w1, const float *d_w2, int32_t hidden, void *stream);

/* The CMA-ES racer's policy for the whole population: CmaEsAgent::updateAction + Controller::forward
 * (CovarianceMatrixAdaptationEvolution/main_torch.cpp:61-71, Controller.cpp:16-23).  Candidate i evaluates
 * tanh(fc3 tanh(fc2 tanh(fc1 obs_i))) with its OWN flat parameter vector d_params[i] (f32[n_params], torch parameters()
 * order: fc1.weight [16][R], fc1.bias, fc2.weight [8][16], fc2.bias, fc3.weight [1][8], fc3.bias; n_params = 16 R + 169)
 * and writes ACT_THROTTLE = throttle (100), ACT_STEER = steer_scale (5) * output.  hidden must be 16 (main_torch.cpp:19).
 * One warp per agent, weights streamed once: HBM bound at 4 * n_params bytes per agent. */
int ok_cmaes_controller(OkEnv *env, const float *d_params, int32_t n_params, int32_t hidden, float throttle, float steer_scale,
                        void *stream);

/* The PPO racers' policy step for every agent in ONE kernel: PPOAgent::updateAction + Actor::forward
 * (RLRacers/PPO/PPOAgent.hpp:79-102, Actor.hpp:9-26): probs = softmax(W2 relu(W1 obs + b1) + b2) clamped to
 * [1e-8, 1 - 1e-8]; an action index drawn with torch::multinomial's distribution (inverse CDF on one uniform per agent:
 * d_uniform[N] if given, else Philox keyed by (agent_id_base + a, step) with `seed`), or the argmax when greedy != 0;
 * log-probability of the choice; (throttle, steering) = d_action_table[action] into the ACT_* buffers
 * (PPOAgent::kActionMap).  Weights are device pointers in torch.nn.Linear layout, shared by all agents: d_w1 [hidden][R],
 * d_b1 [hidden], d_w2 [n_actions][hidden], d_b2 [n_actions]; n_actions <= 8.  Nullable outputs (rows of a rollout
 * buffer): d_action i32[N], d_log_prob f32[N], d_probs f32[N][n_actions], d_obs f32[N][R] (the observation the action
 * was chosen from), d_prev_reward f32[N] / d_prev_done u8[N] (REWARD / DONE of the tick before this call).
 * With d_w1 == NULL only the d_prev_* record is made (the flush after a rollout's last tick). */
typedef struct OkActorIO {
    const float *d_w1, *d_b1, *d_w2, *d_b2;
    int32_t      hidden, n_actions;
    const float *d_action_table;
    const float *d_uniform;
    int32_t      greedy, reserved;
    int32_t     *d_action;
    float       *d_log_prob, *d_probs, *d_obs;
    float       *d_prev_reward;
    uint8_t     *d_prev_done;
} OkActorIO;
int ok_ppo_actor(OkEnv *env, const OkActorIO *io, uint64_t step, uint32_t seed, void *stream);
/* ok_ppo_actor and the tick that consumes its actions (ok_launch_step with the stored actions) as ONE launch: the policy
 * step runs as phase 0 of every tile of the step kernel (ppo_sim.cpp:61-89's updateAction + env.step per tick).  At a few
 * thousand agents a tick is launch-to-completion latency, and two kernels pay it twice.  Same results as the two calls; falls
 * back to them when the env's kernel shape has no fused instantiation (grid / brute-force modes; the staged shapes of
 * populations of 65,536 rays per tick and more, whose tick is not latency). */
int ok_ppo_actor_step(OkEnv *env, const OkActorIO *io, uint64_t step, uint32_t seed, void *stream);
/* ExperienceBuffer::calculateDiscountedRewards (RLRacers/PPO/ExperienceBuffer.hpp:45-62) for all agents: d_rewards
 * f32[steps][N] -> d_out f32[steps][N], ret[t] = r[t] + gamma * ret[t + 1] in binary32 with the reference's operation
 * order; d_done u8[steps][N] (nullable) restarts the sum where an episode ended.  No env needed beyond its device. */
int ok_discounted_returns(OkEnv *env, const float *d_rewards, const uint8_t *d_done, float *d_out, int32_t steps, int64_t n,
                          float gamma, void *stream);

/* End-to-end host call (what the C++ shim's Environment::step and the reference-facing plugin
 * use): H2D of the two action arrays, one tick, D2H of obs / reward / done / crashed (each
 * nullable), then a stream synchronise.  Host buffers should be pinned (ok_host_alloc). */
int ok_step_host(OkEnv *env, const float *h_act_throttle, const float *h_act_steer, float *h_obs, float *h_reward,
                 uint8_t *h_done, void *stream);
/* ok_step_host with the observations as 16-bit fixed point -- OPT-IN AND LOSSY: h_obs_q16[a][r] =
 * round-to-nearest-even(clamp(obs, 0, 1) * 65535), obs being the binary32 value ok_step_host delivers
 * (|error| <= 0.5 / 65535 = 7.7e-6 of the sensor range, 0.0015 px at 200 px).  Half the bytes on the host link, which bounds the
 * end-to-end tick.  The device-side buffers (OK_BUF_OBS ...) keep the exact values; h_obs_q16 must be
 * pinned.  No counterpart in the reference (its consumers read Agent::sensor_hits_, Agent.h:66, as float). */
int ok_step_host_q16(OkEnv *env, const float *h_act_throttle, const float *h_act_steer, uint16_t *h_obs_q16, float *h_reward,
                     uint8_t *h_done, void *stream);
/* The object-model tick in ONE upload, one launch and two downloads (what the C++ shim's Environment::step and
 * CollisionChecker::checkCollision use; round 1 made ~24 synchronous buffer copies per tick).  The env's first thirteen
 * buffers (OK_BUF_POS_X .. OK_BUF_SS_Y: pose, speed, acceleration, action, flags, standstill record) are adjacent in
 * device memory; h_state is a byte-for-byte host image of that block (layout from ok_packed_layout).  The call copies
 * h_state to the device, runs one tick (move != 0: ok_launch_step with the actions in the block; 0: ok_cast_rays), copies
 * the block back into h_state and HIT_ABS + HIT_REL into h_hits, and synchronises `stream` once.  Pinned buffers
 * (ok_host_alloc) make the three copies asynchronous. */
typedef struct OkPackedLayout {
    size_t state_bytes;      /* bytes of the state block */
    size_t state_offset[13]; /* byte offset of OK_BUF_POS_X ... OK_BUF_SS_Y (enum order) inside it */
    size_t hits_bytes;       /* bytes of the hits block */
    size_t hit_abs_offset, hit_rel_offset;
} OkPackedLayout;
int ok_packed_layout(const OkEnv *env, OkPackedLayout *out);
int ok_step_packed(OkEnv *env, void *h_state, void *h_hits, int32_t move, void *stream);
int ok_host_alloc(void **h_ptr, size_t bytes); /* cudaMallocHost */
int ok_host_free(void *h_ptr);

/* ---- buffers ------------------------------------------------------------------------------- */
/* device pointer, shape (shape[1] = 0 for per-agent vectors; shape[2] = 2 for the xy buffers) and
 * OkDType of an agent buffer: the DLPack export of the Python layer is built from this. */
int ok_get_buffer(OkEnv *env, int32_t which /*OkBuffer*/, void **d_ptr, int64_t shape[3], int32_t *dtype);
/* convenience copies (synchronise `stream`) */
int ok_read_buffer(OkEnv *env, int32_t which, void *h_dst, size_t bytes, void *stream);
int ok_write_buffer(OkEnv *env, int32_t which, const void *h_src, size_t bytes, void *stream);
int ok_sync(OkEnv *env, void *stream);

/* ---- introspection (bench / profiling) ----------------------------------------------------- */
/* Population counters (metrics: what the reference's apps print per episode -- colony averages, how many
 * agents are still driving -- needs these): agents alive (= not crashed) / crashed / timed out / done in the env's buffers as
 * they are now, counted on the device.  Synchronises `stream`. */
typedef struct OkPopulationCounters {
    uint64_t agents, alive, crashed, timed_out, done;
} OkPopulationCounters;
int ok_population_counters(OkEnv *env, OkPopulationCounters *out, void *stream);
typedef struct OkLaunchStats {
    uint64_t kernel_launches; /* kernels launched by this env since creation */
    int32_t  grid_blocks, block_threads, smem_bytes, tiles;
} OkLaunchStats;
int ok_launch_stats(const OkEnv *env, OkLaunchStats *out);
/* Work counters of the beam kernel (profiling aid, off by default): out[0] = rays cast, out[1] = rays its first pass
 * could not settle from the table entry alone (queued for the second pass), out[2] = rays that fell back to the
 * uniform-grid walk, since the previous call.  Synchronises the device; enable != 0 keeps counting, 0 stops. */
int ok_debug_stats(OkEnv *env, uint64_t out[4], int32_t enable);
/* Index self-checks: a library built with -DOK_CHECKED=1 (tools/checked_build.sh) counts every table row, chunk,
 * segment, ray or agent index that would leave its array; *count is the total since the library was loaded (always 0 for
 * the product build, whose *checks_compiled_in is 0).  Synchronises the device. */
int ok_debug_violations(OkEnv *env, uint64_t *count, int32_t *checks_compiled_in);
/* Timeline of the beam kernel's last launch (profiling aid, off by default): per CTA and per tile it processed, six
 * 64-bit words {tile | smid << 32, then %globaltimer (ns) at: tile start, end of the thread-per-agent phase, end of the
 * first ray pass, end of the second ray pass, end of the reward phase}.  Copies min(words available, capacity_words) to
 * h_out (layout [grid_blocks][tiles_per_cta][6], zero = unused) and returns the count; then re-arms the trace with room
 * for `tiles_per_cta` tiles per CTA (0 = off).  Synchronises the device. */
int64_t ok_debug_trace(OkEnv *env, uint64_t *h_out, int64_t capacity_words, int32_t tiles_per_cta);
/* Feedback tiling of the staged beam kernels.  The kernel leaves every tile's time (ray phase / rest) in a small device
 * buffer; this call synchronises `stream`, reads it and re-cuts the tiles so that a tile's MEASURED cost -- not its agent
 * count -- is equal across the SMs (tracks differ by up to 15 % per agent).  Results never depend on the tiling.  The
 * library does this on its own after the 16th, 64th and 256th launch of an env and every 4,096 launches from then on
 * (never inside a stream capture; OK_AUTO_BALANCE=0 turns that off); call it yourself after a change of workload.
 * out[0] / out[1] (nullable): slowest tile over mean tile before (measured) and after (predicted). */
int ok_balance_schedule(OkEnv *env, void *stream, float out[2]);
/* The beam kernel's tiling of the population (profiling aid; ok_debug_trace's tile ids index it): copies up to `capacity`
 * tiles as {track, count, first agent} to h_out[capacity][3] and returns the number of tiles. */
int64_t ok_debug_tiles(OkEnv *env, int64_t *h_out, int64_t capacity);
/* evaluates the kernels' sincosf (the glibc-2.39 restatement, ok_math.cuh) on `n` host floats: lets a test
 * compare the DEVICE function with libm / the oracle directly (tests/test_gpu_math.py) */
int ok_eval_sincosf(OkEnv *env, const float *h_in, float *h_sin, float *h_cos, int64_t n);

/* Host-link probe (measurement only, no env needed): moves `bytes` between a pinned host buffer and device memory
 * `iters` times on one stream and returns the achieved GB/s in *gbps_out.  mode 0: device->host DMA copies; 1: a kernel
 * storing through the host mapping (what ok_step_host does with the lidar observations); 2: host->device DMA copies.
 * bench.py and tools/pcie_ceiling.py run it on all ranks at once to measure the box's ceiling for the end-to-end path. */
int ok_pcie_probe(int32_t device, size_t bytes, int32_t iters, int32_t mode, double *gbps_out);

/* OK_RAYCAST_BEAM's candidate table, host side (no device needed; built on first use, ok_beam.hpp): the
 * segments a ray starting at (x, y) with direction `angle_rad` can reach within *d_complete px, nearest first.
 * Copies min(count, capacity) indices to h_items and returns count; OK_BEAM_NOT_COVERED = the start cell is not
 * covered (the kernel then uses the grid walk); other negative values are OkStatus errors.  Test / inspection hook. */
#define OK_BEAM_NOT_COVERED (-100)
int32_t ok_beam_lookup(OkEnv *env, int32_t track_id, float x, float y, float angle_rad, uint16_t *h_items,
                       int32_t capacity, float *d_complete);
/* the same, also returning what the kernel's first pass works with: *n_inline = how many of the leading candidates are
 * carried in the table entry itself (at most 4), *d_inline = the distance up to which those are provably the ONLY
 * contenders (every other listed segment is met at a parameter >= *d_inline by every ray of the cell and bin) */
int32_t ok_beam_lookup_ex(OkEnv *env, int32_t track_id, float x, float y, float angle_rad, uint16_t *h_items,
                          int32_t capacity, float *d_complete, float *d_inline, int32_t *n_inline);
/* size of the track's beam table in bytes (builds it if needed) */
int64_t ok_beam_table_bytes(OkEnv *env, int32_t track_id);
/* Beam tables are shared between the envs of a process through a cache (host copies and, per device, the tables the
 * kernels read: ~250 MB per track at the default resolution).  This drops the cache's own references; a table is freed
 * when the last env that uses it is destroyed. */
void ok_release_caches(void);

#ifdef __cplusplus
}
#endif
#endif /* OPENKITCHEN_B200_H */
