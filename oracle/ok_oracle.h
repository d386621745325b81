/*
 * ok_oracle.h -- CPU restatement of OpenKitchen's per-tick hot path (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle, not the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (openkitchen_b200/) never includes, links or calls anything in this directory.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Parity pin: oracle/ref_harness.cpp links the UNMODIFIED reference Agent.cpp and
 * RaceTrack.cpp (oracle/_ref/libokref.so) and tests/test_oracle_vs_ref.py checks this
 * restatement against it bit-for-bit; tests/golden/ holds vectors minted from that build.
 */
#ifndef OK_ORACLE_H
#define OK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Buffer ids: identical numbering to include/openkitchen_b200.h (OkBuffer). */
enum {
    OKO_BUF_POS_X = 0, OKO_BUF_POS_Y, OKO_BUF_ROT, OKO_BUF_SPEED, OKO_BUF_ACCEL,
    OKO_BUF_ACT_THROTTLE, OKO_BUF_ACT_STEER,
    OKO_BUF_CRASHED, OKO_BUF_TIMED_OUT, OKO_BUF_DONE,
    OKO_BUF_SS_CTR, OKO_BUF_SS_X, OKO_BUF_SS_Y,
    OKO_BUF_TRACK_ID,
    OKO_BUF_HIT_ABS, OKO_BUF_HIT_REL, OKO_BUF_OBS, OKO_BUF_HIT_SEG, OKO_BUF_HIT_T,
    OKO_BUF_MIN_DIST2,
    OKO_BUF_NEAREST_IDX, OKO_BUF_PREV_IDX, OKO_BUF_REWARD, OKO_BUF_FITNESS,
    OKO_BUF_RESET_PT, OKO_BUF_START_X, OKO_BUF_START_Y,
    OKO_BUF_COUNT
};

enum { OKO_MOVE_VELOCITY = 0, OKO_MOVE_ACCELERATION = 1 };

enum {
    OKO_REWARD_NONE = 0,
    OKO_REWARD_Q_PROGRESS = 1,     /* QAgent.hpp:150-168 */
    OKO_REWARD_CMAES_PROGRESS = 2, /* main_eigen.cpp:147-163 */
    OKO_REWARD_CONSTANT = 3,       /* ppo_sim.cpp:76 */
    OKO_REWARD_DISPLACEMENT = 4,   /* ReinforceContinuous/reinforce_sim.cpp:59-73 */
    OKO_REWARD_MIN_RAY = 5,        /* DQAgent.hpp:161-180 */
    OKO_REWARD_TRACK_INDEX = 6,    /* MiscUtils.hpp:64-71 */
    OKO_REWARD_LANE_CENTER = 7     /* WorldModelVaeRnn/main.cpp:336-342 */
};

typedef struct OkoConfig {
    int32_t  movement_mode;        /* Agent.h:20-25 */
    int32_t  reward_mode;
    int32_t  auto_reset;           /* GuidedCostLearning/test.cpp:102-111 style reset of crashed agents */
    int32_t  auto_reset_stride;    /* next pt_idx = (prev + stride) mod pts */
    float    sensor_range;         /* Agent.h:10  (200) */
    float    speed_limit;          /* Agent.h:11  (100) */
    float    dt;                   /* Agent.cpp:84,110 (0.016) */
    float    collision_dist2;      /* CollisionChecker.cu:167 (2.0) */
    float    sensor_offset;        /* Agent.h:61 (0) */
    uint32_t standstill_period;    /* Environment.h:19 (200) */
    float    standstill_threshold; /* Environment.h:20 (20) */
    uint32_t pad_;
    uint64_t agent_id_base;        /* global id of agent 0: the Philox counter of the synthetic action stream is
                                      (agent_id_base + a, step), so a shard reproduces its slice of an unsharded run */
} OkoConfig;

typedef struct OkoEnv OkoEnv;

void    oko_config_default(OkoConfig *cfg);
OkoEnv *oko_create(const OkoConfig *cfg);
void    oko_destroy(OkoEnv *env);
void    oko_update_config(OkoEnv *env, const OkoConfig *cfg); /* constants of a live env */
void    oko_set_threads(int n); /* OpenMP threads for the agent loop (1 = the reference's single thread) */
int     oko_get_max_threads(void);

/* RaceTrack ctor chain (RaceTrack.cpp:3-14) on the four raw CSV columns. Returns track id. */
int oko_add_track(OkoEnv *env, const float *x_m, const float *y_m, const float *w_right, const float *w_left, int n);
/* getTrackDataFromCsv (RaceTrack.cpp:127-164) + the above. Returns track id or -1. */
int oko_load_track_csv(OkoEnv *env, const char *path);
int oko_num_tracks(const OkoEnv *env);
int oko_track_points(const OkoEnv *env, int track);
int oko_track_segments(const OkoEnv *env, int track);
/* which: 0 x, 1 y, 2 w_right, 3 w_left, 4 heading, 5 LI(xy), 6 LO(xy), 7 RI(xy), 8 RO(xy), 9 segments(x1,y1,x2,y2) */
const float *oko_track_array(const OkoEnv *env, int track, int which);

int     oko_alloc_agents(OkoEnv *env, int64_t n, int rays, const float *ray_deg, const int32_t *track_id);
int64_t oko_num_agents(const OkoEnv *env);
void   *oko_buffer(OkoEnv *env, int which); /* host pointer, layout as the product buffer of the same id */

/* Environment::resetAgent (Environment.cpp:79-122) with explicit idx / alpha / heading offset
 * instead of raylib GetRandomValue. lane_alpha / heading_off may be NULL. */
void oko_reset_agents(OkoEnv *env, const int64_t *agent_idx, const int32_t *pt_idx, const float *lane_alpha,
                      const float *heading_off, int64_t n);

/* CollisionChecker::checkCollision (CollisionChecker.cu:113-172) on the current poses. */
void oko_cast_rays(OkoEnv *env);
/* Environment::step stages 1-2 (Environment.cpp:125-146); actions may be NULL (use stored). */
void oko_step(OkoEnv *env, const float *act_throttle, const float *act_steer);
/* synthetic Philox4x32-10 action stream; writes the ACT_* buffers */
void oko_fill_random_actions(OkoEnv *env, uint64_t step, uint32_t seed);

/* scalar helpers exposed for unit tests */
void     oko_sincosf(float x, float *s, float *c);
int32_t  oko_nearest_index(const OkoEnv *env, int track, float x, float y);      /* RaceTrack.cpp:16-31 */
float    oko_dist_lane_center(const OkoEnv *env, int track, float x, float y);   /* RaceTrack.cpp:53-72 */
float    oko_dist_boundary(const OkoEnv *env, int track, float x, float y);      /* RaceTrack.cpp:33-51 */
void     oko_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float    oko_normalize_angle_deg(float a);                                       /* Utils.h:3-14 */

#ifdef __cplusplus
}
#endif
#endif
