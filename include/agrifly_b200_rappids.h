/* agrifly_b200_rappids.h -- C ABI of the batched RAPPIDS planner (C5 / K6 / N3).
 *
 * Drop-in boundary for the reference's depth-image planner call
 *   RectangularPyramidPlanner::DepthImagePlanner::FindLowestCostTrajectory
 *   (Components/Components/DepthImagePlanner/DepthImagePlanner.cpp:91-214, declared DepthImagePlanner.hpp:160-171)
 * as Rappids_Simulator makes it once per depth image and vehicle
 *   (Simulator/Rappids_Simulator/main.cpp:484-503),
 * run for a whole population in one kernel launch: one planner invocation per vehicle, each on its own
 * depth image, initial state and candidate list.  Same library as agrifly_b200.h (libagrifly_b200.so),
 * same conventions: plain C, int return codes (AGF_OK / negative AGF_E*), opaque handle, caller-owned
 * host buffers, one handle per GPU, no CPU fallback.
 *
 * Determinism.  The reference bounds a planner call by wall-clock time (DepthImagePlanner.cpp:119-123);
 * the batched call evaluates a FIXED number of candidates per vehicle instead (the finite-generator exit
 * the reference offers, DepthImagePlanner.cpp:128-134), in order, with the reference's pruning: cost
 * check, input feasibility, velocity check, pyramid collision check, pyramids generated on demand and
 * reused by later candidates of the same image.  Results are a pure function of the inputs.
 *
 * Frames: camera-fixed, x right, y down, z into the image; trajectories start at the focal point.
 */
#ifndef AGRIFLY_B200_RAPPIDS_H_
#define AGRIFLY_B200_RAPPIDS_H_

#include <stddef.h>
#include <stdint.h>

#include "agrifly_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* TrajectoryTestResult bit mask (DepthImagePlanner.hpp:38-44) */
#define AGF_RAPPIDS_LOW_COST 1
#define AGF_RAPPIDS_DYNAMICS_FEASIBLE 2
#define AGF_RAPPIDS_VELOCITY_ADMISSIBLE 4
#define AGF_RAPPIDS_COLLISION_FREE 8

#define AGF_RAPPIDS_MAX_PYRAMIDS 32      /* pyramids kept per vehicle; SetMaxNumberOfPyramids beyond this is clamped */
#define AGF_RAPPIDS_PYRAMID_DOUBLES 17   /* depth, right, top, left, bottom, plane normals [4][3] (Pyramid.hpp:62-75) */
#define AGF_RAPPIDS_MAX_BOXES 4          /* synthetic scene description, agf_rappids_render_scenes */

enum { AGF_RAPPIDS_COST_DIRECTION = 0, /* -dir . p(T) / T      DepthImagePlanner.hpp:431-449                   */
       AGF_RAPPIDS_COST_GOAL = 1 };    /* -(|G| - |G - p(T)|)/T  Simulator/Rappids_Simulator/main.cpp:95-109 */

/* DepthImagePlanner constructor arguments (DepthImagePlanner.cpp:27-62) and the limits it initialises
 * (:44-52, SetDynamicFeasiblityParameters DepthImagePlanner.hpp:245-253, SetMaxNumberOfPyramids :260-262),
 * plus the sampling box of RandomTrajectoryGenerator (DepthImagePlanner.hpp:334-381). */
typedef struct agf_rappids_cfg {
  int32_t width, height;               /* depth image, CV_16UC1, row-major                              */
  double depth_scale;                  /* metres per pixel unit                                         */
  double focal_length, cx, cy;
  double true_radius, planning_radius, min_checking_dist;
  double min_thrust, max_thrust, max_angvel, min_section_time, max_velocity;
  int32_t max_pyramids;                /* 1 .. AGF_RAPPIDS_MAX_PYRAMIDS: SetMaxNumberOfPyramids; <= 0: the reference's default
                                          (unlimited, DepthImagePlanner.cpp:52) -- device storage holds AGF_RAPPIDS_MAX_PYRAMIDS, and
                                          agf_rappids_result.pyramid_cap_hit reports a vehicle that needed more (its result may
                                          then differ from the reference's); > AGF_RAPPIDS_MAX_PYRAMIDS: AGF_EUNSUPPORTED      */
  int32_t cost_kind;                   /* AGF_RAPPIDS_COST_*                                            */
  double cost_vec[3];                  /* direction / goal shared by the population (agf_rappids_set_goals overrides per vehicle) */
  double sample_min_x, sample_max_x, sample_min_y, sample_max_y;   /* pixels  */
  double sample_min_depth, sample_max_depth;                       /* metres  */
  double sample_min_time, sample_max_time;                         /* seconds */
  int32_t math;                        /* AGF_MATH_PARITY (bit-comparable with the oracle) or AGF_MATH_FAST */
  int32_t device;                      /* CUDA device ordinal, -1 = current                             */
} agf_rappids_cfg;

/* what FindLowestCostTrajectory returns and the counters the planner keeps (DepthImagePlanner.hpp:176-215) */
typedef struct agf_rappids_result {
  int32_t found;                       /* return value of FindLowestCostTrajectory                      */
  int32_t best_index;                  /* index of the returned candidate, -1 if none                   */
  int32_t n_generated, n_cost_checks, n_collision_checks, n_velocity_checks, n_collision_free;
  int32_t n_pyramids;
  double best_cost;
  double best_coeffs[18];              /* RapidTrajectoryGenerator::GetTrajectory().GetCoeffs(): [6][3], t^5 first */
  double best_tf;
  int32_t pyramid_cap_hit;             /* 1: a candidate needed a new pyramid when max_pyramids existed already and was rejected
                                          as colliding (DepthImagePlanner.cpp:246-249); with max_pyramids <= 0 ("unlimited") this
                                          marks a result that may differ from the reference's                                */
  int32_t reserved_;
} agf_rappids_result;

typedef struct agf_rappids agf_rappids;

/* Fills `cfg` with Rappids_Simulator's settings for a width x height image (main.cpp:121-122,167-169,360,
 * 484-489: depth scale 10/256, f = cx = width/2, cy = height/2, radii 0.116/0.174 m, 0.5 m) and the
 * DepthImagePlanner defaults (thrust 5..30 m/s^2, 20 rad/s, 0.02 s, 5 m/s; sampling box 10-90 % of the
 * image, 1.5-3 m, 2-3 s). */
int agf_rappids_cfg_default(int32_t width, int32_t height, agf_rappids_cfg* cfg);

/* One handle plans for `n_vehicles` vehicles with at most `max_candidates` candidates each. */
int agf_rappids_create(const agf_rappids_cfg* cfg, size_t n_vehicles, int32_t max_candidates, agf_rappids** out);
int agf_rappids_destroy(agf_rappids* p);
size_t agf_rappids_size(const agf_rappids* p);
void* agf_rappids_stream(const agf_rappids* p);

/* Depth images of vehicles first .. first+count-1 from host memory, [count][height][width] uint16
 * (the cv::Mat the reference's constructor takes). */
int agf_rappids_set_images(agf_rappids* p, const uint16_t* images, size_t first, size_t count);
/* Synthetic scenes rasterised on the device: row_bg [count][height] uint16 (value of every pixel of a row),
 * boxes [count][AGF_RAPPIDS_MAX_BOXES][5] int32 (x0, x1, y0, y1, value; x1/y1 exclusive; value 0 = unused);
 * pixel = min(row value, values of the boxes covering it). */
int agf_rappids_render_scenes(agf_rappids* p, const uint16_t* row_bg, const int32_t* boxes, size_t first, size_t count);
int agf_rappids_get_images(agf_rappids* p, uint16_t* images, size_t first, size_t count);

/* Initial state of the candidate trajectories, [count][3] each: velocity, acceleration and gravity in the camera
 * frame (the RapidTrajectoryGenerator constructor arguments, Rappids_Simulator main.cpp:491-497; position is 0). */
int agf_rappids_set_states(agf_rappids* p, const double* vel0, const double* acc0, const double* grav, size_t first,
                           size_t count);
/* Per-vehicle cost vector (exploration direction or goal in the camera frame), [count][3]. */
int agf_rappids_set_goals(agf_rappids* p, const double* goals, size_t first, size_t count);
/* Candidate list, [count][k][4]: end position (camera frame) and duration; candidates come to rest
 * (RandomTrajectoryGenerator::GetNextCandidateTrajectory, DepthImagePlanner.hpp:383-393).  k <= max_candidates. */
int agf_rappids_set_candidates(agf_rappids* p, const double* candidates, int32_t k, size_t first, size_t count);
/* Draw k candidates per vehicle on the device from the sampling box (counter-based Philox keyed by `seed`,
 * counter = (global vehicle index, candidate)); first_global_index keeps draws independent of the sharding. */
int agf_rappids_sample_candidates(agf_rappids* p, int32_t k, uint64_t seed, uint64_t first_global_index);
int agf_rappids_get_candidates(agf_rappids* p, double* candidates, size_t first, size_t count);

/* The planner call for every vehicle: evaluates the current candidate lists.  Asynchronous on the handle's stream. */
int agf_rappids_plan(agf_rappids* p);
int agf_rappids_sync(agf_rappids* p);
/* Order in which the planning pass hands vehicles to its warps.  BY_LAST_WORK (default): descending device time of each
 * vehicle's previous plan on this handle (index order on the first call) -- plans differ in length by an order of magnitude
 * and a planner runs at image rate on slowly changing scenes, so starting the long plans first removes the tail of a launch.
 * INDEX: vehicle index order.  Results do not depend on the order (the reference has no counterpart: one planner object
 * per vehicle, DepthImagePlanner.cpp:91). */
#define AGF_RAPPIDS_DISPATCH_INDEX 0
#define AGF_RAPPIDS_DISPATCH_BY_LAST_WORK 1
int agf_rappids_set_dispatch(agf_rappids* p, int32_t mode);
/* Validation knobs, read from the environment when a handle is created (results are bit-identical for every setting; the
 * GPU tests compare them): AGF_RAPPIDS_FRAME_JUMP=<k> -- iterations of InflatePyramid's spiral expansion taken in one step
 * when their frame holds no blocking pixel (default 8, 0 = line by line as the reference scans); AGF_RAPPIDS_SHRINK_FOLD=0 --
 * one shrink update per pixel instead of one reduction per 32-pixel span of an edge region. */
/* Device clock cycles the last plan spent on each vehicle (what BY_LAST_WORK sorts by), [count] uint32. */
int agf_rappids_get_plan_work(agf_rappids* p, uint32_t* cycles, size_t first, size_t count);

int agf_rappids_get_results(agf_rappids* p, agf_rappids_result* out, size_t first, size_t count);
/* The returned trajectories as records for the in-kernel tracking loop, [count][AGF_OFFTRAJ_DOUBLES] (agrifly_b200.h,
 * agf_batch_set_offboard_trajectories): the primitive in SingleAxisTrajectory's own variables (p0 = 0, v0, a0 of the
 * vehicle, alpha / beta / gamma of the best candidate -- the object Rappids_Simulator copies into `_traj`, main.cpp:524-526),
 * gravity in the trajectory frame and the end time; trajAtt is written as identity and trajOffset as zero, to be filled in
 * by the caller (estState.att * depthCamAtt and estState.pos, main.cpp:519-522).  Vehicles without a trajectory get tf = 0. */
int agf_rappids_get_tracking_primitives(agf_rappids* p, double* records, size_t first, size_t count);
/* The same records written on the device, field-major [AGF_OFFTRAJ_DOUBLES][n_dst] (vehicle v of this planner -> column
 * dst_first + v), straight into a step batch's trajectory table (agf_batch_offboard_trajectories_device_ptr): the planner's
 * result reaches the tracking loop without a host hop.  att [n][4] / offset [n][3]: host arrays or NULL (identity / zero).
 * Runs on the planner's stream; the call returns after the kernel has finished. */
int agf_rappids_export_tracking_primitives(agf_rappids* p, double* dev_dst, size_t n_dst, size_t dst_first, const double* att,
                                           const double* offset);
/* TrajectoryTestResult of every candidate, [count][k] bytes (the `trajectories` vector of the reference call). */
int agf_rappids_get_candidate_flags(agf_rappids* p, uint8_t* flags, size_t first, size_t count);
/* GetPyramids(): [count][AGF_RAPPIDS_MAX_PYRAMIDS][AGF_RAPPIDS_PYRAMID_DOUBLES], depth order, unused records NaN. */
int agf_rappids_get_pyramids(agf_rappids* p, double* pyramids, size_t first, size_t count);

/* Population statistics of the last plan, reduced on the device: [0] vehicles with a trajectory, [1] candidates
 * generated, [2] cost checks passed, [3] input-feasible, [4] velocity-admissible, [5] collision-free, [6] pyramids,
 * [7] sum of best costs over vehicles with a trajectory.  `host_out` receives 8 doubles. */
int agf_rappids_reduce_stats(agf_rappids* p, double* host_out);
int agf_rappids_reduce_stats_device(agf_rappids* p, double* dev_out);

/* Device time of the planner kernel (CUDA events on the handle's stream): mean ms over the plans since the
 * last call, and their number. */
int agf_rappids_plan_kernel_time(agf_rappids* p, double* ms, uint64_t* launches);
uint64_t agf_rappids_launch_count(const agf_rappids* p);

#ifdef __cplusplus
}
#endif
#endif
