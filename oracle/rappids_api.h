/* oracle/rappids_api.h -- TEST INFRASTRUCTURE.
 *
 * C interface shared by the two CPU oracles of the RAPPIDS planner path:
 *   oracle/_ref/libagf_rappids_ref_*.so   the UNMODIFIED reference sources
 *       Components/Components/DepthImagePlanner/DepthImagePlanner.cpp
 *       Components/Components/TrajectoryGenerator/{RapidTrajectoryGenerator,SingleAxisTrajectory}.cpp
 *     driven by oracle/ref_harness_rappids.cpp
 *   oracle/libagf_rappids_port_*.so       the independent restatement oracle/port/agf_rappids_port.cpp
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load these.  The product
 * (agri-fly_b200/, include/) never does.
 *
 * Determinism: the reference bounds planning by wall-clock time (DepthImagePlanner.cpp:119-123).
 * Both oracles run it with an unbounded time budget and a FINITE candidate generator (the
 * `returnVal < 0` exit the reference provides for exactly that, DepthImagePlanner.cpp:128-134).
 */
#ifndef ORC_RAPPIDS_API_H_
#define ORC_RAPPIDS_API_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_rappids_cfg {
  int32_t width, height;                 /* depth image size (CV_16UC1)                                 */
  double depth_scale;                    /* metres per pixel unit                                       */
  double focal_length, cx, cy;           /* pinhole intrinsics [pixels]                                 */
  double true_radius, planning_radius;   /* DepthImagePlanner ctor: physical / planning vehicle radius  */
  double min_checking_dist;              /* minimumCollisionDistance                                    */
  double min_thrust, max_thrust;         /* SetDynamicFeasiblityParameters (defaults 5, 30)             */
  double max_angvel, min_section_time;   /* defaults 20, 0.02                                           */
  double max_velocity;                   /* _maximumAllowedVelocity is 5 in the reference (no setter); must be 5 for the ref oracle */
  int32_t max_pyramids;                  /* SetMaxNumberOfPyramids; <=0 = unlimited                     */
  int32_t cost_kind;                     /* 0: -dir.p(T)/T (DepthImagePlanner.hpp:431-449); 1: -(|G|-|G-p(T)|)/T (Rappids_Simulator/main.cpp:95-109) */
  double cost_vec[3];                    /* exploration direction, or goal in the camera frame          */
} orc_rappids_cfg;

/* candidate sampling box of RandomTrajectoryGenerator (DepthImagePlanner.hpp:334-352 defaults:
 * pixels in [0.1,0.9] of the image, depth in [1.5,3], duration in [2,3]) */
typedef struct orc_rappids_sampler {
  double min_x, max_x, min_y, max_y, min_depth, max_depth, min_time, max_time;
} orc_rappids_sampler;

#define ORC_RAPPIDS_PYRAMID_DOUBLES 17   /* depth, right, top, left, bottom, normals[4][3] */

typedef struct orc_rappids_out {
  int32_t found, best_index;
  int32_t n_generated, n_cost_checks, n_collision_checks, n_velocity_checks, n_collision_free;
  int32_t n_pyramids;
  double best_cost;
  double best_coeffs[18];                /* GetTrajectory().GetCoeffs(): [6][3], t^5 first  */
  double best_tf;
} orc_rappids_out;

const char* orc_rappids_flavour(void);

/* One planner invocation: FindLowestCostTrajectory on a fresh DepthImagePlanner.
 *   candidates != NULL : [n][4] = goal position (camera frame) + duration, evaluated in order
 *   candidates == NULL : drawn from RandomTrajectoryGenerator(std::mt19937(seed)) with `sampler`
 *                        (NULL = the reference's default box); the draws are written to candidates_out
 *   results[n]         : TrajectoryTestResult bit mask of each candidate (DepthImagePlanner.hpp:38-44)
 *   pyramids           : up to max_pyr_out records of ORC_RAPPIDS_PYRAMID_DOUBLES doubles (depth order) */
int orc_rappids_plan(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3],
                     const double acc0[3], const double grav[3], int32_t n_candidates,
                     const double* candidates, int32_t seed, const orc_rappids_sampler* sampler,
                     orc_rappids_out* out, uint8_t* results, double* candidates_out,
                     double* pyramids, int32_t max_pyr_out);

/* Many independent planner invocations on `threads` host threads (CPU baseline timing):
 * images [n][h][w], vel0/acc0/grav [n][3], candidates [n][k][4]; out [n]; results [n][k] (may be NULL). */
int orc_rappids_plan_many(const orc_rappids_cfg* cfg, int32_t n, const uint16_t* images,
                          const double* vel0, const double* acc0, const double* grav,
                          int32_t n_candidates, const double* candidates, orc_rappids_out* out,
                          uint8_t* results, int32_t threads);

/* The reference's own, algorithm-independent check of the collision test (DepthImagePlanner::IsCollisionFreeGroundTruth,
 * DepthImagePlanner.cpp:1031-1097: the trajectory sampled every 0.1 s, field-of-view test, then every pixel's ray against
 * the sphere around the vehicle; used by MeasureConservativeness, :972-1003).  candidates [n][4] as above;
 * free_out[i] = 1 when the ground truth finds candidate i collision free. */
int orc_rappids_ground_truth(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3],
                             const double acc0[3], const double grav[3], int32_t n_candidates,
                             const double* candidates, uint8_t* free_out);

/* PORT ONLY: depth-image pixels read by InflatePyramid (in the reference's scan order) over all plans since the last call */
uint64_t orc_rappids_pixels_read(void);

/* pieces, for unit pins */
int orc_rappids_solve_cubic(double a, double b, double c, double roots[3]);
int orc_rappids_solve_quartic(double a, double b, double c, double d, double roots[4]);
/* RapidTrajectoryGenerator from (0, vel0, acc0, grav) to (goal, 0, 0) in T: alpha/beta/gamma per axis [3][3],
 * CheckInputFeasibility result, CheckVelocityFeasibility result */
int orc_rappids_primitive(const double vel0[3], const double acc0[3], const double grav[3],
                          const double goal[3], double T, double fmin, double fmax, double wmax,
                          double min_section, double vmax, double abg[9], int32_t* input_res,
                          int32_t* vel_res);

#ifdef __cplusplus
}
#endif
#endif
