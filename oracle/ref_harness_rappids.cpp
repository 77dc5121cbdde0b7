// oracle/ref_harness_rappids.cpp -- TEST INFRASTRUCTURE (builds oracle/_ref/libagf_rappids_ref_*.so).
//
// Drives the UNMODIFIED reference RAPPIDS planner
//   Components/Components/DepthImagePlanner/DepthImagePlanner.cpp
//   Components/Components/TrajectoryGenerator/{RapidTrajectoryGenerator,SingleAxisTrajectory}.cpp
//   Common/Common/Math/{RootFinder,Trajectory}.hpp
// (compiled where they lie under /root/reference by oracle/Makefile against oracle/shim/opencv2)
// through the C interface of oracle/rappids_api.h.  No reference source is copied; this file only
// calls the reference's public API (FindLowestCostTrajectory with a caller-supplied cost function
// and candidate generator, DepthImagePlanner.hpp:160-171) and reads private members for parity dumps.
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <iostream>
#include <limits>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "Components/DepthImagePlanner/DepthImagePlanner.hpp"
#undef private
#undef protected

#include "rappids_api.h"

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "ref-glibc"
#endif

using RapidQuadrocopterTrajectoryGenerator::RapidTrajectoryGenerator;
using RectangularPyramidPlanner::DepthImagePlanner;
using RectangularPyramidPlanner::TrajectoryTest;

namespace {

struct CostObj {
  int kind;
  Vec3d v;
  // kind 0 == DepthImagePlanner::ExplorationCost::GetCost (DepthImagePlanner.hpp:436-441)
  // kind 1 == Rappids_Simulator ExplorationCost::GetTrajCost with the goal already in the camera
  //           frame (Simulator/Rappids_Simulator/main.cpp:95-109)
  static double Wrapper(void* p, RapidTrajectoryGenerator& traj) {
    CostObj* c = (CostObj*)p;
    double duration = traj.GetFinalTime();
    if (c->kind == 0) return -c->v.Dot(traj.GetPosition(duration)) / duration;
    Vec3d Pi_C = traj.GetPosition(duration);
    Vec3d G_C = c->v;
    Vec3d S_C = Vec3d(0, 0, 0);
    double SG = (G_C - S_C).GetNorm2();
    double PiG = (G_C - Pi_C).GetNorm2();
    return -(SG - PiG) / duration;
  }
};

struct GenObj {
  int n, i;
  const double* cands;                                   // [n][4] or NULL
  DepthImagePlanner::RandomTrajectoryGenerator* rnd;     // used when cands == NULL
  double* cands_out;                                     // [n][4] or NULL
  static int Wrapper(void* p, RapidTrajectoryGenerator& nextTraj) {
    GenObj* g = (GenObj*)p;
    if (g->i >= g->n) return -1;
    if (g->cands) {
      // the calls RandomTrajectoryGenerator::GetNextCandidateTrajectory makes (DepthImagePlanner.hpp:383-393)
      const double* c = g->cands + 4 * g->i;
      nextTraj.Reset();
      nextTraj.SetGoalPosition(Vec3d(c[0], c[1], c[2]));
      nextTraj.SetGoalVelocity(Vec3d(0, 0, 0));
      nextTraj.SetGoalAcceleration(Vec3d(0, 0, 0));
      nextTraj.Generate(c[3]);
    } else {
      g->rnd->GetNextCandidateTrajectory(nextTraj);
    }
    if (g->cands_out) {
      double* o = g->cands_out + 4 * g->i;
      for (int a = 0; a < 3; a++) o[a] = nextTraj._axis[a]._pf;
      o[3] = nextTraj._tf;
    }
    g->i++;
    return 0;
  }
};

void configure(DepthImagePlanner& pl, const orc_rappids_cfg* cfg) {
  pl.SetDynamicFeasiblityParameters(cfg->min_thrust, cfg->max_thrust, cfg->max_angvel, cfg->min_section_time);
  pl._maximumAllowedVelocity = cfg->max_velocity;
  if (cfg->max_pyramids > 0) pl.SetMaxNumberOfPyramids(cfg->max_pyramids);
}

int plan_one(const orc_rappids_cfg* cfg, const uint16_t* image, const double* vel0, const double* acc0,
             const double* grav, int n, const double* cands, int seed, const orc_rappids_sampler* sm,
             orc_rappids_out* out, uint8_t* results, double* cands_out, double* pyramids, int max_pyr_out) {
  cv::Mat img;
  img.rows = cfg->height;
  img.cols = cfg->width;
  img.data = (unsigned char*)image;
  DepthImagePlanner planner(img, cfg->depth_scale, cfg->focal_length, cfg->cx, cfg->cy, cfg->true_radius,
                            cfg->planning_radius, cfg->min_checking_dist);
  configure(planner, cfg);
  planner.SetRandomSeed(seed);
  RapidTrajectoryGenerator traj(Vec3d(0, 0, 0), Vec3d(vel0[0], vel0[1], vel0[2]),
                                Vec3d(acc0[0], acc0[1], acc0[2]), Vec3d(grav[0], grav[1], grav[2]));
  CostObj cost{cfg->cost_kind, Vec3d(cfg->cost_vec[0], cfg->cost_vec[1], cfg->cost_vec[2])};
  std::unique_ptr<DepthImagePlanner::RandomTrajectoryGenerator> rnd;
  if (!cands) {
    if (sm)
      rnd.reset(new DepthImagePlanner::RandomTrajectoryGenerator((int)sm->min_x, (int)sm->max_x, (int)sm->min_y,
                                                                 (int)sm->max_y, sm->min_depth, sm->max_depth,
                                                                 sm->min_time, sm->max_time, seed, &planner));
    else
      rnd.reset(new DepthImagePlanner::RandomTrajectoryGenerator(&planner));
  }
  GenObj gen{n, 0, cands, rnd.get(), cands_out};
  std::vector<TrajectoryTest> trajectories;
  trajectories.reserve(n);
  bool found = planner.FindLowestCostTrajectory(traj, trajectories, 1.0e3 /* unbounded: int(1e9 us) */, &cost,
                                                &CostObj::Wrapper, &gen, &GenObj::Wrapper);
  memset(out, 0, sizeof(*out));
  out->found = found ? 1 : 0;
  out->best_index = -1;
  out->n_generated = planner.GetNumTrajectoriesGenerated();
  out->n_cost_checks = planner.GetNumCostChecks();
  out->n_collision_checks = planner.GetNumCollisionChecks();
  out->n_velocity_checks = planner.GetNumVelocityChecks();
  out->n_collision_free = planner.GetNumCollisionFree();
  out->n_pyramids = planner.GetNumPyramids();
  out->best_cost = std::numeric_limits<double>::max();
  for (size_t i = 0; i < trajectories.size(); i++) {
    if (results) results[i] = (uint8_t)trajectories[i].result;
    if (trajectories[i].result & RectangularPyramidPlanner::CollisionFree) out->best_index = (int)i;
  }
  if (found) {
    out->best_cost = CostObj::Wrapper(&cost, traj);
    CommonMath::Trajectory t = traj.GetTrajectory();
    std::vector<Vec3d> c = t.GetCoeffs();
    for (int k = 0; k < 6; k++)
      for (int a = 0; a < 3; a++) out->best_coeffs[3 * k + a] = c[k][a];
    out->best_tf = t.GetEndTime();
  }
  if (pyramids) {
    std::vector<RectangularPyramidPlanner::Pyramid> ps = planner.GetPyramids();
    for (int i = 0; i < (int)ps.size() && i < max_pyr_out; i++) {
      double* p = pyramids + ORC_RAPPIDS_PYRAMID_DOUBLES * i;
      p[0] = ps[i].depth;
      p[1] = ps[i].rightPixBound;
      p[2] = ps[i].topPixBound;
      p[3] = ps[i].leftPixBound;
      p[4] = ps[i].bottomPixBound;
      for (int f = 0; f < 4; f++)
        for (int a = 0; a < 3; a++) p[5 + 3 * f + a] = ps[i].planeNormals[f][a];
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char* orc_rappids_flavour(void) { return ORC_FLAVOUR; }

int orc_rappids_plan(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3],
                     const double acc0[3], const double grav[3], int32_t n_candidates,
                     const double* candidates, int32_t seed, const orc_rappids_sampler* sampler,
                     orc_rappids_out* out, uint8_t* results, double* candidates_out, double* pyramids,
                     int32_t max_pyr_out) {
  return plan_one(cfg, image, vel0, acc0, grav, n_candidates, candidates, seed, sampler, out, results,
                  candidates_out, pyramids, max_pyr_out);
}

int orc_rappids_plan_many(const orc_rappids_cfg* cfg, int32_t n, const uint16_t* images, const double* vel0,
                          const double* acc0, const double* grav, int32_t k, const double* candidates,
                          orc_rappids_out* out, uint8_t* results, int32_t threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  const size_t npix = (size_t)cfg->width * cfg->height;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      for (int i = (int)((int64_t)n * t / threads); i < (int)((int64_t)n * (t + 1) / threads); i++)
        plan_one(cfg, images + npix * i, vel0 + 3 * i, acc0 + 3 * i, grav + 3 * i, k, candidates + (size_t)4 * k * i,
                 0, nullptr, out + i, results ? results + (size_t)k * i : nullptr, nullptr, nullptr, 0);
    });
  }
  for (auto& th : pool) th.join();
  return 0;
}

int orc_rappids_ground_truth(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3], const double acc0[3],
                             const double grav[3], int32_t n, const double* cands, uint8_t* free_out) {
  cv::Mat img;
  img.rows = cfg->height;
  img.cols = cfg->width;
  img.data = (unsigned char*)image;
  DepthImagePlanner planner(img, cfg->depth_scale, cfg->focal_length, cfg->cx, cfg->cy, cfg->true_radius,
                            cfg->planning_radius, cfg->min_checking_dist);
  configure(planner, cfg);
  for (int i = 0; i < n; i++) {
    const double* c = cands + 4 * i;
    RapidTrajectoryGenerator traj(Vec3d(0, 0, 0), Vec3d(vel0[0], vel0[1], vel0[2]), Vec3d(acc0[0], acc0[1], acc0[2]),
                                  Vec3d(grav[0], grav[1], grav[2]));
    traj.SetGoalPosition(Vec3d(c[0], c[1], c[2]));
    traj.SetGoalVelocity(Vec3d(0, 0, 0));
    traj.SetGoalAcceleration(Vec3d(0, 0, 0));
    traj.Generate(c[3]);
    free_out[i] = planner.IsCollisionFreeGroundTruth(traj.GetTrajectory()) ? 1 : 0;
  }
  return 0;
}

int orc_rappids_solve_cubic(double a, double b, double c, double roots[3]) {
  return (int)RootFinder::solve_cubic<double>(a, b, c, roots);
}
int orc_rappids_solve_quartic(double a, double b, double c, double d, double roots[4]) {
  return (int)RootFinder::solve_quartic<double>(a, b, c, d, roots);
}

int orc_rappids_primitive(const double vel0[3], const double acc0[3], const double grav[3], const double goal[3],
                          double T, double fmin, double fmax, double wmax, double min_section, double vmax,
                          double abg[9], int32_t* input_res, int32_t* vel_res) {
  RapidTrajectoryGenerator traj(Vec3d(0, 0, 0), Vec3d(vel0[0], vel0[1], vel0[2]),
                                Vec3d(acc0[0], acc0[1], acc0[2]), Vec3d(grav[0], grav[1], grav[2]));
  traj.SetGoalPosition(Vec3d(goal[0], goal[1], goal[2]));
  traj.SetGoalVelocity(Vec3d(0, 0, 0));
  traj.SetGoalAcceleration(Vec3d(0, 0, 0));
  traj.Generate(T);
  for (int a = 0; a < 3; a++) {
    abg[3 * a + 0] = traj.GetAxisParamAlpha(a);
    abg[3 * a + 1] = traj.GetAxisParamBeta(a);
    abg[3 * a + 2] = traj.GetAxisParamGamma(a);
  }
  *input_res = (int)traj.CheckInputFeasibility(fmin, fmax, wmax, min_section);
  *vel_res = (int)traj.CheckVelocityFeasibility(vmax);
  return 0;
}

}  // extern "C"
