// agf_rappids.cu -- host side of the batched RAPPIDS planner behind include/agrifly_b200_rappids.h:
// the handle, device buffers (depth images [n][H][W] + transposed copies [n][W][H], states, candidate lists,
// results), the small kernels around the planner (scene rasteriser, image transpose, Philox candidate sampler,
// population statistics) and the C ABI.  The planner kernel itself is agf_rappids_plan.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <utility>
#include <vector>

#include "agf_rappids_plan.cuh"
#include "agrifly_b200_rappids.h"

namespace agf {
int fail_from(int code, const char* what, int cuda_error);  // agf_batch.cu (sets agf_last_error_string)
}
namespace agfr {
cudaError_t launch_plan_parity(const PlanParams& P, int grid, cudaStream_t stream);
cudaError_t launch_plan_fast(const PlanParams& P, int grid, cudaStream_t stream);
cudaError_t plan_blocks_per_sm_parity(int* blocks, int* regs);
cudaError_t plan_blocks_per_sm_fast(int* blocks, int* regs);
}  // namespace agfr

namespace {

int fail(int code, const char* what, cudaError_t e = cudaSuccess) { return agf::fail_from(code, what, (int)e); }
#define AGFR_CUDA(call)                                         \
  do {                                                          \
    cudaError_t e_ = (call);                                    \
    if (e_ != cudaSuccess) return fail(AGF_ECUDA, #call, e_);   \
  } while (0)

static_assert(sizeof(agf_rappids_result) == sizeof(agfr::ResultRec), "result record layout");
static_assert(AGF_RAPPIDS_MAX_PYRAMIDS == agfr::kMaxPyr, "pyramid capacity");

// pixel = min(row value, boxes covering it); written row-major (T == false) or transposed (T == true), the
// thread index running along the contiguous dimension of the destination in both cases
template<bool T>
__global__ void render_kernel(uint16_t* __restrict__ dst, const uint16_t* __restrict__ row_bg,
                              const int32_t* __restrict__ boxes, int W, int H, size_t count, size_t vstride) {
  const size_t npix = (size_t)W * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count * npix;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t v = idx / npix;
    const int r = (int)(idx - v * npix);
    const int x = T ? r / H : r % W;
    const int y = T ? r % H : r / W;
    unsigned val = row_bg[v * H + y];
    const int32_t* b = boxes + v * AGF_RAPPIDS_MAX_BOXES * 5;
#pragma unroll
    for (int k = 0; k < AGF_RAPPIDS_MAX_BOXES; k++) {
      const int x0 = b[5 * k], x1 = b[5 * k + 1], y0 = b[5 * k + 2], y1 = b[5 * k + 3], bv = b[5 * k + 4];
      if (bv > 0 && x >= x0 && x < x1 && y >= y0 && y < y1) val = min(val, (unsigned)bv);
    }
    dst[v * vstride + r] = (uint16_t)val;
  }
}

// [H][W] -> [W][H] of every vehicle's scene record through 32x32 shared-memory tiles (both sides coalesced)
__global__ void transpose_kernel(uint16_t* __restrict__ dst, const uint16_t* __restrict__ src, int W, int H, size_t vstride) {
  __shared__ uint16_t tile[32][34];
  const size_t v = blockIdx.z;
  const uint16_t* s = src + v * vstride;
  uint16_t* d = dst + v * vstride;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int x = bx + threadIdx.x, y = by + j;
    if (x < W && y < H) tile[j][threadIdx.x] = s[(size_t)y * W + x];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int y = by + threadIdx.x, x = bx + j;
    if (x < W && y < H) d[(size_t)x * H + y] = tile[threadIdx.x][j];
  }
}

// Minimum of every group of 32 consecutive pixels of every line of `plane` ([L lines][pitch] per vehicle, vehicles `vstride`
// apart in both the plane and the table); pixels <= ignore count as 65535 (they are never "seen" by the planner,
// DepthImagePlanner.cpp:506).  One thread per (line, group).
__global__ void groupmin_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ plane, int pitch, int G, int L, size_t lines,
                                int ignore, size_t vstride) {
  const size_t total = lines * (size_t)G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t line = idx / G;
    const int g = (int)(idx - line * G);
    const size_t v = line / L;
    const int l = (int)(line - v * L);
    const uint16_t* src = plane + v * vstride + (size_t)l * pitch + (size_t)g * 32;
    const int cnt = min(32, pitch - g * 32);
    unsigned m = 65535u;
    for (int k = 0; k < cnt; k++) {
      const unsigned p = src[k];
      if ((int)p > ignore) m = min(m, p);
    }
    out[v * vstride + (size_t)l * G + g] = (uint16_t)m;
  }
}

// Dispatch order of the planning pass: vehicles by DESCENDING work of their previous plan (clock cycles, written by the
// planning kernel), 2048 linear buckets, one block.  Plans differ in length by an order of magnitude (pyramids inflated);
// handed out in index order the launch ends with a tail in which a few warps finish long plans on an otherwise empty GPU
// (ncu, round 2: SMs active 42 % of a 16 384-plan launch), and a planner runs at image rate on scenes that change little
// between frames, so the last frame's work is a good forecast.  Only the order changes -- results do not depend on it.
// *nlong = the number of vehicles at the head of the order whose work was at least `coop_factor` times the launch's ideal
// length (total work / warp slots of the grid; whole buckets): only those can stretch the launch beyond what the rest of the
// population needs anyway, and they are planned by a whole CTA each (agf_rappids_plan.cuh, plan_vehicle).
__global__ void __launch_bounds__(1024) dispatch_order_kernel(const unsigned* __restrict__ work, int n, int* __restrict__ order,
                                                              int* __restrict__ nlong, float coop_factor, float coop_cap, int slots) {
  __shared__ unsigned hist[2048];
  __shared__ unsigned wmax;
  __shared__ unsigned long long wsum;
  const int t = threadIdx.x;
  if (t == 0) { wmax = 1; wsum = 0; }
  for (int i = t; i < 2048; i += 1024) hist[i] = 0;
  __syncthreads();
  unsigned m = 0;
  unsigned long long sum = 0;
  for (int v = t; v < n; v += 1024) {
    m = max(m, work[v]);
    sum += work[v];
  }
  atomicMax(&wmax, m);
  atomicAdd(&wsum, sum);
  __syncthreads();
  const unsigned long long mx = wmax;
  for (int v = t; v < n; v += 1024) atomicAdd(&hist[2047 - (unsigned)((unsigned long long)work[v] * 2047ull / mx)], 1u);
  __syncthreads();
  if (t == 0) {
    // buckets 0 .. cut-1 hold work >= (2048 - cut) / 2047 * max >= coop_factor * mean
    int cut = 0;
    if (coop_factor > 0.0f) {
      const double thr = (double)coop_factor * (double)wsum / (double)slots;
      const double b = thr * 2047.0 / (double)mx;  // bin = 2047 - floor(work * 2047 / max) < cut  <=>  floor(..) > 2047 - cut
      cut = b >= 2047.0 ? 0 : 2047 - (int)b;       // floor(work * 2047 / max) >= floor(b) + 1 > b
      if (cut < 0) cut = 0;
      // ... unless together they hold more than `coop_cap` of the total work: then the launch is bound by its total work, not
      // by its longest plan (many long plans pack well), and speculation -- which spends about twice the warp time of the
      // sequential loop on a vehicle -- would only cost (measured on the hard scene family: 245 -> 260-277 ms)
      double accw = 0.0;
      for (int i = 0; i < cut; i++) accw += (double)hist[i] * ((double)(2047 - i) + 0.5) / 2047.0 * (double)mx;
      if (accw > (double)coop_cap * (double)wsum) cut = 0;
    }
    unsigned acc = 0;
    for (int i = 0; i < 2048; i++) {
      if (i == cut) *nlong = (int)acc;
      const unsigned c = hist[i];
      hist[i] = acc;
      acc += c;
    }
  }
  __syncthreads();
  for (int v = t; v < n; v += 1024) order[atomicAdd(&hist[2047 - (unsigned)((unsigned long long)work[v] * 2047ull / mx)], 1u)] = v;
}

// agf_rappids_export_tracking_primitives: one thread per vehicle, field-major destination
__global__ void export_prims_kernel(const double* __restrict__ prims, const double* __restrict__ state,
                                    const agf_rappids_result* __restrict__ res, const double* __restrict__ att,
                                    const double* __restrict__ offset, size_t n, double* __restrict__ dst, size_t n_dst, size_t dst_first) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  double r[AGF_OFFTRAJ_DOUBLES];
  for (int a = 0; a < 3; a++) {
    r[6 * a + 0] = 0.0;
    r[6 * a + 1] = state[v * 12 + a];
    r[6 * a + 2] = state[v * 12 + 3 + a];
    r[6 * a + 3] = prims[v * 9 + 3 * a];
    r[6 * a + 4] = prims[v * 9 + 3 * a + 1];
    r[6 * a + 5] = prims[v * 9 + 3 * a + 2];
    r[18 + a] = state[v * 12 + 6 + a];
    r[26 + a] = offset ? offset[v * 3 + a] : 0.0;
  }
  r[21] = res[v].found ? res[v].best_tf : 0.0;
  r[22] = att ? att[v * 4] : 1.0;
  for (int a = 1; a < 4; a++) r[22 + a] = att ? att[v * 4 + a] : 0.0;
  for (int k = 0; k < AGF_OFFTRAJ_DOUBLES; k++) dst[(size_t)k * n_dst + dst_first + v] = r[k];
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

struct SampleBox {
  double x0, x1, y0, y1, d0, d1, t0, t1, cx, cy, f;
};
// RandomTrajectoryGenerator::GetNextCandidateTrajectory (DepthImagePlanner.hpp:383-393) with counter-based draws:
// one Philox block per (vehicle, candidate) -> pixel x, pixel y, depth, duration
__global__ void sample_kernel(double* __restrict__ cands, size_t n, int k, int kcap, uint64_t seed, uint64_t first,
                              const SampleBox B) {
  const size_t total = n * (size_t)k;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t v = idx / k;
    const int i = (int)(idx - v * k);
    const uint64_t gv = first + v;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)gv, (uint32_t)(gv >> 32), (uint32_t)i, 0x52415050u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const double s = 1.0 / 4294967296.0;
    const double px = B.x0 + (B.x1 - B.x0) * (((double)r.x + 0.5) * s);
    const double py = B.y0 + (B.y1 - B.y0) * (((double)r.y + 0.5) * s);
    const double dp = B.d0 + (B.d1 - B.d0) * (((double)r.z + 0.5) * s);
    const double T = B.t0 + (B.t1 - B.t0) * (((double)r.w + 0.5) * s);
    double4 c;
    c.x = dp * ((px - B.cx) / B.f);
    c.y = dp * ((py - B.cy) / B.f);
    c.z = dp * 1;
    c.w = T;
    *reinterpret_cast<double4*>(cands + (v * kcap + i) * 4) = c;
  }
}

// K4-style reduction of the per-vehicle results: warp shuffles, one atomic per warp and statistic
__global__ void stats_kernel(const agf_rappids_result* __restrict__ res, size_t n, double* __restrict__ out) {
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
    const agf_rappids_result& r = res[v];
    a[0] += r.found;
    a[1] += r.n_generated;
    a[2] += r.n_cost_checks;
    a[3] += r.n_collision_checks;
    a[4] += r.n_velocity_checks;
    a[5] += r.n_collision_free;
    a[6] += r.n_pyramids;
    if (r.found) a[7] += r.best_cost;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + k, a[k]);
  }
}

struct Handle {
  agf_rappids_cfg cfg;
  size_t n = 0;
  int kcap = 0, k = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  // One SCENE RECORD per vehicle: image [H][W] | transposed image [W][H] | group minima of the rows [H][GW] | of the columns
  // [W][GH] (agf_rappids_plan.cuh), records `vstride` elements apart in one allocation.  A planning warp works on one vehicle
  // at a time and touches all four planes; as four population-wide arrays they lay on four different 2 MB pages per warp.
  uint16_t* scene = nullptr;
  size_t vstride = 0;  // in uint16 elements, a multiple of 64 (128 bytes)
  uint16_t *img = nullptr, *imgT = nullptr;     // = scene + plane offset; vehicle v's plane starts at + v * vstride
  uint16_t *gminR = nullptr, *gminC = nullptr;
  int GW() const { return (cfg.width + 31) / 32; }
  int GH() const { return (cfg.height + 31) / 32; }
  int ignore_value() const { return (int)(uint16_t)(cfg.true_radius / cfg.depth_scale); }
  int rebuild_groupmin(size_t first, size_t count) {  // after the images of these vehicles changed
    const int W = cfg.width, H = cfg.height;
    const size_t tr = count * (size_t)H * GW(), tc = count * (size_t)W * GH();
    const int gr = (int)((tr + 255) / 256 < 148 * 16 ? (tr + 255) / 256 : 148 * 16);
    const int gc = (int)((tc + 255) / 256 < 148 * 16 ? (tc + 255) / 256 : 148 * 16);
    groupmin_kernel<<<gr, 256, 0, stream>>>(gminR + first * vstride, img + first * vstride, W, GW(), H, count * (size_t)H, ignore_value(), vstride);
    groupmin_kernel<<<gc, 256, 0, stream>>>(gminC + first * vstride, imgT + first * vstride, H, GH(), W, count * (size_t)W, ignore_value(), vstride);
    launches += 2;
    return cudaGetLastError() == cudaSuccess ? AGF_OK : AGF_ECUDA;
  }
  double *state = nullptr, *cands = nullptr, *pyr = nullptr, *stats = nullptr, *prims = nullptr;
  uint8_t* flags = nullptr;
  double* ccost = nullptr;  // candidate pass -> planning pass (agf_rappids_plan.cuh)
  agf_rappids_result* results = nullptr;
  int* next = nullptr;
  int* order = nullptr;      // dispatch order of the planning pass (dispatch_order_kernel)
  int* nlong = nullptr;      // [0] vehicles at the head of the order planned by a whole CTA each, [1] their work counter
  float coop_factor = 0.3f;  // a vehicle is "long" when its previous plan took this fraction of (total work / warp slots);
                             // AGF_RAPPIDS_COOP_FACTOR at create time overrides, 0: off
  float coop_cap = 0.04f;    // ... provided all of them together hold at most this share of the total work (AGF_RAPPIDS_COOP_CAP)
  unsigned* work = nullptr;  // cycles per vehicle of the last plan
  bool have_work = false;
  int dispatch = 1;          // 1: by the previous plan's work (default), 0: index order
  int shrink_fold = 1;       // PlanParams::shrinkFold; AGF_RAPPIDS_SHRINK_FOLD=0 at create time: one update per pixel
  int frame_jump = 8;        // PlanParams::frameJump; AGF_RAPPIDS_FRAME_JUMP=<k> at create time overrides (0: line by line only)
  void* stage = nullptr;  // device staging for scene descriptions
  size_t stage_bytes = 0;
  int grid = 0, regs = 0, blocks_per_sm = 0;
  uint64_t launches = 0;
  // planner-kernel timing: a fixed ring of event pairs; when it is full the oldest (long complete) pair is folded into
  // the running totals, so a caller that plans at image rate for hours never grows the handle
  enum { EVENT_RING = 32 };
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  size_t ev_head = 0, ev_count = 0;
  double timed_ms = 0;
  uint64_t timed_launches = 0;
  cudaError_t fold_oldest_event() {
    auto& e = events[ev_head];
    float t = 0;
    cudaError_t r = cudaEventSynchronize(e.second);
    if (r == cudaSuccess) r = cudaEventElapsedTime(&t, e.first, e.second);
    if (r != cudaSuccess) return r;
    timed_ms += t;
    timed_launches++;
    ev_head = (ev_head + 1) % EVENT_RING;
    ev_count--;
    return cudaSuccess;
  }

  ~Handle() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& e : events) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    cudaFree(scene);
    cudaFree(prims);
    cudaFree(state);
    cudaFree(cands);
    cudaFree(pyr);
    cudaFree(stats);
    cudaFree(flags);
    cudaFree(ccost);
    cudaFree(results);
    cudaFree(next);
    cudaFree(order);
    cudaFree(nlong);
    cudaFree(work);
    cudaFree(stage);
    if (stream) cudaStreamDestroy(stream);
  }
  size_t npix() const { return (size_t)cfg.width * cfg.height; }
  int range(size_t first, size_t count) const {
    if (first > n || count > n - first) return fail(AGF_ERANGE, "vehicle range outside the handle");
    return AGF_OK;
  }
  int ensure_stage(size_t bytes) {
    if (bytes <= stage_bytes) return AGF_OK;
    cudaFree(stage);
    stage = nullptr;
    stage_bytes = 0;
    AGFR_CUDA(cudaMalloc(&stage, bytes));
    stage_bytes = bytes;
    return AGF_OK;
  }
  int retranspose(size_t first, size_t count) {
    const int W = cfg.width, H = cfg.height;
    size_t done = 0;
    while (done < count) {  // gridDim.z <= 65535
      const size_t c = count - done < 32768 ? count - done : 32768;
      dim3 g((W + 31) / 32, (H + 31) / 32, (unsigned)c), b(32, 8);
      transpose_kernel<<<g, b, 0, stream>>>(imgT + (first + done) * vstride, img + (first + done) * vstride, W, H, vstride);
      done += c;
    }
    AGFR_CUDA(cudaGetLastError());
    return AGF_OK;
  }
};

Handle* H_(agf_rappids* p) { return reinterpret_cast<Handle*>(p); }
const Handle* H_(const agf_rappids* p) { return reinterpret_cast<const Handle*>(p); }

}  // namespace

extern "C" {

int agf_rappids_cfg_default(int32_t width, int32_t height, agf_rappids_cfg* c) {
  if (!c || width <= 0 || height <= 0) return fail(AGF_EINVAL, "bad arguments");
  memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  c->depth_scale = 10.0 / 256.0;       // Simulator/Rappids_Simulator/main.cpp:121-122
  c->focal_length = width / 2.0;       // :360
  c->cx = width / 2.0;                 // :485-486
  c->cy = height / 2.0;
  c->true_radius = 0.116;              // :167-169
  c->planning_radius = 0.174;
  c->min_checking_dist = 0.5;
  c->min_thrust = 5;                   // DepthImagePlanner.cpp:44-52
  c->max_thrust = 30;
  c->max_angvel = 20;
  c->min_section_time = 0.02;
  c->max_velocity = 5;
  c->max_pyramids = AGF_RAPPIDS_MAX_PYRAMIDS;
  c->cost_kind = AGF_RAPPIDS_COST_DIRECTION;
  c->cost_vec[0] = 0;
  c->cost_vec[1] = 0;
  c->cost_vec[2] = 1;
  c->sample_min_x = 0.1 * width;       // DepthImagePlanner.hpp:334-352
  c->sample_max_x = 0.9 * width;
  c->sample_min_y = 0.1 * height;
  c->sample_max_y = 0.9 * height;
  c->sample_min_depth = 1.5;
  c->sample_max_depth = 3.0;
  c->sample_min_time = 2.0;
  c->sample_max_time = 3.0;
  c->math = AGF_MATH_FAST;
  c->device = -1;
  return AGF_OK;
}

int agf_rappids_create(const agf_rappids_cfg* cfg, size_t n, int32_t max_candidates, agf_rappids** out) {
  if (!out) return fail(AGF_EINVAL, "out is null");
  *out = nullptr;
  if (!cfg) return fail(AGF_EINVAL, "cfg is null");
  if (n == 0 || n > (size_t)1 << 30) return fail(AGF_EINVAL, "n_vehicles out of range");
  if (max_candidates <= 0) return fail(AGF_EINVAL, "max_candidates must be > 0");
  if (cfg->width < 8 || cfg->height < 8 || cfg->width > 4096 || cfg->height > 4096)
    return fail(AGF_EINVAL, "image size out of range");
  if (!(cfg->depth_scale > 0) || !(cfg->focal_length > 0) || !(cfg->min_checking_dist > 0) ||
      !(cfg->true_radius > 0) || !(cfg->planning_radius > 0))
    return fail(AGF_EINVAL, "camera / vehicle geometry must be positive");
  if (!(cfg->min_section_time > 0)) return fail(AGF_EINVAL, "min_section_time must be > 0");
  if (cfg->cost_kind != AGF_RAPPIDS_COST_DIRECTION && cfg->cost_kind != AGF_RAPPIDS_COST_GOAL)
    return fail(AGF_EINVAL, "bad cost_kind");
  if (cfg->math != AGF_MATH_PARITY && cfg->math != AGF_MATH_FAST) return fail(AGF_EINVAL, "bad math");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(AGF_ENODEVICE, "no CUDA device available: agrifly_b200 has no CPU execution path");
  }
  int dev = cfg->device;
  if (dev < 0) AGFR_CUDA(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(AGF_EINVAL, "device ordinal out of range");
  AGFR_CUDA(cudaSetDevice(dev));
  Handle* h = new (std::nothrow) Handle();
  if (!h) return fail(AGF_ENOMEM, "host allocation");
  h->cfg = *cfg;
  h->cfg.device = dev;
  if (h->cfg.max_pyramids > AGF_RAPPIDS_MAX_PYRAMIDS) {
    delete h;
    return fail(AGF_EUNSUPPORTED, "max_pyramids exceeds AGF_RAPPIDS_MAX_PYRAMIDS (device storage per vehicle)");
  }
  if (h->cfg.max_pyramids <= 0) h->cfg.max_pyramids = AGF_RAPPIDS_MAX_PYRAMIDS;  // "unlimited": see pyramid_cap_hit
  h->n = n;
  h->kcap = max_candidates;
  if (const char* fj = getenv("AGF_RAPPIDS_FRAME_JUMP")) h->frame_jump = atoi(fj);
  if (const char* sf = getenv("AGF_RAPPIDS_SHRINK_FOLD")) h->shrink_fold = atoi(sf) != 0;
  if (const char* cf = getenv("AGF_RAPPIDS_COOP_FACTOR")) h->coop_factor = (float)atof(cf);
  if (const char* cc = getenv("AGF_RAPPIDS_COOP_CAP")) h->coop_cap = (float)atof(cc);
  h->device = dev;
  const size_t npix = h->npix();
#define AGFR_ALLOC(ptr, bytes)                                  \
  do {                                                          \
    cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));       \
    if (e_ != cudaSuccess) {                                    \
      delete h;                                                 \
      return fail(AGF_ENOMEM, "cudaMalloc " #ptr, e_);          \
    }                                                           \
  } while (0)
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(AGF_ECUDA, "cudaStreamCreate", e);
  }
  const size_t nGR = (size_t)cfg->height * ((cfg->width + 31) / 32), nGC = (size_t)cfg->width * ((cfg->height + 31) / 32);
  h->vstride = (2 * npix + nGR + nGC + 63) / 64 * 64;
  AGFR_ALLOC(h->scene, n * h->vstride * sizeof(uint16_t));
  h->img = h->scene;
  h->imgT = h->scene + npix;
  h->gminR = h->scene + 2 * npix;
  h->gminC = h->scene + 2 * npix + nGR;
  AGFR_ALLOC(h->state, n * 12 * sizeof(double));
  AGFR_ALLOC(h->cands, n * (size_t)h->kcap * 4 * sizeof(double));
  AGFR_ALLOC(h->flags, n * (size_t)h->kcap);
  AGFR_ALLOC(h->ccost, n * (size_t)h->kcap * sizeof(double));
  AGFR_ALLOC(h->results, n * sizeof(agf_rappids_result));
  AGFR_ALLOC(h->pyr, n * (size_t)AGF_RAPPIDS_MAX_PYRAMIDS * AGF_RAPPIDS_PYRAMID_DOUBLES * sizeof(double));
  AGFR_ALLOC(h->stats, 8 * sizeof(double));
  AGFR_ALLOC(h->prims, n * 9 * sizeof(double));
  AGFR_ALLOC(h->next, sizeof(int));
  AGFR_ALLOC(h->order, n * sizeof(int));
  AGFR_ALLOC(h->nlong, 2 * sizeof(int));
  AGFR_ALLOC(h->work, n * sizeof(unsigned));
#undef AGFR_ALLOC
  // images start empty (everything at the far plane would be 65535; zero = "ignored" pixels), states zero with
  // the shared cost vector
  cudaMemsetAsync(h->scene, 0, n * h->vstride * sizeof(uint16_t), h->stream);
  cudaMemset2DAsync(h->gminR, h->vstride * sizeof(uint16_t), 0xFF, (nGR + nGC) * sizeof(uint16_t), n, h->stream);  // nothing seen
  cudaMemsetAsync(h->results, 0, n * sizeof(agf_rappids_result), h->stream);
  cudaMemsetAsync(h->flags, 0, n * (size_t)h->kcap, h->stream);
  {
    std::vector<double> st(n * 12, 0.0);
    for (size_t v = 0; v < n; v++) {
      st[v * 12 + 7] = 9.81;  // level camera: gravity along +y
      for (int a = 0; a < 3; a++) st[v * 12 + 9 + a] = cfg->cost_vec[a];
    }
    e = cudaMemcpyAsync(h->state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
      delete h;
      return fail(AGF_ECUDA, "state upload", e);
    }
  }
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  e = (h->cfg.math == AGF_MATH_PARITY) ? agfr::plan_blocks_per_sm_parity(&h->blocks_per_sm, &h->regs)
                                       : agfr::plan_blocks_per_sm_fast(&h->blocks_per_sm, &h->regs);
  if (e != cudaSuccess || h->blocks_per_sm < 1) {
    delete h;
    return fail(AGF_ECUDA, "planner kernel occupancy query", e);
  }
  // one resident wave of CTAs; warps fetch vehicles from a counter until the population is done
  h->grid = nsm * h->blocks_per_sm;
  const size_t need = (n + agfr::kWarps - 1) / agfr::kWarps;
  if ((size_t)h->grid > need) h->grid = (int)need;
  *out = reinterpret_cast<agf_rappids*>(h);
  return AGF_OK;
}

int agf_rappids_destroy(agf_rappids* p) {
  if (p) delete H_(p);
  return AGF_OK;
}
size_t agf_rappids_size(const agf_rappids* p) { return p ? H_(p)->n : 0; }
void* agf_rappids_stream(const agf_rappids* p) { return p ? (void*)H_(p)->stream : nullptr; }
uint64_t agf_rappids_launch_count(const agf_rappids* p) { return p ? H_(p)->launches : 0; }

int agf_rappids_set_images(agf_rappids* p, const uint16_t* images, size_t first, size_t count) {
  if (!p || !images) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  if (count)
    AGFR_CUDA(cudaMemcpy2DAsync(h->img + first * h->vstride, h->vstride * sizeof(uint16_t), images, h->npix() * sizeof(uint16_t),
                                h->npix() * sizeof(uint16_t), count, cudaMemcpyHostToDevice, h->stream));
  if (int rc = h->retranspose(first, count)) return rc;
  h->launches += 1;
  if (int rc = h->rebuild_groupmin(first, count)) return fail(rc, "group-minimum kernel launch");
  AGFR_CUDA(cudaStreamSynchronize(h->stream));  // the caller may reuse its buffer
  return AGF_OK;
}

int agf_rappids_render_scenes(agf_rappids* p, const uint16_t* row_bg, const int32_t* boxes, size_t first, size_t count) {
  if (!p || !row_bg || !boxes) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const int W = h->cfg.width, Hh = h->cfg.height;
  const size_t bg_bytes = count * Hh * sizeof(uint16_t), box_bytes = count * AGF_RAPPIDS_MAX_BOXES * 5 * sizeof(int32_t);
  const size_t box_off = (bg_bytes + 15) & ~(size_t)15;
  if (int rc = h->ensure_stage(box_off + box_bytes)) return rc;
  uint16_t* d_bg = reinterpret_cast<uint16_t*>(h->stage);
  int32_t* d_box = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(h->stage) + box_off);
  AGFR_CUDA(cudaMemcpyAsync(d_bg, row_bg, bg_bytes, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaMemcpyAsync(d_box, boxes, box_bytes, cudaMemcpyHostToDevice, h->stream));
  const size_t total = count * h->npix();
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  render_kernel<false><<<grid, 256, 0, h->stream>>>(h->img + first * h->vstride, d_bg, d_box, W, Hh, count, h->vstride);
  render_kernel<true><<<grid, 256, 0, h->stream>>>(h->imgT + first * h->vstride, d_bg, d_box, W, Hh, count, h->vstride);
  AGFR_CUDA(cudaGetLastError());
  h->launches += 2;
  if (int rc = h->rebuild_groupmin(first, count)) return fail(rc, "group-minimum kernel launch");
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_get_images(agf_rappids* p, uint16_t* images, size_t first, size_t count) {
  if (!p || !images) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  if (count)
    AGFR_CUDA(cudaMemcpy2DAsync(images, h->npix() * sizeof(uint16_t), h->img + first * h->vstride, h->vstride * sizeof(uint16_t),
                                h->npix() * sizeof(uint16_t), count, cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_set_states(agf_rappids* p, const double* vel0, const double* acc0, const double* grav, size_t first,
                           size_t count) {
  if (!p || !vel0 || !acc0 || !grav) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t pitch = 12 * sizeof(double), w = 3 * sizeof(double);
  char* base = reinterpret_cast<char*>(h->state + first * 12);
  AGFR_CUDA(cudaMemcpy2DAsync(base, pitch, vel0, w, w, count, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaMemcpy2DAsync(base + w, pitch, acc0, w, w, count, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaMemcpy2DAsync(base + 2 * w, pitch, grav, w, w, count, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_set_goals(agf_rappids* p, const double* goals, size_t first, size_t count) {
  if (!p || !goals) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t pitch = 12 * sizeof(double), w = 3 * sizeof(double);
  char* base = reinterpret_cast<char*>(h->state + first * 12);
  AGFR_CUDA(cudaMemcpy2DAsync(base + 3 * w, pitch, goals, w, w, count, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_set_candidates(agf_rappids* p, const double* candidates, int32_t k, size_t first, size_t count) {
  if (!p || !candidates) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (k <= 0 || k > h->kcap) return fail(AGF_EINVAL, "k outside (0, max_candidates]");
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t w = (size_t)k * 4 * sizeof(double);
  AGFR_CUDA(cudaMemcpy2DAsync(h->cands + first * h->kcap * 4, (size_t)h->kcap * 4 * sizeof(double), candidates, w, w,
                              count, cudaMemcpyHostToDevice, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  h->k = k;
  return AGF_OK;
}

int agf_rappids_sample_candidates(agf_rappids* p, int32_t k, uint64_t seed, uint64_t first_global_index) {
  if (!p) return fail(AGF_EINVAL, "null handle");
  Handle* h = H_(p);
  if (k <= 0 || k > h->kcap) return fail(AGF_EINVAL, "k outside (0, max_candidates]");
  AGFR_CUDA(cudaSetDevice(h->device));
  const agf_rappids_cfg& c = h->cfg;
  SampleBox B{c.sample_min_x, c.sample_max_x, c.sample_min_y, c.sample_max_y, c.sample_min_depth, c.sample_max_depth,
              c.sample_min_time, c.sample_max_time, c.cx, c.cy, c.focal_length};
  const size_t total = h->n * (size_t)k;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  sample_kernel<<<grid, 256, 0, h->stream>>>(h->cands, h->n, k, h->kcap, seed, first_global_index, B);
  AGFR_CUDA(cudaGetLastError());
  h->launches += 1;
  h->k = k;
  return AGF_OK;
}

int agf_rappids_get_candidates(agf_rappids* p, double* candidates, size_t first, size_t count) {
  if (!p || !candidates) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (h->k <= 0) return fail(AGF_EINVAL, "no candidates set");
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t w = (size_t)h->k * 4 * sizeof(double);
  AGFR_CUDA(cudaMemcpy2DAsync(candidates, w, h->cands + first * h->kcap * 4, (size_t)h->kcap * 4 * sizeof(double), w,
                              count, cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_plan(agf_rappids* p) {
  if (!p) return fail(AGF_EINVAL, "null handle");
  Handle* h = H_(p);
  if (h->k <= 0) return fail(AGF_EINVAL, "no candidates: call agf_rappids_set_candidates or agf_rappids_sample_candidates");
  AGFR_CUDA(cudaSetDevice(h->device));
  const agf_rappids_cfg& c = h->cfg;
  agfr::PlanParams P;
  P.img = h->img;
  P.imgT = h->imgT;
  P.gminR = h->gminR;
  P.gminC = h->gminC;
  P.vstride = h->vstride;
  P.GW = h->GW();
  P.GH = h->GH();
  P.state = h->state;
  P.cands = h->cands;
  P.flags = h->flags;
  P.ccost = h->ccost;
  P.results = h->results;
  P.pyramids = h->pyr;
  P.prims = h->prims;
  P.next = h->next;
  P.work = h->work;
  P.order = nullptr;
  P.nlong = nullptr;
  P.nextLong = h->nlong + 1;
  P.n = (int)h->n;
  P.k = h->k;
  P.kcap = h->kcap;
  P.W = c.width;
  P.H = c.height;
  P.scale = c.depth_scale;
  P.f = c.focal_length;
  P.cx = c.cx;
  P.cy = c.cy;
  P.rPlan = c.planning_radius;
  P.minDist = c.min_checking_dist;
  P.fminA = c.min_thrust;
  P.fmaxA = c.max_thrust;
  P.wmaxA = c.max_angvel;
  P.minSec = c.min_section_time;
  P.vmax = c.max_velocity;
  P.maxPyr = c.max_pyramids;
  P.costKind = c.cost_kind;
  P.shrinkFold = h->shrink_fold;
  P.frameJump = (c.width % 8 == 0 && c.height % 8 == 0) ? h->frame_jump : 0;  // frame_strip reads 16-byte vectors of pixel lines
  // the integer constants of InflatePyramid, evaluated in double on the host exactly as the reference does
  // (DepthImagePlanner.cpp:460,506,608)
  P.edgeOff = (int)(c.focal_length * c.true_radius / c.min_checking_dist);
  P.ignore = (int)(uint16_t)(c.true_radius / c.depth_scale);
  P.num = (int)(c.focal_length * c.planning_radius / c.depth_scale);
  if (h->events.empty()) {  // the whole ring at first use: the ring indices below run modulo EVENT_RING
    h->events.reserve(Handle::EVENT_RING);
    for (int k = 0; k < Handle::EVENT_RING; k++) {
      cudaEvent_t a, b;
      AGFR_CUDA(cudaEventCreate(&a));
      AGFR_CUDA(cudaEventCreate(&b));
      h->events.emplace_back(a, b);
    }
  }
  if (h->ev_count == size_t(Handle::EVENT_RING)) AGFR_CUDA(h->fold_oldest_event());
  AGFR_CUDA(cudaMemsetAsync(h->next, 0, sizeof(int), h->stream));
  auto& ev = h->events[(h->ev_head + h->ev_count) % Handle::EVENT_RING];
  h->ev_count++;
  AGFR_CUDA(cudaEventRecord(ev.first, h->stream));
  if (h->dispatch == 1 && h->have_work) {  // inside the timed region
    AGFR_CUDA(cudaMemsetAsync(h->nlong, 0, 2 * sizeof(int), h->stream));
    dispatch_order_kernel<<<1, 1024, 0, h->stream>>>(h->work, (int)h->n, h->order, h->nlong, h->coop_factor, h->coop_cap, h->grid * agfr::kWarps);
    AGFR_CUDA(cudaGetLastError());
    h->launches += 1;
    P.order = h->order;
    if (h->coop_factor > 0.0f) P.nlong = h->nlong;
  }
  cudaError_t e = (c.math == AGF_MATH_PARITY) ? agfr::launch_plan_parity(P, h->grid, h->stream)
                                              : agfr::launch_plan_fast(P, h->grid, h->stream);
  if (e != cudaSuccess) return fail(AGF_ECUDA, "planner kernel launch", e);
  AGFR_CUDA(cudaEventRecord(ev.second, h->stream));
  h->launches += 2;
  h->have_work = true;
  return AGF_OK;
}

int agf_rappids_set_dispatch(agf_rappids* p, int32_t mode) {
  if (!p) return fail(AGF_EINVAL, "null handle");
  if (mode != AGF_RAPPIDS_DISPATCH_INDEX && mode != AGF_RAPPIDS_DISPATCH_BY_LAST_WORK) return fail(AGF_EINVAL, "unknown dispatch mode");
  H_(p)->dispatch = mode;
  return AGF_OK;
}

int agf_rappids_get_plan_work(agf_rappids* p, uint32_t* cycles, size_t first, size_t count) {
  if (!p || !cycles) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  if (!h->have_work) return fail(AGF_EINVAL, "no plan yet");
  AGFR_CUDA(cudaSetDevice(h->device));
  AGFR_CUDA(cudaMemcpyAsync(cycles, h->work + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_sync(agf_rappids* p) {
  if (!p) return fail(AGF_EINVAL, "null handle");
  AGFR_CUDA(cudaSetDevice(H_(p)->device));
  AGFR_CUDA(cudaStreamSynchronize(H_(p)->stream));
  return AGF_OK;
}

int agf_rappids_get_results(agf_rappids* p, agf_rappids_result* out, size_t first, size_t count) {
  if (!p || !out) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  AGFR_CUDA(cudaMemcpyAsync(out, h->results + first, count * sizeof(agf_rappids_result), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_get_tracking_primitives(agf_rappids* p, double* records, size_t first, size_t count) {
  if (!p || !records) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  if (!count) return AGF_OK;
  AGFR_CUDA(cudaSetDevice(h->device));
  std::vector<double> abg(count * 9), st(count * 12);
  std::vector<agf_rappids_result> res(count);
  AGFR_CUDA(cudaMemcpyAsync(abg.data(), h->prims + first * 9, abg.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaMemcpyAsync(st.data(), h->state + first * 12, st.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaMemcpyAsync(res.data(), h->results + first, count * sizeof(agf_rappids_result), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < count; i++) {
    double* r = records + i * AGF_OFFTRAJ_DOUBLES;
    for (int a = 0; a < 3; a++) {
      r[6 * a + 0] = 0.0;                // trajectories start at the focal point
      r[6 * a + 1] = st[i * 12 + a];      // vel0
      r[6 * a + 2] = st[i * 12 + 3 + a];  // acc0
      r[6 * a + 3] = abg[i * 9 + 3 * a];
      r[6 * a + 4] = abg[i * 9 + 3 * a + 1];
      r[6 * a + 5] = abg[i * 9 + 3 * a + 2];
      r[18 + a] = st[i * 12 + 6 + a];     // gravity in the trajectory frame
      r[26 + a] = 0.0;
    }
    r[21] = res[i].found ? res[i].best_tf : 0.0;
    r[22] = 1.0; r[23] = r[24] = r[25] = 0.0;
  }
  return AGF_OK;
}

int agf_rappids_export_tracking_primitives(agf_rappids* p, double* dev_dst, size_t n_dst, size_t dst_first, const double* att,
                                           const double* offset) {
  if (!p || !dev_dst) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (dst_first + h->n > n_dst) return fail(AGF_ERANGE, "destination table too small for this planner's vehicles");
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t ab = att ? h->n * 4 * sizeof(double) : 0, ob = offset ? h->n * 3 * sizeof(double) : 0;
  double *d_att = nullptr, *d_off = nullptr;
  if (ab + ob) {
    if (int rc = h->ensure_stage(ab + ob)) return rc;
    if (att) {
      d_att = reinterpret_cast<double*>(h->stage);
      AGFR_CUDA(cudaMemcpyAsync(d_att, att, ab, cudaMemcpyHostToDevice, h->stream));
    }
    if (offset) {
      d_off = reinterpret_cast<double*>(reinterpret_cast<char*>(h->stage) + ab);
      AGFR_CUDA(cudaMemcpyAsync(d_off, offset, ob, cudaMemcpyHostToDevice, h->stream));
    }
  }
  export_prims_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, h->stream>>>(h->prims, h->state, h->results, d_att, d_off, h->n, dev_dst,
                                                                            n_dst, dst_first);
  AGFR_CUDA(cudaGetLastError());
  h->launches += 1;
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_get_candidate_flags(agf_rappids* p, uint8_t* flags, size_t first, size_t count) {
  if (!p || !flags) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (h->k <= 0) return fail(AGF_EINVAL, "no candidates set");
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  AGFR_CUDA(cudaMemcpy2DAsync(flags, (size_t)h->k, h->flags + first * h->kcap, (size_t)h->kcap, (size_t)h->k, count,
                              cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_get_pyramids(agf_rappids* p, double* pyramids, size_t first, size_t count) {
  if (!p || !pyramids) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = h->range(first, count)) return rc;
  AGFR_CUDA(cudaSetDevice(h->device));
  const size_t rec = (size_t)AGF_RAPPIDS_MAX_PYRAMIDS * AGF_RAPPIDS_PYRAMID_DOUBLES;
  AGFR_CUDA(cudaMemcpyAsync(pyramids, h->pyr + first * rec, count * rec * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_reduce_stats_device(agf_rappids* p, double* dev_out) {
  if (!p || !dev_out) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  AGFR_CUDA(cudaSetDevice(h->device));
  AGFR_CUDA(cudaMemsetAsync(dev_out, 0, 8 * sizeof(double), h->stream));
  const int grid = (int)((h->n + 255) / 256 < 148 * 4 ? (h->n + 255) / 256 : 148 * 4);
  stats_kernel<<<grid, 256, 0, h->stream>>>(h->results, h->n, dev_out);
  AGFR_CUDA(cudaGetLastError());
  h->launches += 1;
  return AGF_OK;
}

int agf_rappids_reduce_stats(agf_rappids* p, double* host_out) {
  if (!p || !host_out) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  if (int rc = agf_rappids_reduce_stats_device(p, h->stats)) return rc;
  AGFR_CUDA(cudaMemcpyAsync(host_out, h->stats, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  return AGF_OK;
}

int agf_rappids_plan_kernel_time(agf_rappids* p, double* ms, uint64_t* launches) {
  if (!p || !ms || !launches) return fail(AGF_EINVAL, "null argument");
  Handle* h = H_(p);
  AGFR_CUDA(cudaSetDevice(h->device));
  AGFR_CUDA(cudaStreamSynchronize(h->stream));
  while (h->ev_count) AGFR_CUDA(h->fold_oldest_event());
  *launches = h->timed_launches;
  *ms = h->timed_launches ? h->timed_ms / (double)h->timed_launches : 0.0;
  h->timed_ms = 0;
  h->timed_launches = 0;
  return AGF_OK;
}

}  // extern "C"
