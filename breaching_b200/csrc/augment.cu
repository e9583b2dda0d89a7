// Candidate augmentations on the device (reference attacks/auxiliaries/augmentations.py, applied in the closure
// at optimization_based_attack.py:149-153).  The candidate x passes through a short pipeline of *linear* views before the model sees
// it; the closure's gradient is pulled back through the transposed pipeline.  Supported steps (the ones that keep the shape):
//   discrete_shift  (Jitter :9-18)            torch.roll by two random offsets in [-lim, lim), one draw per forward, whole batch
//   flip            (Flip :57-64)             horizontal flip with probability p, one draw per forward
//   colorjitter     (ColorJitter :70-89)      (x - mean[n, c]) / std[n, c], constants drawn once per attacker
//   continuous_shift (RandomTransform :141-205) bilinear grid_sample (align_corners = True) on the grid
//                                             g(i, j) = (lin[j] + sx[n], lin[i] + sy[n]), lin = linspace(-1, 1, S), per-image random shifts
//                                             of at most shift / (S - 1); padding "circular" maps g -> ((g + 1) mod 1) - 1 as the
//                                             reference does (which samples the top-left quadrant twice per axis -- reproduced)
// Random draws come from Philox keyed by (seed, iteration, step), so the forward view and the transposed pull-back of one
// iteration see the same draws without any host round trip; `bre_augment_*` take the draws explicitly (parity tests).
#include <math.h>

#include "../../include/breaching_b200.h"
#include "common.cuh"
#include "augment.cuh"

namespace bre {
namespace {

__device__ __forceinline__ void philox4(uint64_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t (&out)[4]) {
  uint32_t ctr[4] = {a, b, c, 0x85EBCA6Bu};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = ctr[0]; out[1] = ctr[1]; out[2] = ctr[2]; out[3] = ctr[3];
}
__device__ __forceinline__ float u01(uint32_t v) { return ((float)(v >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// one thread block: the draws of this iteration -> AugDraws in global memory (read by the view / pull-back kernels)
__global__ void aug_draw_kernel(AugPlan plan, const Scalars* sc, AugDraws* draws, int N) {
  pdl_prologue();
  const int it = sc != nullptr ? sc->it : 0;
  for (int s = threadIdx.x; s < plan.n_steps; s += blockDim.x) {
    uint32_t r[4];
    philox4(plan.seed, (uint32_t)it, (uint32_t)s, 0u, r);
    if (plan.kind[s] == AUG_SHIFT) {
      const int lim = (int)plan.p0[s];
      draws->o1[s] = lim > 0 ? (int)(r[0] % (uint32_t)(2 * lim)) - lim : 0;   // randint(-lim, lim)
      draws->o2[s] = lim > 0 ? (int)(r[1] % (uint32_t)(2 * lim)) - lim : 0;
    } else if (plan.kind[s] == AUG_FLIP) {
      draws->o1[s] = u01(r[0]) < plan.p0[s] ? 1 : 0;
    }
  }
  // continuous shift: two uniforms per image
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    uint32_t r[4];
    philox4(plan.seed, (uint32_t)it, 0xC0FFEEu, (uint32_t)n, r);
    draws->sx[n] = u01(r[0]);
    draws->sy[n] = u01(r[1]);
  }
}

// index map of the permutation steps: output pixel (i, j) reads input pixel (y, x)
__device__ __forceinline__ void map_back(const AugPlan& plan, const AugDraws& d, int H, int W, int i, int j, int& y, int& x) {
  y = i; x = j;
  for (int s = plan.n_steps - 1; s >= 0; --s) {
    if (plan.kind[s] == AUG_FLIP) { if (d.o1[s]) x = W - 1 - x; }
    else if (plan.kind[s] == AUG_SHIFT) {         // out[i, j] = in[(i - o1) mod H, (j - o2) mod W]
      y = ((y - d.o1[s]) % H + H) % H;
      x = ((x - d.o2[s]) % W + W) % W;
    }
  }
}
__device__ __forceinline__ void map_forward(const AugPlan& plan, const AugDraws& d, int H, int W, int y, int x, int& i, int& j) {
  i = y; j = x;
  for (int s = 0; s < plan.n_steps; ++s) {
    if (plan.kind[s] == AUG_SHIFT) { i = ((i + d.o1[s]) % H + H) % H; j = ((j + d.o2[s]) % W + W) % W; }
    else if (plan.kind[s] == AUG_FLIP) { if (d.o1[s]) j = W - 1 - j; }
  }
}

// continuous shift: source coordinate of output index `o` along an axis of extent S, for the uniform u of this image
__device__ __forceinline__ void cs_coord(const AugPlan& plan, float u, int o, int S, int& i0, float& frac) {
  const float lin = S > 1 ? -1.f + 2.f * (float)o / (float)(S - 1) : -1.f;
  const float delta = plan.cs_shift / (float)(S - 1);
  float g = lin + (u - 0.5f) * 2.f * delta;
  if (plan.cs_circular) g = (g + 1.f) - floorf(g + 1.f) - 1.f;     // python's (g + 1) % 1 - 1
  const float pos = (g + 1.f) * 0.5f * (float)(S - 1);              // align_corners = True
  const float fl = floorf(pos);
  i0 = (int)fl;
  frac = pos - fl;
}

// forward view: permutation steps, then (optionally) the continuous shift, then the colour affine
__global__ void __launch_bounds__(256) aug_view_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C, int H, int W,
                                                       AugPlan plan, const AugDraws* __restrict__ draws) {
  pdl_prologue();
  const long long total = (long long)N * C * H * W;
  const AugDraws& d = *draws;   // read through the pointer (uniform addresses: served by the L1 / constant path)
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % W);
    long long t = e / W;
    const int i = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const float* src = x + ((long long)n * C + c) * H * W;
    float v;
    if (plan.cs_enabled) {   // the continuous shift is the outermost spatial step: sample the permuted image bilinearly
      int y0, x0; float fy, fx;
      cs_coord(plan, d.sy[n], i, H, y0, fy);
      cs_coord(plan, d.sx[n], j, W, x0, fx);
      v = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int yy = y0 + a, xx = x0 + b;
          const float w = (a ? fy : 1.f - fy) * (b ? fx : 1.f - fx);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            int sy, sx;
            map_back(plan, d, H, W, yy, xx, sy, sx);
            v = fmaf(w, src[(long long)sy * W + sx], v);
          }
        }
    } else {
      int sy, sx;
      map_back(plan, d, H, W, i, j, sy, sx);
      v = src[(long long)sy * W + sx];
    }
    if (plan.cj_scale != nullptr) v = fmaf(v, plan.cj_scale[n * C + c], plan.cj_shift[n * C + c]);
    out[e] = v;
  }
}

// pull-back without continuous shift: gx[y, x] = scale * g[forward(y, x)]
__global__ void __launch_bounds__(256) aug_pull_perm_kernel(const float* __restrict__ g, float* __restrict__ gx, int N, int C, int H, int W,
                                                            AugPlan plan, const AugDraws* __restrict__ draws) {
  pdl_prologue();
  const long long total = (long long)N * C * H * W;
  const AugDraws& d = *draws;   // read through the pointer (uniform addresses: served by the L1 / constant path)
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(e % W);
    long long t = e / W;
    const int yq = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    int i, j;
    map_forward(plan, d, H, W, yq, xq, i, j);
    float v = g[(((long long)n * C + c) * H + i) * W + j];
    if (plan.cj_scale != nullptr) v *= plan.cj_scale[n * C + c];
    gx[e] = v;
  }
}

// pull-back through the continuous shift, separable and deterministic (no atomics): first along x, then along y.
//   tmp[n, c, i, xx] = sum_j wx(j -> xx) g[n, c, i, j]          out[n, c, yy, xx] = sum_i wy(i -> yy) tmp[n, c, i, xx]
// (every thread walks the output index of its axis in order; O(S) per element, S <= a few hundred)
__global__ void __launch_bounds__(256) aug_pull_cs_kernel(const float* __restrict__ g, float* __restrict__ out, int N, int C, int H, int W,
                                                          AugPlan plan, const AugDraws* __restrict__ draws, int axis) {
  pdl_prologue();
  const long long total = (long long)N * C * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(e % W);
    long long t = e / W;
    const int yq = (int)(t % H); t /= H;
    const int n = (int)(t / C);
    const float* plane = g + (e - ((long long)yq * W + xq));
    const float u = axis == 0 ? draws->sx[n] : draws->sy[n];
    const int S = axis == 0 ? W : H, q = axis == 0 ? xq : yq;
    float acc = 0.f;
    for (int o = 0; o < S; ++o) {
      int i0; float fr;
      cs_coord(plan, u, o, S, i0, fr);
      float w = 0.f;
      if (i0 == q) w = 1.f - fr;
      else if (i0 + 1 == q) w = fr;
      if (w != 0.f) acc = fmaf(w, axis == 0 ? plane[(long long)yq * W + o] : plane[(long long)o * W + xq], acc);
    }
    out[e] = acc;
  }
}

inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int launch_aug_draw(const AugPlan& plan, const Scalars* sc, AugDraws* draws, int N, cudaStream_t s) {
  if (N > AUG_MAX_BATCH) { set_error("augmentations: batch larger than AUG_MAX_BATCH"); return -4; }
  BRE_KLAUNCH(aug_draw_kernel, 1, 64, 0, s, plan, sc, draws, N);
  BRE_CHECK_LAUNCH();
  return 0;
}
int launch_aug_view(const float* x, float* out, int N, int C, int H, int W, const AugPlan& plan, const AugDraws* draws, cudaStream_t s) {
  BRE_KLAUNCH(aug_view_kernel, grid_for((long long)N * C * H * W), 256, 0, s, x, out, N, C, H, W, plan, draws);
  BRE_CHECK_LAUNCH();
  return 0;
}
// gx <- pull-back of g (g is clobbered when the continuous shift is enabled: it serves as the intermediate); tmp: same size
int launch_aug_pull(float* g, float* tmp, float* gx, int N, int C, int H, int W, const AugPlan& plan, const AugDraws* draws, cudaStream_t s) {
  const int grid = grid_for((long long)N * C * H * W);
  const float* src = g;
  if (plan.cs_enabled) {
    BRE_KLAUNCH(aug_pull_cs_kernel, grid, 256, 0, s, (const float*)g, tmp, N, C, H, W, plan, draws, 0);
    BRE_KLAUNCH(aug_pull_cs_kernel, grid, 256, 0, s, (const float*)tmp, g, N, C, H, W, plan, draws, 1);
  }
  BRE_KLAUNCH(aug_pull_perm_kernel, grid, 256, 0, s, src, gx, N, C, H, W, plan, draws);
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace bre

using namespace bre;

// ---- stand-alone entry points with explicit draws (parity tests against the reference modules) -----------------------------------
static int fill_plan(AugPlan* plan, AugDraws* d, int32_t n_steps, const int32_t* kinds, const int32_t* o1, const int32_t* o2, float cs_shift,
                     int32_t cs_circular, const float* sx, const float* sy, int32_t N, const float* cj_scale, const float* cj_shift) {
  if (n_steps < 0 || n_steps > AUG_MAX_STEPS || N > AUG_MAX_BATCH) { set_error("bre_augment: too many steps / images"); return BRE_ERR_INVALID; }
  memset(plan, 0, sizeof(*plan));
  memset(d, 0, sizeof(*d));
  plan->n_steps = n_steps;
  for (int s = 0; s < n_steps; ++s) { plan->kind[s] = kinds[s]; d->o1[s] = o1[s]; d->o2[s] = o2 ? o2[s] : 0; }
  plan->cs_enabled = sx != nullptr; plan->cs_shift = cs_shift; plan->cs_circular = cs_circular;
  for (int n = 0; n < N && sx != nullptr; ++n) { d->sx[n] = sx[n]; d->sy[n] = sy[n]; }
  plan->cj_scale = cj_scale; plan->cj_shift = cj_shift;
  return 0;
}

extern "C" int bre_augment_view(const float* x, float* out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t n_steps, const int32_t* kinds,
                                const int32_t* o1, const int32_t* o2, float cs_shift, int32_t cs_circular, const float* sx, const float* sy,
                                const float* cj_scale, const float* cj_shift, int32_t transpose, float* scratch, void* stream) {
  if (!x || !out || N <= 0 || C <= 0 || H <= 0 || W <= 0) { set_error("bre_augment_view: bad arguments"); return BRE_ERR_INVALID; }
  AugPlan plan; AugDraws host;
  { const int frc = fill_plan(&plan, &host, n_steps, kinds, o1, o2, cs_shift, cs_circular, sx, sy, N, cj_scale, cj_shift); if (frc != 0) return frc; }
  cudaStream_t s = (cudaStream_t)stream;
  AugDraws* dev = nullptr;
  BRE_CUDA_CHECK(cudaMallocAsync((void**)&dev, sizeof(AugDraws), s));
  BRE_CUDA_CHECK(cudaMemcpyAsync(dev, &host, sizeof(AugDraws), cudaMemcpyHostToDevice, s));
  int rc = 0;
  if (!transpose) rc = launch_aug_view(x, out, N, C, H, W, plan, dev, s);
  else {
    if (plan.cs_enabled && !scratch) { set_error("bre_augment_view: the transposed continuous shift needs a scratch buffer"); rc = BRE_ERR_INVALID; }
    else rc = launch_aug_pull(const_cast<float*>(x), scratch, out, N, C, H, W, plan, dev, s);
  }
  cudaStreamSynchronize(s);
  cudaFreeAsync(dev, s);
  return rc;
}
