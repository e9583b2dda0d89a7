// Shared helpers for the breaching_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace bre {

void set_error(const std::string& msg);

#define BRE_CUDA_CHECK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _err = (expr);                                                                      \
    if (_err != cudaSuccess) {                                                                      \
      ::bre::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_err) + " at " +       \
                       __FILE__ + ":" + std::to_string(__LINE__));                                  \
      return -2;                                                                                    \
    }                                                                                               \
  } while (0)

#define BRE_CHECK_LAUNCH()                                                                          \
  do {                                                                                              \
    cudaError_t _err = cudaGetLastError();                                                          \
    if (_err != cudaSuccess) {                                                                      \
      ::bre::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_err) + " at " +  \
                       __FILE__ + ":" + std::to_string(__LINE__));                                  \
      return -2;                                                                                    \
    }                                                                                               \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of a double; result valid in thread 0.  `scratch` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
  double r = 0.0;
  if (warp == 0) {
    r = lane < nwarps ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Device-resident scalar block shared by all stages of one iteration (no host round trips).
// ---------------------------------------------------------------------------------------------
struct Scalars {
  // matching reduction (objectives.py list reductions), accumulated in double
  double dot, nG, ng, sq, l1w;
  // objective pieces of the current evaluation
  double match, task_loss, tv, norm, di, feat;
  // coefficients of v = c1*g + c2*G + c3*w*sign(G-g)
  float c1, c2, c3, pad0;
  // trial state
  double fmin;           // minimal objective so far (optimization_based_attack.py:103,119-121)
  double last_objective;
  double grad_norm_sq;   // for grad clipping
  int it;                // iterations executed (0-based index of the *next* step)
  int recorded;          // len(stats["Trial_k_Val"])
  int stopped;           // non-finite objective seen (:131-133)
  int improved;          // scratch: this iteration improved fmin
};

}  // namespace bre
