// Small-batch linear layers (the classification head at batch 1-16): fprop / dgrad / wgrad of  out[n][co] = sum_k in[n][k] W[co][k]
// with the optional K-concatenated second source of the tangent sweeps.  At these sizes the contraction is a handful of
// matrix-vector products over a [Co][Ci] weight matrix (ResNet-18 head: 397 x 512 = 0.8 MB): HBM/L2-latency bound, no reuse to
// tile for -- the 64x64-tile implicit GEMM spent 16 us per launch on a single row of tiles (profiles/launches_r1_summary.txt:
// 7 % of a config-2 iteration in the five head launches).  fp32 throughout (bit-class parity with the fp32 SIMT back end: the
// same products, summed in a fixed order).
#include "igemm.cuh"

namespace bre {
namespace {

constexpr int LS_MAXN = 16;

// ---- fprop: one warp per output channel, all rows --------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) linear_small_fprop_kernel(GemmArgs a, int vec) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int Co = a.g.Co, Ci = a.g.Ci, N = a.g.N;
  if (warp >= Co) return;
  float acc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n] = 0.f;
  for (int s = 0; s < a.nsrc; ++s) {
    const float* __restrict__ w = a.wgt[s] + (long long)warp * Ci;
    const float* __restrict__ x = a.act[s];
    if (vec) {
      for (int k = lane * 4; k < Ci; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          if (n < N) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + n * a.x_sN + k));
            acc[n] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[n]))));
          }
        }
      }
    } else {
      for (int k = lane; k < Ci; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int n = 0; n < NB; ++n)
          if (n < N) acc[n] = fmaf(wv, __ldg(x + n * a.x_sN + k), acc[n]);
      }
    }
  }
  const float b = a.bias != nullptr ? __ldg(a.bias + warp) : 0.f;
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    const float t = warp_sum(acc[n]);
    if (lane == 0 && n < N) a.out[(long long)n * Co + warp] = t + b;
  }
}

// ---- dgrad: block = 32 input channels x 32 groups of output channels (one warp per group, lanes along ci: coalesced 128-byte
//      rows of W); every thread keeps four independent weight loads in flight; fixed-order reduction over the groups ----------------
template <int NB>
__global__ void __launch_bounds__(1024) linear_small_dgrad_kernel(GemmArgs a) {
  pdl_prologue();
  constexpr int G = NB >= 16 ? 8 : (NB >= 8 ? 16 : 32);   // groups of output channels = warps per block (17 KB of partials)
  __shared__ float part[G][NB][33];
  const int cx = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int ci = blockIdx.x * 32 + cx;
  const int Co = a.g.Co, Ci = a.g.Ci, N = a.g.N;
  float acc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n] = 0.f;
  if (ci < Ci) {
    for (int s = 0; s < a.nsrc; ++s) {
      const float* __restrict__ w = a.wgt[s];
      const float* __restrict__ dy = a.act[s];
      int co = grp;
      for (; co + 3 * G < Co; co += 4 * G) {
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(w + (long long)(co + u * G) * Ci + ci);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int n = 0; n < NB; ++n)
            if (n < N) acc[n] = fmaf(wv[u], __ldg(dy + (long long)n * Co + co + u * G), acc[n]);
        }
      }
      for (; co < Co; co += G) {
        const float wv = __ldg(w + (long long)co * Ci + ci);
#pragma unroll
        for (int n = 0; n < NB; ++n)
          if (n < N) acc[n] = fmaf(wv, __ldg(dy + (long long)n * Co + co), acc[n]);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NB; ++n) part[grp][n][cx] = acc[n];
  __syncthreads();
  for (int idx = threadIdx.x; idx < NB * 32; idx += 32 * G) {
    const int n = idx >> 5, c = idx & 31;
    const int cc = blockIdx.x * 32 + c;
    if (n >= N || cc >= Ci) continue;
    float t = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < G; ++g2) t += part[g2][n][c];
    float* o = a.out + (long long)n * a.x_sN + (long long)cc * a.x_sC;
    *o = a.accumulate ? *o + t : t;
  }
}

// ---- wgrad: one thread per weight element -------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) linear_small_wgrad_kernel(GemmArgs a) {
  pdl_prologue();
  const int Co = a.g.Co, Ci = a.g.Ci, N = a.g.N;
  const long long total = (long long)Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Ci), ci = (int)(i - (long long)co * Ci);
    float t = 0.f;
    for (int s = 0; s < a.nsrc; ++s) {
      const float* __restrict__ x = a.act[s];    // in  [N][Ci]
      const float* __restrict__ dy = a.wgt[s];   // dout [N][Co]
#pragma unroll
      for (int n = 0; n < NB; ++n)
        if (n < N) t = fmaf(__ldg(dy + (long long)n * Co + co), __ldg(x + n * a.x_sN + ci), t);
    }
    a.out[i] = a.accumulate ? a.out[i] + t : t;
  }
}

template <int NB>
int launch_nb(const GemmArgs& a, cudaStream_t stream) {
  const int Co = a.g.Co, Ci = a.g.Ci;
  if (a.mode == GEMM_FPROP) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    int vec = (Ci % 4 == 0) && (a.x_sN % 4 == 0);
    for (int s = 0; s < a.nsrc; ++s) vec = vec && al16(a.act[s]) && al16(a.wgt[s]);
    BRE_KLAUNCH((linear_small_fprop_kernel<NB>), ceil_div((long long)Co * 32, 256), 256, 0, stream, a, vec);
  } else if (a.mode == GEMM_DGRAD) {
    BRE_KLAUNCH((linear_small_dgrad_kernel<NB>), ceil_div(Ci, 32), 32 * (NB >= 16 ? 8 : (NB >= 8 ? 16 : 32)), 0, stream, a);
  } else {
    long long blocks = ((long long)Co * Ci + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    BRE_KLAUNCH((linear_small_wgrad_kernel<NB>), (int)blocks, 256, 0, stream, a);
  }
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool linear_small_supported(const GemmArgs& a) {
  const ConvGeom& g = a.g;
  return g.R == 1 && g.S == 1 && g.H == 1 && g.W == 1 && g.Ho == 1 && g.Wo == 1 && g.stride == 1 && g.pad == 0 && g.N >= 1 && g.N <= LS_MAXN &&
         a.x_sC == 1 && a.epi.kind == 0 && a.nsrc >= 1 && a.nsrc <= 2 && !(a.mode == GEMM_FPROP && a.accumulate);
}

int launch_linear_small(const GemmArgs& a, cudaStream_t stream) {
  const int N = a.g.N;
  if (N <= 1) return launch_nb<1>(a, stream);
  if (N <= 2) return launch_nb<2>(a, stream);
  if (N <= 4) return launch_nb<4>(a, stream);
  if (N <= 8) return launch_nb<8>(a, stream);
  return launch_nb<16>(a, stream);
}

}  // namespace bre
