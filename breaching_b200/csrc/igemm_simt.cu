// fp32 SIMT implicit-GEMM back end (CUDA cores).  Bit-faithful fp32 arithmetic: this is the parity
// back end and the fallback for shapes the tcgen05 TF32 back end (igemm_tc.cu) does not cover
// (tiny channel counts such as the 3-channel stem, ragged K).
//
// Tile 64x64x16, 256 threads, 4x4 outputs per thread, register-prefetch double buffering, deterministic
// split-K (partials to a workspace, the last-arriving CTA of a tile reduces them in fixed order).
#include "igemm.cuh"

namespace bre {

namespace {

struct Dims {
  int M, Nc, K;
  int steps_per_src, total_steps, steps_per_split;
  int vecA, vecB, vecOut;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int MODE>
__global__ void __launch_bounds__(IG_THREADS) igemm_simt_kernel(GemmArgs a, Dims d) {
  __shared__ __align__(16) float As[2][IG_BK][IG_BM + 4];
  __shared__ __align__(16) float Bs[2][IG_BK][IG_BN + 4];
  __shared__ int s_last;

  const ConvGeom g = a.g;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * IG_BM, n0 = blockIdx.y * IG_BN;
  const int z = blockIdx.z;
  const int s_begin = z * d.steps_per_split;
  const int s_end = min(d.total_steps, s_begin + d.steps_per_split);

  // thread mappings of the two loader flavours
  const int kc_row = tid >> 2, kc_k = (tid & 3) << 2;    // K-contiguous operand: 1 row, 4 consecutive k
  const int mc_k = tid >> 4, mc_col = (tid & 15) << 2;   // M/N-contiguous operand: 1 k, 4 consecutive rows
  const int HoWo = g.Ho * g.Wo, HW = g.H * g.W;

  // ---- per-thread fixed decode -----------------------------------------------------------------
  // FPROP A / DGRAD A: one row (pixel) per thread
  bool a_valid = false;
  int a_img = 0, a_y = 0, a_x = 0;   // FPROP: (h0, w0) = top-left of the window ; DGRAD: (h, w) of the input pixel
  if (MODE == GEMM_FPROP) {
    const int m = m0 + kc_row;
    a_valid = m < d.M;
    if (a_valid) {
      a_img = m / HoWo;
      const int rem = m - a_img * HoWo;
      const int p = rem / g.Wo, q = rem - p * g.Wo;
      a_y = p * g.stride - g.pad;
      a_x = q * g.stride - g.pad;
    }
  } else if (MODE == GEMM_DGRAD) {
    const int m = m0 + kc_row;
    a_valid = m < d.M;
    if (a_valid) {
      a_img = m / HW;
      const int rem = m - a_img * HW;
      a_y = rem / g.W;
      a_x = rem - a_y * g.W;
    }
  }
  // WGRAD B: 4 consecutive columns n = (rs, c)
  int wb_c = 0, wb_r = 0, wb_s = 0;
  if (MODE == GEMM_WGRAD) {
    const int n = n0 + mc_col;
    if (n < d.Nc) {
      const int rs = n / g.Ci;
      wb_c = n - rs * g.Ci;
      wb_r = rs / g.S;
      wb_s = rs - wb_r * g.S;
    }
  }

  float ra[4], rb[4];

  auto load_tiles = [&](int step) {
    const int src = step / d.steps_per_src;
    const int kbase = (step - src * d.steps_per_src) * IG_BK;
    const float* __restrict__ act = a.act[src];
    const float* __restrict__ wgt = a.wgt[src];
#pragma unroll
    for (int j = 0; j < 4; ++j) { ra[j] = 0.f; rb[j] = 0.f; }

    if (MODE == GEMM_FPROP) {
      // A(m, k) = in[img, h0 + r, w0 + s, c],  k = (r, s, c)
      const int kk = kbase + kc_k;
      if (a_valid && kk < d.K) {
        if (d.vecA) {
          const int rs = kk / g.Ci, c = kk - rs * g.Ci;
          const int r = rs / g.S, s = rs - r * g.S;
          const int h = a_y + r, w = a_x + s;
          if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
            const float4 v = ldg4(act + a_img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + c);
            ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kj = kk + j;
            if (kj < d.K) {
              const int rs = kj / g.Ci, c = kj - rs * g.Ci;
              const int r = rs / g.S, s = rs - r * g.S;
              const int h = a_y + r, w = a_x + s;
              if (h >= 0 && h < g.H && w >= 0 && w < g.W)
                ra[j] = __ldg(act + a_img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + (long long)c * a.x_sC);
            }
          }
        }
      }
      // B(n, k) = W[n][k]
      const int n = n0 + kc_row;
      if (n < d.Nc && kk < d.K) {
        const float* wp = wgt + (long long)n * d.K + kk;
        if (d.vecB) {
          const float4 v = ldg4(wp);
          rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kk + j < d.K) rb[j] = __ldg(wp + j);
        }
      }
    } else if (MODE == GEMM_DGRAD) {
      // A(m, k) = dout[img, (h + pad - r)/stride, (w + pad - s)/stride, ko],  k = (r, s, ko)
      const int kk = kbase + kc_k;
      if (a_valid && kk < d.K) {
        if (d.vecA) {
          const int rs = kk / g.Co, ko = kk - rs * g.Co;
          const int r = rs / g.S, s = rs - r * g.S;
          const int hp = a_y + g.pad - r, wp = a_x + g.pad - s;
          if (hp >= 0 && wp >= 0) {
            const int p = hp / g.stride, q = wp / g.stride;
            if (p * g.stride == hp && q * g.stride == wp && p < g.Ho && q < g.Wo) {
              const float4 v = ldg4(act + ((long long)(a_img * g.Ho + p) * g.Wo + q) * g.Co + ko);
              ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kj = kk + j;
            if (kj < d.K) {
              const int rs = kj / g.Co, ko = kj - rs * g.Co;
              const int r = rs / g.S, s = rs - r * g.S;
              const int hp = a_y + g.pad - r, wp = a_x + g.pad - s;
              if (hp >= 0 && wp >= 0) {
                const int p = hp / g.stride, q = wp / g.stride;
                if (p * g.stride == hp && q * g.stride == wp && p < g.Ho && q < g.Wo)
                  ra[j] = __ldg(act + ((long long)(a_img * g.Ho + p) * g.Wo + q) * g.Co + ko);
              }
            }
          }
        }
      }
      // B(n = ci, k) = W[ko][r][s][ci]
      const int k1 = kbase + mc_k;
      const int n = n0 + mc_col;
      if (k1 < d.K && n < d.Nc) {
        const int rs = k1 / g.Co, ko = k1 - rs * g.Co;
        const float* wp = wgt + ((long long)ko * (g.R * g.S) + rs) * g.Ci + n;
        if (d.vecB) {
          const float4 v = ldg4(wp);
          rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < d.Nc) rb[j] = __ldg(wp + j);
        }
      }
    } else {  // GEMM_WGRAD
      const int k1 = kbase + mc_k;  // pixel (img, p, q)
      if (k1 < d.K) {
        // A(m = ko, k) = dout[pixel][ko]
        const int m = m0 + mc_col;
        if (m < d.M) {
          const float* dp = wgt + (long long)k1 * g.Co + m;
          if (d.vecA) {
            const float4 v = ldg4(dp);
            ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (m + j < d.M) ra[j] = __ldg(dp + j);
          }
        }
        // B(n = (r, s, c), k) = in[img, p*stride - pad + r, q*stride - pad + s, c]
        const int n = n0 + mc_col;
        if (n < d.Nc) {
          const int img = k1 / HoWo;
          const int rem = k1 - img * HoWo;
          const int p = rem / g.Wo, q = rem - p * g.Wo;
          const int hb = p * g.stride - g.pad, wb = q * g.stride - g.pad;
          if (d.vecB) {
            const int h = hb + wb_r, w = wb + wb_s;
            if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
              const float4 v = ldg4(act + img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + wb_c);
              rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
            }
          } else {
            int c = wb_c, r = wb_r, s = wb_s;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j < d.Nc) {
                const int h = hb + r, w = wb + s;
                if (h >= 0 && h < g.H && w >= 0 && w < g.W)
                  rb[j] = __ldg(act + img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + (long long)c * a.x_sC);
              }
              if (++c == g.Ci) { c = 0; if (++s == g.S) { s = 0; ++r; } }
            }
          }
        }
      }
    }
  };

  auto store_tiles = [&](int buf) {
    if (MODE == GEMM_FPROP) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[buf][kc_k + j][kc_row] = ra[j];
        Bs[buf][kc_k + j][kc_row] = rb[j];
      }
    } else if (MODE == GEMM_DGRAD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][kc_k + j][kc_row] = ra[j];
      *reinterpret_cast<float4*>(&Bs[buf][mc_k][mc_col]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
      *reinterpret_cast<float4*>(&As[buf][mc_k][mc_col]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(&Bs[buf][mc_k][mc_col]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (s_begin < s_end) {
    load_tiles(s_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int step = s_begin; step < s_end; ++step) {
    const int buf = (step - s_begin) & 1;
    const bool more = step + 1 < s_end;
    if (more) load_tiles(step + 1);
#pragma unroll
    for (int kk = 0; kk < IG_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- split-K: deterministic reduction by the last-arriving CTA of this output tile --------------
  const int splits = gridDim.z;
  if (splits > 1) {
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    float* wsb = a.ws + ((long long)tile * splits + z) * (IG_BM * IG_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(wsb + (ty * 4 + i) * IG_BN + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(a.counters + tile, 1);
      s_last = (prev == splits - 1);
      if (s_last) a.counters[tile] = 0;  // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* wst = a.ws + (long long)tile * splits * (IG_BM * IG_BN);
    for (int zz = 0; zz < splits; ++zz) {
      const float* p = wst + (long long)zz * (IG_BM * IG_BN);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p + (ty * 4 + i) * IG_BN + tx * 4));
        acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
      }
    }
  }

  // ---- epilogue ----------------------------------------------------------------------------------
  const int n = n0 + tx * 4;
  if (n >= d.Nc) return;
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE == GEMM_FPROP && a.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < d.Nc) bias4[j] = __ldg(a.bias + n + j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
    long long row;
    int cs = 1;
    if (MODE == GEMM_DGRAD) {
      const int img = m / HW;
      row = img * a.x_sN + (long long)(m - img * HW) * a.x_sP;
      cs = a.x_sC;
    } else {
      row = (long long)m * d.Nc;
    }
    float* op = a.out + row + (long long)n * cs;
    if (d.vecOut) {
      float4 v = make_float4(acc[i][0] + bias4[0], acc[i][1] + bias4[1], acc[i][2] + bias4[2], acc[i][3] + bias4[3]);
      if (a.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(op);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(op) = v;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < d.Nc) {
          float v = acc[i][j] + bias4[j];
          float* q = op + (long long)j * cs;
          if (a.accumulate) v += *q;
          *q = v;
        }
      }
    }
  }
}

// ---- dgrad onto a tensor with very few channels (the 3-channel candidate behind the stem conv) ----------------
// The GEMM formulation would run an N=64-wide tile for Ci=3 columns and multiply stride^2-1 out of stride^2 taps by
// structural zeros.  Here one thread owns one input pixel and all CI channels, walks only the taps that hit the
// output grid, and streams dout with 128-bit loads; lanes of a warp are mapped to pixels of the same stride-parity
// class so the tap loop is warp-uniform and the weight loads are broadcasts.
template <int CI>
__global__ void __launch_bounds__(128) dgrad_small_ci_kernel(GemmArgs a, int vec) {
  const ConvGeom g = a.g;
  const int img = blockIdx.z, h = blockIdx.y;
  const int Wc = (g.W + g.stride - 1) / g.stride;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int cls = slot / Wc, w = (slot - cls * Wc) * g.stride + cls;
  if (cls >= g.stride || w >= g.W) return;
  float acc[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) acc[c] = 0.f;
  const int RS = g.R * g.S;
  for (int src = 0; src < a.nsrc; ++src) {
    const float* __restrict__ dout = a.act[src];
    const float* __restrict__ wgt = a.wgt[src];
    for (int r = 0; r < g.R; ++r) {
      const int hp = h + g.pad - r;
      if (hp < 0) break;
      const int p = hp / g.stride;
      if (p * g.stride != hp || p >= g.Ho) continue;
      for (int s = 0; s < g.S; ++s) {
        const int wp = w + g.pad - s;
        if (wp < 0) break;
        const int q = wp / g.stride;
        if (q * g.stride != wp || q >= g.Wo) continue;
        const float* dp = dout + ((long long)(img * g.Ho + p) * g.Wo + q) * g.Co;
        const float* wq = wgt + (long long)(r * g.S + s) * g.Ci;
        if (vec) {
          for (int ko = 0; ko < g.Co; ko += 4) {
            const float4 dv = __ldg(reinterpret_cast<const float4*>(dp + ko));
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* wj = wq + (long long)(ko + j) * RS * g.Ci;
#pragma unroll
              for (int c = 0; c < CI; ++c) acc[c] = fmaf(dd[j], __ldg(wj + c), acc[c]);
            }
          }
        } else {
          for (int ko = 0; ko < g.Co; ++ko) {
            const float dd = __ldg(dp + ko);
            const float* wj = wq + (long long)ko * RS * g.Ci;
#pragma unroll
            for (int c = 0; c < CI; ++c) acc[c] = fmaf(dd, __ldg(wj + c), acc[c]);
          }
        }
      }
    }
  }
  float* op = a.out + img * a.x_sN + (long long)(h * g.W + w) * a.x_sP;
#pragma unroll
  for (int c = 0; c < CI; ++c) {
    float* q = op + (long long)c * a.x_sC;
    *q = a.accumulate ? *q + acc[c] : acc[c];
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_dgrad_small_ci(const GemmArgs& a, cudaStream_t stream) {
  const ConvGeom& g = a.g;
  bool vec = g.Co % 4 == 0;
  for (int s = 0; s < a.nsrc; ++s) vec = vec && aligned16(a.act[s]);
  const int Wc = (g.W + g.stride - 1) / g.stride;
  dim3 grid(ceil_div((long long)Wc * g.stride, 128), g.H, g.N), block(128);
  switch (g.Ci) {
    case 1: dgrad_small_ci_kernel<1><<<grid, block, 0, stream>>>(a, vec); break;
    case 2: dgrad_small_ci_kernel<2><<<grid, block, 0, stream>>>(a, vec); break;
    case 3: dgrad_small_ci_kernel<3><<<grid, block, 0, stream>>>(a, vec); break;
    default: dgrad_small_ci_kernel<4><<<grid, block, 0, stream>>>(a, vec); break;
  }
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int launch_igemm_simt(const GemmArgs& a, cudaStream_t stream) {
  Dims d;
  gemm_dims(a, d.M, d.Nc, d.K);
  const ConvGeom& g = a.g;
  if (d.M <= 0 || d.Nc <= 0 || d.K <= 0) { set_error("igemm: empty problem"); return -1; }
  if (a.nsrc < 1 || a.nsrc > 2 || (a.nsrc == 2 && a.mode == GEMM_WGRAD)) { set_error("igemm: bad nsrc"); return -1; }
  if (a.mode == GEMM_DGRAD && g.Ci <= 4 && g.H <= 65535 && g.N <= 65535) return launch_dgrad_small_ci(a, stream);
  d.steps_per_src = ceil_div(d.K, IG_BK);
  d.total_steps = d.steps_per_src * a.nsrc;

  bool ptr_ok = true;
  for (int s = 0; s < a.nsrc; ++s) ptr_ok = ptr_ok && aligned16(a.act[s]) && aligned16(a.wgt[s]);
  const bool x_vec = (a.x_sC == 1) && (g.Ci % 4 == 0) && (a.x_sP % 4 == 0) && (a.x_sN % 4 == 0);
  if (a.mode == GEMM_FPROP) {
    d.vecA = ptr_ok && x_vec;
    d.vecB = ptr_ok && (d.K % 4 == 0);
    d.vecOut = aligned16(a.out) && (d.Nc % 4 == 0);
  } else if (a.mode == GEMM_DGRAD) {
    d.vecA = ptr_ok && (g.Co % 4 == 0);
    d.vecB = ptr_ok && (g.Ci % 4 == 0);
    d.vecOut = aligned16(a.out) && x_vec;
  } else {
    d.vecA = ptr_ok && (g.Co % 4 == 0);
    d.vecB = ptr_ok && x_vec;
    d.vecOut = aligned16(a.out) && (d.Nc % 4 == 0);
  }

  const int tm = ceil_div(d.M, IG_BM), tn = ceil_div(d.Nc, IG_BN);
  const long long tiles = (long long)tm * tn;
  int splits = a.splits;
  if (splits <= 0) {
    splits = 1;
    if (tiles < 2 * kNumSMs) {
      splits = ceil_div(2 * kNumSMs, tiles);
      const int max_by_k = d.total_steps / 4 > 0 ? d.total_steps / 4 : 1;
      if (splits > max_by_k) splits = max_by_k;
    }
  }
  if (splits > d.total_steps) splits = d.total_steps;
  if (splits > 1 && (a.ws == nullptr || a.counters == nullptr)) splits = 1;
  if (splits > 1 && tiles * splits > a.ws_tiles) splits = (int)(a.ws_tiles / tiles) > 1 ? (int)(a.ws_tiles / tiles) : 1;
  d.steps_per_split = ceil_div(d.total_steps, splits);
  splits = ceil_div(d.total_steps, d.steps_per_split);  // no empty splits
  if (tn > 65535 || splits > 65535) { set_error("igemm: grid too large"); return -1; }

  dim3 grid(tm, tn, splits), block(IG_THREADS);
  if (a.mode == GEMM_FPROP) igemm_simt_kernel<GEMM_FPROP><<<grid, block, 0, stream>>>(a, d);
  else if (a.mode == GEMM_DGRAD) igemm_simt_kernel<GEMM_DGRAD><<<grid, block, 0, stream>>>(a, d);
  else igemm_simt_kernel<GEMM_WGRAD><<<grid, block, 0, stream>>>(a, d);
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace bre
