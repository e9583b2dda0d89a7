// Implicit-GEMM convolution family (fprop / dgrad / wgrad, optional K-concatenated second source).
//
// These are the dense contractions of the hot path: model forward (objectives.py:44), the first backward
// (objectives.py:45) and -- via the K-concatenated "dual source" form [W | v] . [a_dot ; a] -- both halves
// of the second backward (optimization_based_attack.py:165), see DESIGN.md section 3.
#pragma once
#include "common.cuh"

namespace bre {

enum GemmMode { GEMM_FPROP = 0, GEMM_DGRAD = 1, GEMM_WGRAD = 2 };

struct ConvGeom {
  int N, H, W, Ci;   // input-shaped tensor
  int Ho, Wo, Co;    // output-shaped tensor
  int R, S, stride, pad;
};

// Optional consumer fused into the FPROP epilogue of the tcgen05 back end: the BN + residual + ReLU op that follows a
// convolution (layers.cu bnact_fwd_kernel, kind 1) or its tangent (bnact_tan_fwd_kernel, kind 2).  The GEMM result is still
// written to `out` when `out` is non-null (the backward sweeps need the pre-BN value; nobody reads the pre-BN tangent).
struct GemmEpilogue {
  int kind;                 // 0 = none, 1 = value, 2 = tangent
  int has_bn, relu;
  int round_out;            // store out2 on the TF32 grid (it feeds tensor-core GEMMs, see tf32_rna)
  float* out2;              // [M][Nc] output of the fused op
  const float* res;         // residual branch (value / tangent), may be null
  const float* scale;       // alpha = gamma * invstd
  const float* shift;       // kind 1: beta - mean * alpha
  const float* inv;         // kind 2: invstd, -mean * invstd  (xhat = pre * inv + nrm)
  const float* nrm;
  const float* v_gamma;     // kind 2: direction components of gamma / beta
  const float* v_beta;
  const float* pre;         // kind 2: forward pre-BN value and post-activation value (ReLU mask)
  const float* post;
};

struct GemmArgs {
  int mode;
  GemmEpilogue epi;
  ConvGeom g;
  int nsrc;               // 1 or 2 (dual source: K-concatenation, fprop / dgrad only)
  const float* act[2];    // fprop: in ; dgrad: dout ; wgrad: in
  const float* wgt[2];    // fprop/dgrad: weights OHWI ; wgrad: dout
  // addressing of the input-shaped tensor (gathered operand of fprop/wgrad, *output* of dgrad):
  // offset(img, pixel, c) = img * x_sN + pixel * x_sP + c * x_sC   (NHWC: HWC, C, 1 ; NCHW: CHW, 1, HW)
  long long x_sN;
  int x_sP, x_sC;
  float* out;
  const float* bias;      // fprop only, per output channel (may be null)
  int accumulate;         // out += result
  float* ws;              // split-K workspace (>= ws_tiles * 4096 floats)
  int* counters;          // >= max tiles ints, zero-initialised, self-resetting
  int ws_tiles;
  int splits;             // 0 = choose automatically
  int force_fp32;         // engine: run this contraction on the fp32 kernels even on the tensor-core back end (precision knob)
  // bit s: wgt[s] (fprop / dgrad) was fully written before the *predecessor* kernel of this launch started, so the tcgen05 back end
  // may start loading it before griddepcontrol.wait (model weights; the direction v once a serialised launch follows make_v)
  unsigned wgt_static;
};

constexpr int IG_BM = 64, IG_BN = 64, IG_BK = 16, IG_THREADS = 256;

int launch_igemm_simt(const GemmArgs& a, cudaStream_t stream);
// linear layers on <= 16 rows (the classification head at small batch): dedicated fp32 kernels (linear_small.cu)
bool linear_small_supported(const GemmArgs& a);
int launch_linear_small(const GemmArgs& a, cudaStream_t stream);
bool linear_small_preferred(const GemmArgs& a);   // engine dispatch: small-row linear with a short reduction -> these kernels, not the GEMM
// dgrad of a linear layer with <= 32 rows, <= 128 inputs and >= 8192 outputs (the token models' decoder): chunked reduction in
// fp32 registers + a fixed-order fold (linear_small.cu); uses a.ws for the per-chunk partial sums
bool linear_tall_supported(const GemmArgs& a);
int launch_linear_tall(const GemmArgs& a, cudaStream_t stream);
// tcgen05 TF32 back end (igemm_tc.cu); returns BRE_ERR_UNSUPPORTED (-4) for shapes it does not cover.
int launch_igemm_tc(const GemmArgs& a, cudaStream_t stream);
bool igemm_tc_supported(const GemmArgs& a);

inline void gemm_dims(const GemmArgs& a, int& M, int& Nc, int& K) {
  const ConvGeom& g = a.g;
  if (a.mode == GEMM_FPROP) { M = g.N * g.Ho * g.Wo; Nc = g.Co; K = g.R * g.S * g.Ci; }
  else if (a.mode == GEMM_DGRAD) { M = g.N * g.H * g.W; Nc = g.Ci; K = g.R * g.S * g.Co; }
  else { M = g.Co; Nc = g.R * g.S * g.Ci; K = g.N * g.Ho * g.Wo; }
}

}  // namespace bre
