// Candidate augmentations (augment.cu): plan of the linear view pipeline and the per-iteration random draws.
#pragma once
#include "common.cuh"

namespace bre {

constexpr int AUG_MAX_STEPS = 4;     // permutation steps (discrete_shift / flip) in config order
constexpr int AUG_MAX_BATCH = 64;    // per-image uniforms of the continuous shift
enum AugKind { AUG_SHIFT = 1, AUG_FLIP = 2 };

struct AugPlan {
  int n_steps;
  int kind[AUG_MAX_STEPS];
  float p0[AUG_MAX_STEPS];           // discrete_shift: lim ; flip: p
  int cs_enabled, cs_circular;       // continuous_shift (applied after the permutation steps)
  float cs_shift;
  const float* cj_scale;             // colour affine per (n, c): out = in * scale + shift (composite of all colorjitter steps), may be null
  const float* cj_shift;
  unsigned long long seed;
};
struct AugDraws {
  int o1[AUG_MAX_STEPS], o2[AUG_MAX_STEPS];   // discrete_shift: the two roll offsets ; flip: o1 = flipped?
  float sx[AUG_MAX_BATCH], sy[AUG_MAX_BATCH]; // continuous_shift: uniforms in [0, 1) per image (randgen[:, 0], randgen[:, 1])
};

int launch_aug_draw(const AugPlan& plan, const Scalars* sc, AugDraws* draws, int N, cudaStream_t s);
int launch_aug_view(const float* x, float* out, int N, int C, int H, int W, const AugPlan& plan, const AugDraws* draws, cudaStream_t s);
int launch_aug_pull(float* g, float* tmp, float* gx, int N, int C, int H, int W, const AugPlan& plan, const AugDraws* draws, cudaStream_t s);

}  // namespace bre
