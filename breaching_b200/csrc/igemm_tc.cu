// tcgen05 TF32 implicit-GEMM back end (sm_100a): the dense contractions of the four sweeps on the 5th-generation
// tensor cores.  D[128 x BN] accumulates in TMEM (fp32), one elected thread issues tcgen05.mma.kind::tf32 with
// both operands read from shared memory through UMMA descriptors (128-byte-swizzled K-major or MN-major layouts, as the
// global layout of the operand dictates, so no transposed copies of activations / weights are ever materialised),
// completion is tracked with tcgen05.commit -> mbarrier, the epilogue reads the accumulator back with tcgen05.ld
// (32 lanes x 32 columns per warp).  Operand staging, TC_STAGES-deep ring, two producers:
//   * TMA (default wherever the geometry allows): one thread issues cp.async.bulk.tensor loads -- im2col-mode tensor
//     maps for the gathered activation operand (the hardware walks 128 output pixels x 32 channels of one filter tap,
//     zero-filling the padding halo), tiled maps for the weight operand -- which land directly in the swizzled UMMA
//     layouts and complete on the stage's mbarrier (expect_tx);
//   * cp.async (strided dgrad, whose "every stride-th tap" gather no tensor map expresses, and BRE_TC_TMA=0): four loader
//     warps issue 16-byte LDGSTS with precomputed per-row tap masks, completion via cp.async.mbarrier.arrive.
// Split-K runs inside a thread-block cluster (<= 16 CTAs along z) with a deterministic DSMEM reduction.
//
// fp32 storage everywhere; TF32 (10-bit mantissa) multiplies with fp32 accumulation -- the numeric mode cuDNN uses
// for the reference's GPU path by default (torch.backends.cudnn.allow_tf32).
#include <stdlib.h>
#include <string.h>

#include <array>
#include <map>
#include <mutex>
#include <type_traits>

#include <cuda.h>

#include "igemm.cuh"

namespace bre {

namespace {

constexpr int TC_BM = 128;      // UMMA M
constexpr int TC_BK = 32;       // k-block per pipeline stage (4 MMAs of K = 8)
constexpr int TC_STAGES = 4;       // default ring depth
constexpr int TC_MAX_STAGES = 8;   // deep ring (chosen per launch, TcDims::stage_shift): stage / phase of k-block i are i & (n - 1), (i >> log2 n) & 1
constexpr int TC_THREADS = 128;       // loader / epilogue threads (warps 0-3 <-> TMEM lane quadrants)
constexpr int TC_BLOCK = TC_THREADS + 32;  // + one MMA-issuer warp

struct TcDims {
  int M, Nc, K;
  int kblocks_per_src, total_kblocks, kblocks_per_split;
  int stage_shift;   // log2 of the ring depth of this launch (2 or 3)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();  // never hang the GPU: a lost arrival becomes a launch error
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address, leading / stride byte
// offsets in 16-byte units, version = 1 (Blackwell), layout type 0 = no swizzle ("interleave").
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = (uint64_t)(layout_type & 7u) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Shared-memory layouts of a [ROWS x 32] fp32 operand tile (ROWS = extent along M or N), as the tensor core reads them.
// Both were verified on the B200 with profiles/experiments/umma_layout_probe.cu (one-hot probing of every byte offset):
//   K-major, SWIZZLE_128B (layout type 2): one 128-byte smem row per operand row = the whole 32-wide k-block, its eight
//   16-byte granules XOR-swizzled with (row % 8), 8-row groups at SBO = 1024:
//       byte(row, k) = (row/8) * 1024 + (row%8) * 128 + (((k/4) ^ (row%8)) * 16) + (k%4) * 4
//       MMA covering k in [8j, 8j+8): start + 32 j (the swizzle is a function of the absolute address), LBO = 16
//   (the no-swizzle K-major layout also works -- first version of this kernel -- but forces either uncoalesced global
//   reads or 8-way conflicting shared stores; 128B swizzle gives coalesced 128-byte reads AND conflict-free stores)
//   MN-major for 32-bit operands only exists as SWIZZLE_128B_BASE32B (layout type 1; the plain / 16-byte-atom MN-major
//   layouts produce zeros for kind::tf32): 32 elements (128 B) contiguous along MN, rows of 128 B along k, the four
//   32-byte chunks of a row XOR-swizzled with (k % 4), k-groups of 4 at SBO, MN-groups of 32 at LBO:
//       byte(row, k) = (row/32) * 4096 + (k/4) * 512 + (k%4) * 128 + ((((row%32)/8) ^ (k%4)) * 32) + (row%8) * 4
//       MMA covering k in [8j, 8j+8): start + j*1024, LBO = 4096, SBO = 512
// K-major SWIZZLE_128B: granule = 16 bytes (4 k), `g` = granule column 0..7 of the 32-wide k-block
__device__ __forceinline__ uint32_t off_k128(int row, int g) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((g ^ (row & 7)) << 4));
}
template <int ROWS>
__device__ __forceinline__ uint32_t off_mnmajor(int row, int k) {
  return (uint32_t)((row >> 5) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((((row >> 3) & 3) ^ (k & 3)) << 5) + (row & 7) * 4);
}

// 16-byte asynchronous global -> shared copy; src_bytes = 0 zero-fills the destination (padding, out-of-range taps)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const float* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// arrive on `bar` once all cp.async issued so far by this thread have landed (counts against the barrier's init count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void loader_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte load from the shared memory of CTA `rank` of this cluster at the same offset as local address `saddr`
__device__ __forceinline__ float4 ld_dsmem4(uint32_t saddr, uint32_t rank) {
  uint32_t raddr;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(saddr), "r"(rank));
  float4 v;
  // volatile (not reordered across the cluster barriers, themselves volatile) but no memory clobber: a batch of these is issued
  // back to back instead of one round trip at a time (1.5 us -> of the split-K tail, profiles/experiments/tc_trace.py)
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(raddr));
  return v;
}

// ---- TMA -------------------------------------------------------------------------------------------------
struct TcMaps {
  CUtensorMap act[2];  // gathered operand (im2col mode; WGRAD: the activation, B operand)
  CUtensorMap wgt[2];  // plain operand (tiled mode; WGRAD: dout, A operand)
};

// Strided data gradient as per-parity-class stride-1 gathers (specification + CPU check: oracle/strided_dgrad.py,
// tests/test_strided_dgrad_spec.py).  The output pixels split into the stride x stride classes (ey, ex) = ((y + pad) % stride,
// (x + pad) % stride); class pixels are y = stride * iy + y0, and only the taps r = ey + stride * tr reach them:
//     din[y, x] = sum_{tr, ts} dout[iy + cy - tr, ix + cx - ts] . w[ey + stride tr, ex + stride ts]
// i.e. an im2col load over dout with traversal stride 1, lower corner L = c - (T - 1) and filter offset (T - 1) - t.  One launch
// covers all classes (blockIdx.x enumerates (class, m-tile) pairs); a class no tap reaches writes zeros.  Replaces the cp.async
// producer with its 4x zero-filled taps (round 1: 10 % of a config-2 and 13 % of a config-3 iteration).
constexpr int TC_MAXCLS = 4;   // stride 2
struct ClsPlan {
  int ncls, stride;
  int tile0[TC_MAXCLS + 1];                                  // first m-tile of each class, tile0[ncls] = all tiles
  int Hc[TC_MAXCLS], Wc[TC_MAXCLS], y0[TC_MAXCLS], x0[TC_MAXCLS];   // class pixel grid and its first pixel
  int Ty[TC_MAXCLS], Tx[TC_MAXCLS], Ly[TC_MAXCLS], Lx[TC_MAXCLS], ey[TC_MAXCLS], ex[TC_MAXCLS];
};
struct TcMapsCls {
  CUtensorMap act[2 * TC_MAXCLS];  // [class][source]: im2col maps over dout with the class's corners
  CUtensorMap wgt[2];
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// im2col load of `pixelsPerColumn` pixels x `channelsPerPixel` channels starting at base pixel (w, h, n), filter offset (ow, oh)
__device__ __forceinline__ void tma_im2col(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c, int w, int h, int n, int ow, int oh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow), "h"((uint16_t)oh)
      : "memory");
}
__device__ __forceinline__ void tma_tile2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_tile3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Fused consumer of an FPROP result (GemmEpilogue): NV consecutive channels n..n+NV-1 of GEMM row `row` (element offset
// row = m * Nc).  Mirrors layers.cu bnact_fwd_kernel (kind 1) / bnact_tan_fwd_kernel (kind 2) expression by expression.
template <int NV>
__device__ __forceinline__ void fused_bnact(const GemmEpilogue& e, long long row, int n, float (&v)[NV]) {
  const long long o = row + n;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float u = v[j];
    if (e.kind == 1) {
      if (e.has_bn) u = fmaf(u, __ldg(e.scale + n + j), __ldg(e.shift + n + j));
      if (e.res != nullptr) u += e.res[o + j];
      u = e.relu ? fmaxf(u, 0.f) : u;
      v[j] = e.round_out ? tf32_rna(u) : u;
    } else {
      if (e.has_bn) {
        const float xhat = fmaf(e.pre[o + j], __ldg(e.inv + n + j), __ldg(e.nrm + n + j));
        u = fmaf(__ldg(e.scale + n + j), u, fmaf(__ldg(e.v_gamma + n + j), xhat, __ldg(e.v_beta + n + j)));
      }
      if (e.res != nullptr) u += e.res[o + j];
      if (e.relu && !(e.post[o + j] > 0.f)) u = 0.f;
      v[j] = e.round_out ? tf32_rna(u) : u;
    }
  }
}

// Phase timestamps of CTA (0,0,0) (SM clock), compiled in with -DBRE_TC_TRACE for profiles/experiments/tc_trace.py only.
#ifdef BRE_TC_TRACE
__device__ long long g_tc_trace[16];
#define TC_MARK(i, cond) do { if ((cond) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_trace[i] = clock64(); } while (0)
#else
#define TC_MARK(i, cond) do { } while (0)
#endif

template <int MODE, int BN, bool TMA, bool CLS = false>
__global__ void __launch_bounds__(TC_BLOCK) igemm_tc_kernel(GemmArgs a, TcDims d, int proxy_fence,
                                                            const __grid_constant__ std::conditional_t<CLS, TcMapsCls, TcMaps> maps, ClsPlan plan) {
  static_assert(!CLS || (MODE == GEMM_DGRAD && TMA), "per-class gathers: strided dgrad with the TMA producer only");
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                  // [STAGES][A_BYTES]
  const int nst = 1 << d.stage_shift, stage_mask = nst - 1;
  uint8_t* sB = smem + nst * A_BYTES;                  // [stages][B_BYTES]
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];   // loaders -> MMA issuer (cp.async completion, 128 arrivals)
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];  // MMA issuer -> loaders (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t s_tmem;

  TC_MARK(0, threadIdx.x == 0);
  pdl_launch_dependents();   // the successor may start its own prologue; it blocks in its griddepcontrol.wait
  const ConvGeom g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.y * BN, z = blockIdx.z;
  int m0 = blockIdx.x * TC_BM;
  int kb_begin = z * d.kblocks_per_split;
  int kb_end = min(d.total_kblocks, kb_begin + d.kblocks_per_split);
  const int HoWo = g.Ho * g.Wo, HW = g.H * g.W;
  // per-class view (CLS): which class this CTA serves, its pixel grid / taps, and its own k-block range
  int cls = 0, cHc = 0, cWc = 0, cTy = 0, cTx = 0, cls_kps = 0, Mrows = d.M;
  if (CLS) {
#pragma unroll
    for (int c = 1; c < TC_MAXCLS; ++c)
      if (c < plan.ncls && (int)blockIdx.x >= plan.tile0[c]) cls = c;
    m0 = ((int)blockIdx.x - plan.tile0[cls]) * TC_BM;
    cHc = plan.Hc[cls]; cWc = plan.Wc[cls]; cTy = plan.Ty[cls]; cTx = plan.Tx[cls];
    Mrows = g.N * cHc * cWc;
    cls_kps = cTy * cTx * (g.Co / TC_BK);
    const int total = cls_kps * a.nsrc, per = (total + (int)gridDim.z - 1) / (int)gridDim.z;
    kb_begin = z * per;
    kb_end = min(total, kb_begin + per);
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(&bar_full[s], TMA ? 1 : TC_THREADS); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem;
  TC_MARK(1, threadIdx.x == 0);

  // instruction descriptor (UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4
  constexpr uint32_t a_mn = (MODE == GEMM_WGRAD) ? 1u : 0u;
  constexpr uint32_t b_mn = (MODE == GEMM_FPROP) ? 0u : 1u;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);

  // ---- per-thread fixed decode of the gathered (activation) operand ----------------------------------------
  // Loader thread t handles 16-byte granule column (t % 8) of A rows (t / 8) + 16 j, j = 0..7: the eight lanes of a
  // quarter warp read one full 128-byte line of a pixel row (coalesced) and write one swizzled 128-byte smem row
  // (conflict-free).  Everything that depends only on the row is computed once: the element offset of the row's
  // anchor pixel and a bit mask over the R*S filter taps saying which taps land inside the tensor, so that staging a
  // k-block costs one shift/and + one add + one cp.async per row (the loader warps are issue-latency bound otherwise).
  constexpr int A_VEC = TC_BM * TC_BK / 4 / TC_THREADS;  // 16-byte granules per thread per stage: 8
  constexpr int B_VEC = BN * TC_BK / 4 / TC_THREADS;     // 4 (BN = 64)
  const int gcol = tid & 7, grow = (tid >> 3) & 15;
  int a_off[A_VEC];                  // element offset of the anchor pixel (+ granule column)
  unsigned long long a_taps[A_VEC];  // bit rs: tap (r, s) of this row reads inside the tensor
  int a_yx[A_VEC];                   // slow path (strided dgrad): packed (y << 16) | x and image index in a_off
  if (!TMA && (MODE == GEMM_FPROP || MODE == GEMM_DGRAD)) {
#pragma unroll
    for (int j = 0; j < A_VEC; ++j) {
      const int m = m0 + grow + 16 * j;
      a_off[j] = 0; a_taps[j] = 0ull; a_yx[j] = 0;
      if (m < d.M) {
        if (MODE == GEMM_FPROP) {
          const int img = m / HoWo, rem = m - img * HoWo;
          const int p = rem / g.Wo, q = rem - p * g.Wo;
          const int y0 = p * g.stride - g.pad, x0 = q * g.stride - g.pad;
          a_off[j] = (int)(img * a.x_sN) + (y0 * g.W + x0) * a.x_sP + gcol * 4;
          for (int r = 0; r < g.R; ++r)
            for (int s2 = 0; s2 < g.S; ++s2)
              if (y0 + r >= 0 && y0 + r < g.H && x0 + s2 >= 0 && x0 + s2 < g.W) a_taps[j] |= 1ull << (r * g.S + s2);
        } else {
          const int img = m / HW, rem = m - img * HW;
          const int y = rem / g.W, x = rem - y * g.W;
          // p = (y + pad - r) / stride must be exact.  With yb = y + pad = stride * py + ey the valid taps are r = ey + stride * t
          // and p = py - t: anchor the row at (py, qx) and keep (ey, ex) so that a tap costs two shifts/divides and an fma.
          const int yb = y + g.pad, xb = x + g.pad;
          const int py = yb / g.stride, qx = xb / g.stride;
          const int ey = yb - py * g.stride, ex = xb - qx * g.stride;
          a_off[j] = ((img * g.Ho + py) * g.Wo + qx) * g.Co + gcol * 4;
          a_yx[j] = (ey << 16) | ex;
          for (int r = ey, pp = py; r < g.R; r += g.stride, --pp)
            for (int s2 = ex, qq = qx; s2 < g.S; s2 += g.stride, --qq)
              if (pp >= 0 && pp < g.Ho && qq >= 0 && qq < g.Wo) a_taps[j] |= 1ull << (r * g.S + s2);
        }
      }
    }
  }
  // running decode of the k-block sequence this CTA walks: (source, filter tap, first channel) advance incrementally
  int it_src, it_rs, it_c0;
  {
    const int kch = (MODE == GEMM_DGRAD) ? g.Co : g.Ci;
    it_src = kb_begin / d.kblocks_per_src;
    const int kbase = (kb_begin - it_src * d.kblocks_per_src) * TC_BK;
    it_rs = (MODE == GEMM_WGRAD) ? 0 : kbase / kch;
    it_c0 = (MODE == GEMM_WGRAD) ? kbase : kbase - it_rs * kch;
  }

  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;  // a trailing split may be empty: it contributes zeros
  // TMA producer (one thread): the CTA's first GEMM row fixes the base pixel of every im2col load; per k-block only the
  // filter-tap offsets and the channel coordinate change.
  int base_n = 0, base_h = 0, base_w = 0;
  if (CLS) {
    const int per = cHc * cWc;
    base_n = m0 / per;
    const int rem = m0 - base_n * per;
    const int iy0 = rem / cWc, ix0 = rem - iy0 * cWc;
    base_h = iy0 + plan.Ly[cls]; base_w = ix0 + plan.Lx[cls];
  } else if (TMA && MODE != GEMM_WGRAD) {
    const int per = (MODE == GEMM_FPROP) ? HoWo : HW, wid = (MODE == GEMM_FPROP) ? g.Wo : g.W;
    base_n = m0 / per;
    const int rem = m0 - base_n * per;
    const int p0 = rem / wid, q0 = rem - p0 * wid;
    if (MODE == GEMM_FPROP) { base_h = p0 * g.stride - g.pad; base_w = q0 * g.stride - g.pad; }
    else { base_h = p0 + g.pad - (g.R - 1); base_w = q0 + g.pad - (g.S - 1); }   // stride-1 dgrad: flipped taps
  }
  // Running decode of the producer's k-block sequence.  The producer lane is a single thread on the critical path of the
  // whole CTA: no divisions inside the loop (an integer division costs it ~100 cycles), everything advances incrementally.
  struct KbState { int src, rs, r, s, c0, img, p, q; };
  auto kb_init = [&](int kb) {
    KbState st;
    const int kch = (MODE == GEMM_DGRAD) ? g.Co : g.Ci;
    if (CLS) {   // k = (source, class tap (tr, ts), channel); st.r / st.s hold the class-tap indices, st.rs the weight's filter cell
      st.src = cls_kps > 0 ? kb / cls_kps : 0;
      const int rem = kb - st.src * cls_kps, cpb = g.Co / TC_BK;
      const int tap = rem / cpb;
      st.c0 = (rem - tap * cpb) * TC_BK;
      st.r = cTx > 0 ? tap / cTx : 0; st.s = tap - st.r * cTx;
      st.rs = (plan.ey[cls] + plan.stride * st.r) * g.S + plan.ex[cls] + plan.stride * st.s;
      st.img = st.p = st.q = 0;
      return st;
    }
    st.src = kb / d.kblocks_per_src;
    const int kbase = (kb - st.src * d.kblocks_per_src) * TC_BK;
    if (MODE == GEMM_WGRAD) {
      st.rs = st.r = st.s = 0; st.c0 = kbase;
      st.img = kbase / HoWo;
      const int rem = kbase - st.img * HoWo;
      st.p = rem / g.Wo; st.q = rem - st.p * g.Wo;
    } else {
      st.rs = kbase / kch; st.c0 = kbase - st.rs * kch;
      st.r = st.rs / g.S; st.s = st.rs - st.r * g.S;
      st.img = st.p = st.q = 0;
    }
    return st;
  };
  auto kb_advance = [&](KbState& st) {
    const int kch = (MODE == GEMM_DGRAD) ? g.Co : g.Ci;
    st.c0 += TC_BK;
    if (CLS) {
      if (st.c0 >= kch) {
        st.c0 = 0;
        if (++st.s == cTx) { st.s = 0; if (++st.r == cTy) { st.r = 0; ++st.src; } }
        st.rs = (plan.ey[cls] + plan.stride * st.r) * g.S + plan.ex[cls] + plan.stride * st.s;
      }
      return;
    }
    if (MODE == GEMM_WGRAD) {
      if (st.c0 >= d.kblocks_per_src * TC_BK) { st.c0 = 0; ++st.src; st.img = st.p = st.q = 0; }
      else {
        st.q += TC_BK;
        while (st.q >= g.Wo) { st.q -= g.Wo; if (++st.p == g.Ho) { st.p = 0; ++st.img; } }
      }
    } else if (st.c0 >= kch) {
      st.c0 = 0; ++st.rs;
      if (++st.s == g.S) { st.s = 0; if (++st.r == g.R) { st.r = 0; st.rs = 0; ++st.src; } }
    }
  };
  // WGRAD: the filter tap / channel of the CTA's B columns are fixed
  int wg_c[BN / 32], wg_r[BN / 32], wg_s[BN / 32];
  if (TMA && MODE == GEMM_WGRAD) {
#pragma unroll
    for (int h = 0; h < BN / 32; ++h) {
      const int n = n0 + 32 * h;
      const int rs = n / g.Ci;
      wg_c[h] = n - rs * g.Ci; wg_r[h] = rs / g.S; wg_s[h] = rs - wg_r[h] * g.S;
    }
  }
  // One k-block = the weight-side loads and the activation-side loads on the same stage barrier.  They are separate calls because
  // the weight side of the first ring stages is issued *before* griddepcontrol.wait (below): weights do not depend on the
  // predecessor kernel, so their HBM / L2 round trip overlaps the predecessor's tail.
  auto issue_wgt_tma = [&](int stage, const KbState& st) {
    uint64_t* bar = &bar_full[stage];
    const uint32_t pb = smem_u32(sB) + stage * B_BYTES;
    if (MODE == GEMM_FPROP) {
      tma_tile2d(pb, &maps.wgt[st.src], bar, st.rs * g.Ci + st.c0, n0);
    } else if (MODE == GEMM_DGRAD) {
#pragma unroll
      for (int h = 0; h < BN / 32; ++h) tma_tile3d(pb + h * 4096, &maps.wgt[st.src], bar, n0 + 32 * h, st.rs, st.c0);
    } else {
      const uint32_t pa = smem_u32(sA) + stage * A_BYTES;
#pragma unroll
      for (int h = 0; h < TC_BM / 32; ++h) tma_tile2d(pa + h * 4096, &maps.wgt[st.src], bar, m0 + 32 * h, st.c0);   // k = pixel
    }
  };
  auto issue_act_tma = [&](int stage, const KbState& st) {
    uint64_t* bar = &bar_full[stage];
    const uint32_t pa = smem_u32(sA) + stage * A_BYTES, pb = smem_u32(sB) + stage * B_BYTES;
    if (MODE == GEMM_FPROP) {
      tma_im2col(pa, &maps.act[st.src], bar, st.c0, base_w, base_h, base_n, st.s, st.r);
    } else if (MODE == GEMM_DGRAD) {
      if (CLS) tma_im2col(pa, &maps.act[2 * cls + st.src], bar, st.c0, base_w, base_h, base_n, cTx - 1 - st.s, cTy - 1 - st.r);
      else tma_im2col(pa, &maps.act[st.src], bar, st.c0, base_w, base_h, base_n, g.S - 1 - st.s, g.R - 1 - st.r);
    } else {
#pragma unroll
      for (int h = 0; h < BN / 32; ++h)
        tma_im2col(pb + h * 4096, &maps.act[st.src], bar, wg_c[h], st.q * g.stride - g.pad, st.p * g.stride - g.pad, st.img, wg_s[h], wg_r[h]);
    }
  };

  if (TMA && tid == 0 && (proxy_fence & 2)) {
    for (int sidx = 0; sidx < a.nsrc; ++sidx) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.act[(CLS ? 2 * cls : 0) + sidx])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.wgt[sidx])) : "memory");
    }
  }
  // `nprod` producer lanes (lane 0 of loader warps 0..nprod-1, default 2), k-blocks dealt round-robin: a cp.async.bulk.tensor
  // costs its issuing thread ~140 cycles (profiles/experiments/tc_trace.py); two issuers keep up with the six loads per
  // k-block of the wgrad form, more make no difference
  const int nprod = (proxy_fence >> 2) & 7;
  const bool producer = TMA && (tid & 31) == 0 && warp < nprod && warp < nkb;
  KbState pst;
  memset(&pst, 0, sizeof(pst));
  uint32_t pre_wgt = 0;   // ring stages whose weight loads are already in flight when the wait below returns
  if (producer) {
    pst = kb_init(kb_begin + warp);
    if (MODE != GEMM_WGRAD && a.wgt_static != 0) {
      KbState st = pst;
      for (int i = warp; i < nkb && i < nst; i += nprod) {
        if ((a.wgt_static >> st.src) & 1) {
          mbar_expect_tx(&bar_full[i], A_BYTES + B_BYTES);
          issue_wgt_tma(i, st);
          pre_wgt |= 1u << i;
        }
        for (int u = 0; u < nprod; ++u) kb_advance(st);
      }
    }
  }

  // Everything above touched only kernel parameters, shared memory, TMEM and weights the caller declared constant across the
  // predecessor (GemmArgs::wgt_static); from here on global memory written by the predecessor kernel is read.
  pdl_wait();
  TC_MARK(2, threadIdx.x == 0);

  // Stage one k-block: cp.async (LDGSTS, 16 B, zero-fill for padding / out-of-range taps) straight from global memory
  // into the UMMA operand layouts -- no register staging, so up to TC_STAGES k-blocks of loads stay in flight.
  auto issue_block = [&](int stage) {
    const int src = it_src;
    const int kch = (MODE == GEMM_DGRAD) ? g.Co : g.Ci;
    const int kbase = (MODE == GEMM_WGRAD) ? it_c0 : it_rs * kch + it_c0;
    const float* __restrict__ act = a.act[src];
    const float* __restrict__ wgt = a.wgt[src];
    const uint32_t pa = smem_u32(sA + stage * A_BYTES), pb = smem_u32(sB + stage * B_BYTES);
    if (MODE == GEMM_FPROP) {
      // A(m, k) = in[img, y0 + r, x0 + s, c], k = (r, s, c); the 32-wide k-block lies inside one (r, s) cell (Ci % 32 == 0)
      const int r = it_rs / g.S, s = it_rs - r * g.S;
      const int tapoff = (r * g.W + s) * a.x_sP + it_c0;
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) {
        const bool ok = (a_taps[j] >> it_rs) & 1ull;
        cp_async16(pa + off_k128(grow + 16 * j, gcol), ok ? act + (a_off[j] + tapoff) : act, ok ? 16u : 0u);
      }
      // B(n, k) = W[n][k] (row-major [Co][K])
      const float* q = wgt + (long long)(n0 + grow) * d.K + kbase + gcol * 4;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) cp_async16(pb + off_k128(grow + 16 * j, gcol), q + (long long)(16 * j) * d.K, 16u);
    } else if (MODE == GEMM_DGRAD) {
      // A(m, k) = dout[img, (y + pad - r)/stride, (x + pad - s)/stride, ko], k = (r, s, ko)   (Co % 32 == 0)
      const int rs = it_rs, k0 = it_c0;
      const int r = rs / g.S, s = rs - r * g.S;
      if (g.stride == 1) {
        const int tapoff = k0 - (r * g.Wo + s) * g.Co;
#pragma unroll
        for (int j = 0; j < A_VEC; ++j) {
          const bool ok = (a_taps[j] >> rs) & 1ull;
          cp_async16(pa + off_k128(grow + 16 * j, gcol), ok ? act + (a_off[j] + tapoff) : act, ok ? 16u : 0u);
        }
      } else {
        const bool st2 = g.stride == 2;
#pragma unroll
        for (int j = 0; j < A_VEC; ++j) {
          const bool ok = (a_taps[j] >> rs) & 1ull;
          const int dr = r - (a_yx[j] >> 16), ds = s - (a_yx[j] & 0xffff);
          const int tr = st2 ? dr >> 1 : dr / g.stride, ts = st2 ? ds >> 1 : ds / g.stride;
          cp_async16(pa + off_k128(grow + 16 * j, gcol), ok ? act + (a_off[j] + k0 - (tr * g.Wo + ts) * g.Co) : act, ok ? 16u : 0u);
        }
      }
      // B(n = ci, k) = W[ko][r][s][ci]: contiguous along n -> MN-major.  lanes along n (16 granules = 64 ci), 8 k per pass
      const int n4 = tid & 15, kk = tid >> 4;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) {
        const int k = kk + 8 * j;
        const float* gb = wgt + ((long long)(k0 + k) * (g.R * g.S) + rs) * g.Ci + n0 + n4 * 4;
        cp_async16(pb + off_mnmajor<BN>(4 * n4, k), gb, 16u);
      }
    } else {
      // WGRAD: k = pixel.  A(m = ko, k) = dout[pixel][ko] (MN-major): lanes along m (32 granules = 128 ko), 4 k per pass
      {
        const int m4 = tid & 31, kk = tid >> 5;
#pragma unroll
        for (int j = 0; j < A_VEC; ++j) {
          const int k = kk + 4 * j;
          const int pix = kbase + k;
          const bool kok = pix < d.K;
          const bool mok = kok && (m0 + m4 * 4 < d.M);
          const float* ga = mok ? wgt + (long long)pix * g.Co + m0 + m4 * 4 : wgt;
          cp_async16(pa + off_mnmajor<TC_BM>(4 * m4, k), ga, mok ? 16u : 0u);
        }
      }
      // B(n = (r, s, c), k) = in[img, p*stride - pad + r, q*stride - pad + s, c] (MN-major, Ci % 4 == 0)
      {
        const int n4 = tid & 15, kk = tid >> 4;
        const int n = n0 + n4 * 4;
        const int rs = n / g.Ci, c = n - rs * g.Ci;
        const int r = rs / g.S, s = rs - r * g.S;
#pragma unroll
        for (int j = 0; j < B_VEC; ++j) {
          const int k = kk + 8 * j;
          const int pix = kbase + k;
          bool ok = pix < d.K;
          const float* gb = act;
          if (ok) {
            const int img = pix / HoWo, rem = pix - img * HoWo;
            const int p = rem / g.Wo, q = rem - p * g.Wo;
            const int h = p * g.stride - g.pad + r, w = q * g.stride - g.pad + s;
            ok = h >= 0 && h < g.H && w >= 0 && w < g.W;
            if (ok) gb = act + img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + c;
          }
          cp_async16(pb + off_mnmajor<BN>(4 * n4, k), gb, ok ? 16u : 0u);
        }
      }
    }
    // advance the running (source, tap, channel) decode to the next k-block
    it_c0 += TC_BK;
    if (MODE == GEMM_WGRAD) {
      if (it_c0 >= d.kblocks_per_src * TC_BK) { it_c0 = 0; ++it_src; }
    } else if (it_c0 >= kch) {
      it_c0 = 0;
      if (++it_rs == g.R * g.S) { it_rs = 0; ++it_src; }
    }
  };

  // ---- main loop, warp-specialised: warps 0-3 stream k-blocks into a TC_STAGES-deep ring with cp.async and signal
  //      "full" through cp.async.mbarrier.arrive; one thread of warp 4 waits for "full", issues the four tcgen05.mma of
  //      the k-block and lets tcgen05.commit signal "empty" when the tensor core has consumed the stage.  No CTA-wide
  //      barrier inside the loop.
  if (warp == TC_THREADS / 32) {
    if (tid == TC_THREADS) {
      // descriptors of stage 0 / MMA slice 0; the 14-bit address field counts 16-byte units, so stage and slice offsets are
      // plain additions on the low word (shared memory is < 256 KB: no carry out of the field)
      const uint64_t adesc0 = a_mn ? make_desc(smem_u32(sA), 4096, 512, 1) : make_desc(smem_u32(sA), 16, 1024, 2);
      const uint64_t bdesc0 = b_mn ? make_desc(smem_u32(sB), 4096, 512, 1) : make_desc(smem_u32(sB), 16, 1024, 2);
      constexpr uint32_t a_step = (a_mn ? 1024u : 32u) >> 4, b_step = (b_mn ? 1024u : 32u) >> 4;
      for (int i = 0; i < nkb; ++i) {
        const int stage = i & stage_mask;
        mbar_wait(&bar_full[stage], (uint32_t)((i >> d.stage_shift) & 1));
        TC_MARK(4, i == 0);
        TC_MARK(11, i == nkb - 1);
        // The mbarrier phase completes only after every cp.async of this stage has been performed, so the data is in
        // shared memory when the wait returns; like CUTLASS' sm100 cp.async mainloop no fence.proxy.async is issued
        // here (measured: it costs ~0.5 us per k-block on the single-thread critical path).  BRE_TC_PROXY_FENCE=1 re-enables it.
        if (proxy_fence & 1) fence_proxy_async();
        tc_fence_after();
        const uint64_t adesc_s = adesc0 + (uint64_t)((uint32_t)stage * (A_BYTES >> 4));
        const uint64_t bdesc_s = bdesc0 + (uint64_t)((uint32_t)stage * (B_BYTES >> 4));
#pragma unroll
        for (int j = 0; j < TC_BK / 8; ++j)
          umma_tf32(tmem_d, adesc_s + (uint64_t)(j * a_step), bdesc_s + (uint64_t)(j * b_step), idesc, (i > 0 || j > 0) ? 1u : 0u);
        umma_commit(&bar_empty[stage]);
      }
      umma_commit(&bar_done);
      TC_MARK(5, true);
    }
    __syncwarp();
  } else {
    if (TMA) {
      if (producer) {
        KbState st = pst;
        for (int i = warp; i < nkb; i += nprod) {
          const int stage = i & stage_mask;
          if (i >= nst) mbar_wait(&bar_empty[stage], (uint32_t)(((i >> d.stage_shift) - 1) & 1));
          if (i >= nst || !((pre_wgt >> i) & 1u)) {
            mbar_expect_tx(&bar_full[stage], A_BYTES + B_BYTES);
            issue_wgt_tma(stage, st);
          }
          issue_act_tma(stage, st);
          TC_MARK(3, i == (nkb < nst ? nkb : nst) - 1);
          for (int u = 0; u < nprod; ++u) kb_advance(st);
        }
      }
      __syncwarp();
    } else {
      for (int i = 0; i < nkb; ++i) {
        const int stage = i & stage_mask;
        if (i >= nst) mbar_wait(&bar_empty[stage], (uint32_t)(((i >> d.stage_shift) - 1) & 1));
        issue_block(stage);
        cp_async_arrive(&bar_full[stage]);
      }
    }
    mbar_wait(&bar_done, 0);
    TC_MARK(6, tid == 0);
    tc_fence_after();
  }

  // ---- epilogue: TMEM -> registers; thread `tid` holds row m0 + tid, BN columns in chunks of 32 -----------------
  const int splits = gridDim.z;
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
  float v[32];
  const int m = m0 + tid;
  const bool row_ok = m < Mrows;
  const bool is_loader = warp < TC_THREADS / 32;
  auto out_row = [&](int mm, long long& row, int& cs) {
    if (CLS) {   // class pixel (img, iy, ix) -> input pixel (y0 + stride * iy, x0 + stride * ix)
      const int per = cHc * cWc;
      const int img = mm / per, rem = mm - img * per;
      const int iy = rem / cWc, ix = rem - iy * cWc;
      row = img * a.x_sN + (long long)((plan.y0[cls] + plan.stride * iy) * g.W + plan.x0[cls] + plan.stride * ix) * a.x_sP;
      cs = a.x_sC;
    } else if (MODE == GEMM_DGRAD) {
      const int img = mm / HW;
      row = img * a.x_sN + (long long)(mm - img * HW) * a.x_sP;
      cs = a.x_sC;
    } else {
      row = (long long)mm * d.Nc;
      cs = 1;
    }
  };
  if (!is_loader) {
    // MMA-issuer warp: nothing to write back
  } else if (splits == 1) {
    long long row = 0;
    int cs = 1;
    if (row_ok) out_row(m, row, cs);
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      if (nkb > 0) tmem_ld32(lane_base + c, v);  // warp-collective: executed by all lanes, also for rows beyond M
      else {                                    // (a class no filter tap reaches: zeros)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (!row_ok) continue;
      const int n = n0 + c;
      if (MODE == GEMM_FPROP && a.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + n + j);
      }
      float* op = a.out + row + (long long)n * cs;
      if (MODE == GEMM_FPROP && a.epi.kind != 0) {
        if (a.out != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        fused_bnact<32>(a.epi, row, n, v);
        float* op2 = a.epi.out2 + row + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op2 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else if (cs == 1) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (a.accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(op + j);
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
          }
          *reinterpret_cast<float4*>(op + j) = t;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float* q = op + (long long)j * cs;
          *q = a.accumulate ? *q + v[j] : v[j];
        }
      }
    }
  } else {
    // split-K inside a thread-block cluster (cluster = the `splits` CTAs of one output tile along z): every CTA parks
    // its partial accumulator tile in its own shared memory (the operand ring is idle now), one cluster barrier, then
    // CTA `z` sums rows [z*128/S, (z+1)*128/S) of all S partials straight out of the peers' shared memory (DSMEM,
    // ld.shared::cluster) in fixed rank order -- deterministic, no global workspace, no atomics, no "last CTA" tail.
    float* part = reinterpret_cast<float*>(sA);  // [128 rows][BN] fp32, 16-byte chunks XOR-swizzled by (row & 15)
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      if (nkb > 0) tmem_ld32(lane_base + c, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int chunk = ((c + j) >> 2) ^ (tid & (BN / 4 - 1) & 15);
        *reinterpret_cast<float4*>(part + tid * BN + chunk * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  if (splits > 1) {
    TC_MARK(7, tid == 0);
    cluster_sync_all();   // every thread of every CTA in the cluster (partials visible cluster-wide)
    TC_MARK(8, tid == 0);
    if (is_loader) {
      // Every thread owns BN/(4 S) float4 slots of this CTA's row slice and needs the S peers' copies of each: BN/4 = 16
      // remote 16-byte loads per thread whatever S is.  All of them are issued before the first is consumed (a DSMEM round
      // trip is ~0.5 us; one slot at a time made this phase 2.2 us), then summed in fixed rank order.
      auto reduce_rows = [&](auto s_tag) {
        constexpr int S = decltype(s_tag)::value;
        constexpr int C4 = BN / 4, SLOTS = (TC_BM / S) * C4 / TC_THREADS > 0 ? (TC_BM / S) * C4 / TC_THREADS : 1;
        static_assert(BN >= 64 || S <= 8, "narrow tiles: at most 8 splits");
        const uint32_t part_s = smem_u32(sA);
        float4 t[SLOTS][S];
        int rr[SLOTS], cc[SLOTS];
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
          const int slot = tid + sl * TC_THREADS;
          rr[sl] = z * (TC_BM / S) + slot / C4;
          cc[sl] = slot % C4;
          const uint32_t local = part_s + (uint32_t)(rr[sl] * BN + ((cc[sl] ^ (rr[sl] & (BN / 4 - 1) & 15)) << 2)) * 4u;
#pragma unroll
          for (int q = 0; q < S; ++q) t[sl][q] = ld_dsmem4(local, (uint32_t)q);
        }
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
          const int r = rr[sl], c4 = cc[sl];
          if (m0 + r >= Mrows) continue;
          float4 acc4 = t[sl][0];
#pragma unroll
          for (int q = 1; q < S; ++q) { acc4.x += t[sl][q].x; acc4.y += t[sl][q].y; acc4.z += t[sl][q].z; acc4.w += t[sl][q].w; }
          const int n = n0 + c4 * 4;
          if (MODE == GEMM_FPROP && a.bias != nullptr) {
            acc4.x += __ldg(a.bias + n); acc4.y += __ldg(a.bias + n + 1); acc4.z += __ldg(a.bias + n + 2); acc4.w += __ldg(a.bias + n + 3);
          }
          long long row;
          int cs;
          out_row(m0 + r, row, cs);
          float* op = a.out + row + (long long)n * cs;
          if (MODE == GEMM_FPROP && a.epi.kind != 0) {
            if (a.out != nullptr) *reinterpret_cast<float4*>(op) = acc4;
            float vv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
            fused_bnact<4>(a.epi, row, n, vv);
            *reinterpret_cast<float4*>(a.epi.out2 + row + n) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          } else if (cs == 1) {
            if (a.accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(op);
              acc4.x += o.x; acc4.y += o.y; acc4.z += o.z; acc4.w += o.w;
            }
            *reinterpret_cast<float4*>(op) = acc4;
          } else {
            const float vv[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float* qq = op + (long long)j * cs;
              *qq = a.accumulate ? *qq + vv[j] : vv[j];
            }
          }
        }
      };
      switch (splits) {
        case 2: reduce_rows(std::integral_constant<int, 2>{}); break;
        case 4: reduce_rows(std::integral_constant<int, 4>{}); break;
        case 8: reduce_rows(std::integral_constant<int, 8>{}); break;
        default:
          if constexpr (BN >= 64) reduce_rows(std::integral_constant<int, 16>{});
          break;
      }
    }
    TC_MARK(9, tid == 0);
    cluster_sync_all();   // nobody leaves (and frees its shared memory) while a peer may still be reading it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
  TC_MARK(10, tid == 0);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- tensor maps (host) ----------------------------------------------------------------------------------
// cuTensorMapEncode* are driver entry points; they are fetched through the runtime so that the library does not depend on
// the link order of libcuda.  Maps are cached by (pointer, geometry): the engine's buffers are static, so every map is
// encoded once per engine lifetime and reused by every launch / graph replay.
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaApi {
  EncodeIm2colFn im2col = nullptr;
  EncodeTiledFn tiled = nullptr;
  int driver = 0;
  bool ok = false;
};
const TmaApi& tma_api() {
  static const TmaApi api = [] {
    TmaApi t;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      t.im2col = reinterpret_cast<EncodeIm2colFn>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      t.tiled = reinterpret_cast<EncodeTiledFn>(f);
    cudaDriverGetVersion(&t.driver);
    t.ok = t.im2col != nullptr && t.tiled != nullptr;
    return t;
  }();
  return api;
}

using MapKey = std::array<long long, 16>;
std::mutex g_map_mutex;
std::map<MapKey, CUtensorMap> g_map_cache;

// im2col map over an NHWC view [N][H][W][C] (element strides sN, sP = pixel stride, channels contiguous).  One load =
// `pixels` consecutive base pixels (walked along W, then H, then N inside the box [lower, dim - 1 + upper], step `stride`)
// x `chans` channels; elements outside the tensor read as zero.
bool im2col_map(CUtensorMap* out, const float* base, int N, int H, int W, int C, long long sN, int sP, int lower_w, int lower_h,
                int upper_w, int upper_h, int stride, int chans, int pixels, CUtensorMapSwizzle swz) {
  const MapKey key = {0, (long long)reinterpret_cast<uintptr_t>(base), N, H, W, C, sN, sP, lower_w, lower_h, upper_w, upper_h, stride, chans,
                      pixels, (long long)swz};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *out = it->second; return true; }
  const TmaApi& api = tma_api();
  if (!api.ok) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)sP * 4, (cuuint64_t)W * sP * 4, (cuuint64_t)sN * 4};
  const int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w, upper_h};
  const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUtensorMap tm;
  const CUresult res = api.im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, lower, upper,
                                  (cuuint32_t)chans, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) return false;
  g_map_cache.emplace(key, tm);
  *out = tm;
  return true;
}

bool tiled_map(CUtensorMap* out, const float* base, int rank, const long long* dims_ll, const long long* strides_elems, const int* box_i,
               CUtensorMapSwizzle swz) {
  MapKey key = {1, (long long)reinterpret_cast<uintptr_t>(base), rank, (long long)swz};
  for (int i = 0; i < rank; ++i) { key[4 + i] = dims_ll[i]; key[8 + i] = box_i[i]; if (i + 1 < rank) key[12 + i] = strides_elems[i]; }
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *out = it->second; return true; }
  const TmaApi& api = tma_api();
  if (!api.ok) return false;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { dims[i] = (cuuint64_t)dims_ll[i]; box[i] = (cuuint32_t)box_i[i]; }
  for (int i = 0; i + 1 < rank; ++i) strides[i] = (cuuint64_t)strides_elems[i] * 4;
  CUtensorMap tm;
  const CUresult res = api.tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) return false;
  g_map_cache.emplace(key, tm);
  *out = tm;
  return true;
}

// Does the TMA producer cover this contraction?  (strided dgrad does not: see the header of this file.)
bool tma_eligible(const GemmArgs& a) {
  static const int tma_env = [] { const char* e = getenv("BRE_TC_TMA"); return e ? atoi(e) : 1; }();
  if (!tma_env || !tma_api().ok) return false;
  const ConvGeom& g = a.g;
  if (g.stride > 8 || g.pad > 127 || g.R - 1 - g.pad > 127 || g.S - 1 - g.pad > 127 || g.R > 128 || g.S > 128) return false;
  if (a.mode == GEMM_DGRAD) return g.stride == 1;
  if (a.mode == GEMM_WGRAD) return g.Ci % 32 == 0;
  return true;
}

bool build_maps(const GemmArgs& a, int BN, TcMaps* maps) {
  const ConvGeom& g = a.g;
  const CUtensorMapSwizzle K128 = CU_TENSOR_MAP_SWIZZLE_128B, MN32 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  for (int s = 0; s < a.nsrc; ++s) {
    if (a.mode == GEMM_FPROP) {
      if (!im2col_map(&maps->act[s], a.act[s], g.N, g.H, g.W, g.Ci, a.x_sN, a.x_sP, -g.pad, -g.pad, g.pad - (g.S - 1), g.pad - (g.R - 1),
                      g.stride, TC_BK, TC_BM, K128))
        return false;
      const long long K = (long long)g.R * g.S * g.Ci;
      const long long dims[2] = {K, g.Co}, strides[1] = {K};
      const int box[2] = {TC_BK, BN};
      if (!tiled_map(&maps->wgt[s], a.wgt[s], 2, dims, strides, box, K128)) return false;
    } else if (a.mode == GEMM_DGRAD) {
      // rows = pixels of the input gradient; the gathered tensor is dout [N][Ho][Wo][Co]; tap (r, s) reads (y + pad - r, x + pad - s)
      const int lw = g.pad - (g.S - 1), lh = g.pad - (g.R - 1);
      if (!im2col_map(&maps->act[s], a.act[s], g.N, g.Ho, g.Wo, g.Co, (long long)g.Ho * g.Wo * g.Co, g.Co, lw, lh, lw + g.W - g.Wo,
                      lh + g.H - g.Ho, 1, TC_BK, TC_BM, K128))
        return false;
      const long long dims[3] = {g.Ci, (long long)g.R * g.S, g.Co}, strides[2] = {g.Ci, (long long)g.R * g.S * g.Ci};
      const int box[3] = {32, 1, TC_BK};
      if (!tiled_map(&maps->wgt[s], a.wgt[s], 3, dims, strides, box, MN32)) return false;
    } else {
      // A(m = ko, k = pixel) = dout[pixel][ko]; B(n = (r, s, c), k = pixel) = im2col of the activation, 32 pixels x 32 channels
      const long long npix = (long long)g.N * g.Ho * g.Wo;
      const long long dims[2] = {g.Co, npix}, strides[1] = {g.Co};
      const int box[2] = {32, TC_BK};
      if (!tiled_map(&maps->wgt[s], a.wgt[s], 2, dims, strides, box, MN32)) return false;
      if (!im2col_map(&maps->act[s], a.act[s], g.N, g.H, g.W, g.Ci, a.x_sN, a.x_sP, -g.pad, -g.pad, g.pad - (g.S - 1), g.pad - (g.R - 1),
                      g.stride, 32, TC_BK, MN32))
        return false;
    }
  }
  return true;
}

// ---- per-class plan of a strided dgrad (oracle/strided_dgrad.py class_plan, per axis) ---------------------------------------------
struct AxisPlan { bool any; int y0, Hc, T, L, U; };
inline AxisPlan axis_plan(int H, int Ho, int R, int stride, int pad, int e) {
  AxisPlan p{false, 0, 0, 0, 0, 0};
  p.y0 = ((e - pad) % stride + stride) % stride;
  if (p.y0 >= H) return p;                       // no input pixel of this parity
  p.any = true;
  p.Hc = (H - p.y0 + stride - 1) / stride;
  if (e >= R) return p;                          // pixels exist, no tap reaches them: T = 0
  p.T = (R - e + stride - 1) / stride;
  const int c = (p.y0 + pad - e) / stride;
  p.L = c - (p.T - 1);
  p.U = p.Hc - Ho + p.L;
  return p;
}

bool cls_eligible(const GemmArgs& a) {
  static const int env = [] { const char* e = getenv("BRE_TC_STRIDED_TMA"); return e ? atoi(e) : 1; }();
  const ConvGeom& g = a.g;
  if (!env || a.mode != GEMM_DGRAD || g.stride != 2 || !tma_api().ok) return false;
  static const int tma_env = [] { const char* e = getenv("BRE_TC_TMA"); return e ? atoi(e) : 1; }();
  if (!tma_env || g.R > 16 || g.S > 16) return false;
  for (int ey = 0; ey < 2; ++ey)
    for (int horiz = 0; horiz < 2; ++horiz) {
      const AxisPlan p = horiz ? axis_plan(g.W, g.Wo, g.S, 2, g.pad, ey) : axis_plan(g.H, g.Ho, g.R, 2, g.pad, ey);
      if (p.any && p.T > 0 && (p.L < -128 || p.L > 127 || p.U < -128 || p.U > 127)) return false;
    }
  return true;
}

bool build_cls(const GemmArgs& a, ClsPlan* plan, TcMapsCls* maps) {
  const ConvGeom& g = a.g;
  memset(plan, 0, sizeof(*plan));
  plan->stride = g.stride;
  int tiles = 0, n = 0;
  for (int ey = 0; ey < g.stride; ++ey) {
    const AxisPlan py = axis_plan(g.H, g.Ho, g.R, g.stride, g.pad, ey);
    if (!py.any) continue;
    for (int ex = 0; ex < g.stride; ++ex) {
      const AxisPlan px = axis_plan(g.W, g.Wo, g.S, g.stride, g.pad, ex);
      if (!px.any) continue;
      if (n >= TC_MAXCLS) return false;
      const bool taps = py.T > 0 && px.T > 0;
      plan->Hc[n] = py.Hc; plan->Wc[n] = px.Hc; plan->y0[n] = py.y0; plan->x0[n] = px.y0;
      plan->Ty[n] = taps ? py.T : 0; plan->Tx[n] = taps ? px.T : 0;
      plan->Ly[n] = py.L; plan->Lx[n] = px.L; plan->ey[n] = ey; plan->ex[n] = ex;
      plan->tile0[n] = tiles;
      tiles += ceil_div((long long)g.N * py.Hc * px.Hc, TC_BM);
      for (int s = 0; s < a.nsrc && taps; ++s)
        if (!im2col_map(&maps->act[2 * n + s], a.act[s], g.N, g.Ho, g.Wo, g.Co, (long long)g.Ho * g.Wo * g.Co, g.Co, px.L, py.L, px.U, py.U, 1,
                        TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B))
          return false;
      if (!taps)   // never dereferenced (the class has no k-blocks), but a valid descriptor keeps prefetch.tensormap well defined
        for (int s = 0; s < a.nsrc; ++s) maps->act[2 * n + s] = maps->act[s];
      ++n;
    }
  }
  plan->ncls = n;
  plan->tile0[n] = tiles;
  for (int s = 0; s < a.nsrc; ++s) {
    const long long dims[3] = {g.Ci, (long long)g.R * g.S, g.Co}, strides[2] = {g.Ci, (long long)g.R * g.S * g.Ci};
    const int box[3] = {32, 1, TC_BK};
    if (!tiled_map(&maps->wgt[s], a.wgt[s], 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return false;
  }
  return n > 0;
}

template <int MODE, int BN, bool TMA, bool CLS = false>
int launch_tc(const GemmArgs& a, const TcDims& d0, const std::conditional_t<CLS, TcMapsCls, TcMaps>& maps, cudaStream_t stream,
              const ClsPlan* plan_in = nullptr) {
  TcDims d = d0;
  ClsPlan plan;
  memset(&plan, 0, sizeof(plan));
  if (CLS) plan = *plan_in;
  // CLS: the m-tiles of all classes side by side; the k extent that sizes the split is the largest class's
  const int tm = CLS ? plan.tile0[plan.ncls] : ceil_div(d.M, TC_BM), tn = d.Nc / BN;
  if (CLS) {
    int kmax = 0;
    for (int c = 0; c < plan.ncls; ++c) kmax = kmax > plan.Ty[c] * plan.Tx[c] ? kmax : plan.Ty[c] * plan.Tx[c];
    d.total_kblocks = kmax * (a.g.Co / TC_BK) * a.nsrc;
    if (d.total_kblocks < 1) d.total_kblocks = 1;
  }
  const long long tiles = (long long)tm * tn;
  // split-K factor = cluster size along z: a power of two <= 8 (portable cluster limit) that brings the grid to ~100 CTAs
  static const int max_splits_env = [] { const char* e = getenv("BRE_TC_MAX_SPLITS"); return e ? atoi(e) : 0; }();
  static const int target_ctas_env = [] { const char* e = getenv("BRE_TC_TARGET_CTAS"); return e ? atoi(e) : 0; }();
  const int target = target_ctas_env > 0 ? target_ctas_env : 144;   // re-tuned with the TMA producer: 96 -> 144 is +2 % on config 2, 288 is -4 %
  constexpr int kMaxCluster = BN >= 64 ? 16 : 8;  // non-portable cluster size 16 (8 is the portable limit; narrow tiles: see reduce_rows)
  int splits = a.splits;
  if (splits <= 0) {
    splits = 1;
    while (splits < kMaxCluster && tiles * splits < target && d.total_kblocks / (splits * 2) >= 2) splits *= 2;
  }
  if (max_splits_env > 0 && splits > max_splits_env) splits = max_splits_env;
  int pow2 = 1;
  while (pow2 * 2 <= splits && pow2 < kMaxCluster) pow2 *= 2;
  splits = pow2;
  while (splits > 1 && splits > d.total_kblocks) splits /= 2;
  d.kblocks_per_split = ceil_div(d.total_kblocks, splits);
  if (tn > 65535) { set_error("igemm_tc: grid too large"); return -1; }
  // (an 8-deep ring for single-wave launches was tried: 192 KB of shared memory per CTA stops the next kernel's CTAs from
  // becoming resident during this one's tail -- programmatic dependent launch loses its overlap -- and layer4 got 2x slower)
  // Ring depth: 4 stages (96 KB: two CTAs per SM, and the next kernel's CTAs become resident during this one's tail -- what
  // programmatic dependent launch needs).  Measured on the B200 (profiles/README.md): an 8-deep ring (192 KB, one CTA per SM) is
  // 25-30 % slower on config 2 *and* on the multi-wave batch-8 launches of config 3 -- occupancy beats ring depth.
  // BRE_TC_STAGES=2|4|8 forces a depth for experiments.
  static const int stages_env = [] { const char* e = getenv("BRE_TC_STAGES"); return e ? atoi(e) : 0; }();
  // Short reductions on many tiles (the token models' decoder fprop: 786 tiles x 3-6 k-blocks): a 2-deep ring halves the shared
  // memory so four CTAs share an SM and the per-CTA prologue / epilogue overlap: config 5 1187 -> 1220 it/s (BRE_TC_SHORTK_STAGES=4 turns it off)
  static const int shortk_env = [] { const char* e = getenv("BRE_TC_SHORTK_STAGES"); return e ? atoi(e) : 2; }();
  const bool shortk = shortk_env == 2 && d.kblocks_per_split <= 6 && tm == 1 && tiles * splits > 2LL * kNumSMs;   // (with full 128-row tiles -- config 3's 1x1 convolutions -- it costs 1 %)
  const int stages = (stages_env == 8 || stages_env == 2) ? stages_env : (shortk ? 2 : TC_STAGES);
  d.stage_shift = stages == 8 ? 3 : (stages == 2 ? 1 : 2);
  const size_t smem = (size_t)stages * (TC_BM + BN) * TC_BK * 4;
  const size_t smem_max = (size_t)TC_MAX_STAGES * (TC_BM + BN) * TC_BK * 4;
  static bool attr_done = false;
  if (!attr_done) {
    BRE_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc_kernel<MODE, BN, TMA, CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    BRE_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc_kernel<MODE, BN, TMA, CLS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_done = true;
  }
  static const int proxy_fence_env = [] {
    const char* e = getenv("BRE_TC_PROXY_FENCE");
    const char* pf = getenv("BRE_TC_PREFETCH");
    const char* np = getenv("BRE_TC_PRODUCERS");
    int nprod = np ? atoi(np) : 2;
    if (nprod < 1 || nprod > 4) nprod = 2;
    return (e && atoi(e) ? 1 : 0) | ((pf ? atoi(pf) : 1) ? 2 : 0) | (nprod << 2);
  }();
  {
    cudaError_t lerr = launch_kernel(igemm_tc_kernel<MODE, BN, TMA, CLS>, dim3(tm, tn, splits), dim3(TC_BLOCK), smem, stream, splits, a, d,
                                     proxy_fence_env, maps, plan);
    if (lerr != cudaSuccess) { set_error(std::string("igemm_tc launch failed: ") + cudaGetErrorString(lerr)); return -2; }
  }
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace

static bool narrow_tiles_ok(const GemmArgs& a) {
  static const int env = [] { const char* e = getenv("BRE_TC_NARROW"); return e ? atoi(e) : 1; }();
  if (!env || !tma_eligible(a)) return false;
  if (a.mode == GEMM_DGRAD && a.g.stride != 1) return false;
  if (a.mode == GEMM_WGRAD && (a.g.Ci % 32 != 0)) return false;
  return true;
}

bool igemm_tc_supported(const GemmArgs& a) {
  const ConvGeom& g = a.g;
  int M, Nc, K;
  gemm_dims(a, M, Nc, K);
  for (int s = 0; s < a.nsrc; ++s)
    if (!aligned16(a.act[s]) || !aligned16(a.wgt[s])) return false;
  if (!aligned16(a.out) || (a.splits == 0 && a.ws == nullptr)) return false;
  const bool x_nhwc = a.x_sC == 1 && a.x_sP % 4 == 0 && a.x_sN % 4 == 0;
  // output tiles are 128 x 64; widths that are only a multiple of 32 (token models: d = 96, 3 d = 288) run 128 x 32 tiles, which
  // exist for the TMA producer only
  if (Nc % 64 != 0 && !(Nc % 32 == 0 && narrow_tiles_ok(a))) return false;
  switch (a.mode) {
    case GEMM_FPROP: return x_nhwc && g.Ci % TC_BK == 0 && g.R * g.S <= 64;  // k-block inside one (r, s) cell
    case GEMM_DGRAD: return x_nhwc && g.Co % TC_BK == 0 && g.R * g.S <= 64;   // (Nc = Ci: tile-width rule above)
    case GEMM_WGRAD: return x_nhwc && g.Co % 4 == 0 && g.Ci % 4 == 0 && g.R * g.S <= 64;
    default: return false;
  }
}

int launch_igemm_tc(const GemmArgs& a, cudaStream_t stream) {
  if (!igemm_tc_supported(a)) { set_error("igemm_tc: unsupported shape"); return -4; }
  TcDims d;
  gemm_dims(a, d.M, d.Nc, d.K);
  d.kblocks_per_src = ceil_div(d.K, TC_BK);
  d.total_kblocks = d.kblocks_per_src * a.nsrc;
  d.kblocks_per_split = d.total_kblocks;
  // TMA producer where the geometry allows it and the driver encodes the maps; otherwise the cp.async producer of the same
  // kernel (still tcgen05, still on the GPU: a different loader, not a fallback to another implementation)
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (cls_eligible(a)) {   // strided dgrad: per-parity-class gathers through im2col tensor maps
    ClsPlan plan;
    TcMapsCls cmaps;
    memset(&cmaps, 0, sizeof(cmaps));
    if (build_cls(a, &plan, &cmaps)) return launch_tc<GEMM_DGRAD, 64, true, true>(a, d, cmaps, stream, &plan);
  }
  if (d.Nc % 64 != 0) {   // 128 x 32 tiles (TMA producer only, see igemm_tc_supported)
    if (!build_maps(a, 32, &maps)) { set_error("igemm_tc: tensor-map encoding failed for a narrow-tile shape"); return -4; }
    if (a.mode == GEMM_FPROP) return launch_tc<GEMM_FPROP, 32, true>(a, d, maps, stream);
    if (a.mode == GEMM_DGRAD) return launch_tc<GEMM_DGRAD, 32, true>(a, d, maps, stream);
    return launch_tc<GEMM_WGRAD, 32, true>(a, d, maps, stream);
  }
  const bool tma = tma_eligible(a) && build_maps(a, 64, &maps);
  if (a.mode == GEMM_FPROP) return tma ? launch_tc<GEMM_FPROP, 64, true>(a, d, maps, stream) : launch_tc<GEMM_FPROP, 64, false>(a, d, maps, stream);
  if (a.mode == GEMM_DGRAD) return tma ? launch_tc<GEMM_DGRAD, 64, true>(a, d, maps, stream) : launch_tc<GEMM_DGRAD, 64, false>(a, d, maps, stream);
  return tma ? launch_tc<GEMM_WGRAD, 64, true>(a, d, maps, stream) : launch_tc<GEMM_WGRAD, 64, false>(a, d, maps, stream);
}

}  // namespace bre

#ifdef BRE_TC_TRACE
extern "C" int bre_debug_tc_trace(long long* out16) {
  return cudaMemcpyFromSymbol(out16, bre::g_tc_trace, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif
