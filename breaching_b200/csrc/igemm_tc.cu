// tcgen05 TF32 implicit-GEMM back end (sm_100a): the dense contractions of the four sweeps on the 5th-generation
// tensor cores.  D[128 x BN] accumulates in TMEM (fp32), one elected thread issues tcgen05.mma.kind::tf32 with
// both operands read from shared memory through UMMA descriptors (canonical no-swizzle core-matrix layouts, K-major
// or MN-major as the global layout of the operand dictates, so no transposed copies of activations / weights are
// ever materialised), completion is tracked with tcgen05.commit -> mbarrier, the epilogue reads the accumulator back
// with tcgen05.ld (32 lanes x 32 columns per warp).  Operands are staged global -> registers -> st.shared (the
// activation operand is an im2col gather; a plain 2-D TMA box cannot express it) in a 3-stage ring so that the
// gather of k-block i+1/i+2 overlaps the asynchronous MMAs of k-block i.  Split-K across gridDim.z with the same
// deterministic last-CTA reduction as the SIMT back end.
//
// fp32 storage everywhere; TF32 (10-bit mantissa) multiplies with fp32 accumulation -- the numeric mode cuDNN uses
// for the reference's GPU path by default (torch.backends.cudnn.allow_tf32).
#include "igemm.cuh"

namespace bre {

namespace {

constexpr int TC_BM = 128;      // UMMA M
constexpr int TC_BK = 32;       // k-block per pipeline stage (4 MMAs of K = 8)
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 128;

struct TcDims {
  int M, Nc, K;
  int kblocks_per_src, total_kblocks, kblocks_per_split;
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
//   K-major, no swizzle (layout type 0), core matrix = 8 rows x 16 B:
//       byte(row, k) = (k/4) * (ROWS*16) + row * 16 + (k%4) * 4
//       MMA covering k in [8j, 8j+8): start + 2j*LBO, LBO = ROWS*16, SBO = 128
//   MN-major for 32-bit operands only exists as SWIZZLE_128B_BASE32B (layout type 1; the plain / 16-byte-atom MN-major
//   layouts produce zeros for kind::tf32): 32 elements (128 B) contiguous along MN, rows of 128 B along k, the four
//   32-byte chunks of a row XOR-swizzled with (k % 4), k-groups of 4 at SBO, MN-groups of 32 at LBO:
//       byte(row, k) = (row/32) * 4096 + (k/4) * 512 + (k%4) * 128 + ((((row%32)/8) ^ (k%4)) * 32) + (row%8) * 4
//       MMA covering k in [8j, 8j+8): start + j*1024, LBO = 4096, SBO = 512
template <int ROWS>
__device__ __forceinline__ uint32_t off_kmajor(int row, int k) { return (uint32_t)((k >> 2) * (ROWS * 16) + row * 16 + (k & 3) * 4); }
template <int ROWS>
__device__ __forceinline__ uint32_t off_mnmajor(int row, int k) {
  return (uint32_t)((row >> 5) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((((row >> 3) & 3) ^ (k & 3)) << 5) + (row & 7) * 4);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts4(uint8_t* base, uint32_t off, float4 v) { *reinterpret_cast<float4*>(base + off) = v; }

template <int MODE, int BN>
__global__ void __launch_bounds__(TC_THREADS) igemm_tc_kernel(GemmArgs a, TcDims d) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                  // [STAGES][A_BYTES]
  uint8_t* sB = smem + TC_STAGES * A_BYTES;            // [STAGES][B_BYTES]
  __shared__ __align__(8) uint64_t bar_stage[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  __shared__ int s_last;

  const ConvGeom g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN, z = blockIdx.z;
  const int kb_begin = z * d.kblocks_per_split;
  const int kb_end = min(d.total_kblocks, kb_begin + d.kblocks_per_split);
  const int HoWo = g.Ho * g.Wo, HW = g.H * g.W;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_stage[s], 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem;

  // instruction descriptor (UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4
  constexpr uint32_t a_mn = (MODE == GEMM_WGRAD) ? 1u : 0u;
  constexpr uint32_t b_mn = (MODE == GEMM_FPROP) ? 0u : 1u;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);

  // ---- per-thread fixed decode of the gathered (activation) operand ----------------------------------------
  // FPROP / DGRAD: thread `tid` owns A row m0 + tid.   WGRAD: B rows are (r, s, c) columns, decoded below.
  bool a_valid = false;
  int a_img = 0, a_y = 0, a_x = 0;
  if (MODE == GEMM_FPROP) {
    const int m = m0 + tid;
    a_valid = m < d.M;
    if (a_valid) {
      a_img = m / HoWo;
      const int rem = m - a_img * HoWo;
      const int p = rem / g.Wo, q = rem - p * g.Wo;
      a_y = p * g.stride - g.pad;
      a_x = q * g.stride - g.pad;
    }
  } else if (MODE == GEMM_DGRAD) {
    const int m = m0 + tid;
    a_valid = m < d.M;
    if (a_valid) {
      a_img = m / HW;
      const int rem = m - a_img * HW;
      a_y = rem / g.W;
      a_x = rem - a_y * g.W;
    }
  }

  constexpr int A_VEC = TC_BM * TC_BK / 4 / TC_THREADS;  // float4 per thread per stage: 8
  constexpr int B_VEC = BN * TC_BK / 4 / TC_THREADS;     // 4 (BN = 64) or 8 (BN = 128)
  float4 ra[A_VEC], rb[B_VEC];

  auto load_block = [&](int kb) {
    const int src = kb / d.kblocks_per_src;
    const int kbase = (kb - src * d.kblocks_per_src) * TC_BK;
    const float* __restrict__ act = a.act[src];
    const float* __restrict__ wgt = a.wgt[src];
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == GEMM_FPROP) {
      // A(m, k) = in[img, y + r, x + s, c], k = (r, s, c); the 32-wide k-block lies inside one (r, s) cell (Ci % 32 == 0)
      const int rs = kbase / g.Ci, c0 = kbase - rs * g.Ci;
      const int r = rs / g.S, s = rs - r * g.S;
      const int h = a_y + r, w = a_x + s;
      const bool ok = a_valid && h >= 0 && h < g.H && w >= 0 && w < g.W;
      const float* p = act + a_img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + c0;
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) ra[j] = ok ? ldg4(p + 4 * j) : zero;
      // B(n, k) = W[n][k] (row-major [Co][K]): thread -> row tid % BN, k-chunks (tid / BN) * B_VEC ...
      const int row = tid % BN, kc0 = (tid / BN) * B_VEC;
      const float* q = wgt + (long long)(n0 + row) * d.K + kbase + kc0 * 4;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) rb[j] = ldg4(q + 4 * j);
    } else if (MODE == GEMM_DGRAD) {
      // A(m, k) = dout[img, (y + pad - r)/stride, (x + pad - s)/stride, ko], k = (r, s, ko)   (Co % 32 == 0)
      const int rs = kbase / g.Co, k0 = kbase - rs * g.Co;
      const int r = rs / g.S, s = rs - r * g.S;
      const int hp = a_y + g.pad - r, wp = a_x + g.pad - s;
      bool ok = a_valid && hp >= 0 && wp >= 0;
      int p = 0, q = 0;
      if (ok) {
        p = hp / g.stride; q = wp / g.stride;
        ok = (p * g.stride == hp) && (q * g.stride == wp) && p < g.Ho && q < g.Wo;
      }
      const float* pa = act + ((long long)(a_img * g.Ho + p) * g.Wo + q) * g.Co + k0;
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) ra[j] = ok ? ldg4(pa + 4 * j) : zero;
      // B(n = ci, k) = W[ko][r][s][ci]  -> contiguous along n: MN-major.  thread -> k = tid % 32, n-chunks (tid / 32) ...
      const int k = tid & 31, nc0 = (tid >> 5) * B_VEC;
      const float* pb = wgt + ((long long)(k0 + k) * (g.R * g.S) + rs) * g.Ci + n0 + nc0 * 4;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) rb[j] = ldg4(pb + 4 * j);
    } else {
      // WGRAD: k = pixel.  A(m = ko, k) = dout[pixel][ko] (MN-major).  thread -> k = tid % 32, m-chunks (tid / 32) * 8 ...
      const int k = tid & 31;
      const int pix = kbase + k;
      const bool kok = pix < d.K;
      const int mc0 = (tid >> 5) * A_VEC;
      const float* pa = wgt + (long long)pix * g.Co + m0 + mc0 * 4;
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) ra[j] = kok ? ldg4(pa + 4 * j) : zero;
      // B(n = (r, s, c), k) = in[img, p*stride - pad + r, q*stride - pad + s, c] (MN-major, Ci % 4 == 0)
      int img = 0, hb = 0, wb = 0;
      if (kok) {
        img = pix / HoWo;
        const int rem = pix - img * HoWo;
        const int p = rem / g.Wo, q = rem - p * g.Wo;
        hb = p * g.stride - g.pad; wb = q * g.stride - g.pad;
      }
      const int nc0 = (tid >> 5) * B_VEC;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) {
        const int n = n0 + (nc0 + j) * 4;
        const int rs = n / g.Ci, c = n - rs * g.Ci;
        const int r = rs / g.S, s = rs - r * g.S;
        const int h = hb + r, w = wb + s;
        const bool ok = kok && h >= 0 && h < g.H && w >= 0 && w < g.W;
        rb[j] = ok ? ldg4(act + img * a.x_sN + (long long)(h * g.W + w) * a.x_sP + c) : zero;
      }
    }
  };

  auto store_block = [&](int stage) {
    uint8_t* pa = sA + stage * A_BYTES;
    uint8_t* pb = sB + stage * B_BYTES;
    if (MODE == GEMM_FPROP) {
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) sts4(pa, off_kmajor<TC_BM>(tid, 4 * j), ra[j]);
      const int row = tid % BN, kc0 = (tid / BN) * B_VEC;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) sts4(pb, off_kmajor<BN>(row, 4 * (kc0 + j)), rb[j]);
    } else if (MODE == GEMM_DGRAD) {
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) sts4(pa, off_kmajor<TC_BM>(tid, 4 * j), ra[j]);
      const int k = tid & 31, nc0 = (tid >> 5) * B_VEC;
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) sts4(pb, off_mnmajor<BN>(4 * (nc0 + j), k), rb[j]);
    } else {
      const int k = tid & 31, mc0 = (tid >> 5) * A_VEC, nc0 = (tid >> 5) * B_VEC;
#pragma unroll
      for (int j = 0; j < A_VEC; ++j) sts4(pa, off_mnmajor<TC_BM>(4 * (mc0 + j), k), ra[j]);
#pragma unroll
      for (int j = 0; j < B_VEC; ++j) sts4(pb, off_mnmajor<BN>(4 * (nc0 + j), k), rb[j]);
    }
  };

  // ---- main loop: 3-stage ring, MMAs are asynchronous, completion frees the stage through its mbarrier ----------
  const int nkb = kb_end - kb_begin;
  for (int i = 0; i < nkb; ++i) {
    const int stage = i % TC_STAGES;
    load_block(kb_begin + i);                                      // global loads in flight while earlier MMAs run
    if (i >= TC_STAGES) mbar_wait(&bar_stage[stage], (uint32_t)((i / TC_STAGES - 1) & 1));
    store_block(stage);
    fence_proxy_async();                                           // generic-proxy st.shared -> visible to the MMA's async proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(sA + stage * A_BYTES), sb = smem_u32(sB + stage * B_BYTES);
#pragma unroll
      for (int j = 0; j < TC_BK / 8; ++j) {
        const uint64_t adesc = a_mn ? make_desc(sa + j * 1024, 4096, 512, 1) : make_desc(sa + 2 * j * (TC_BM * 16), TC_BM * 16, 128, 0);
        const uint64_t bdesc = b_mn ? make_desc(sb + j * 1024, 4096, 512, 1) : make_desc(sb + 2 * j * (BN * 16), BN * 16, 128, 0);
        umma_tf32(tmem_d, adesc, bdesc, idesc, (i > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(&bar_stage[stage]);
    }
  }
  if (tid == 0) umma_commit(&bar_done);
  mbar_wait(&bar_done, 0);
  tc_fence_after();

  // ---- epilogue: TMEM -> registers; thread `tid` holds row m0 + tid, BN columns in chunks of 32 -----------------
  const int splits = gridDim.z;
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
  float v[32];
  if (splits > 1) {
    float* wsb = a.ws + ((long long)tile * splits + z) * (TC_BM * BN) + (long long)tid * BN;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      if (nkb > 0) tmem_ld32(lane_base + c, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(wsb + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(a.counters + tile, 1);
      s_last = (prev == splits - 1);
      if (s_last) a.counters[tile] = 0;
    }
    __syncthreads();
  }
  const bool writer = (splits == 1) || s_last;
  if (writer) {
    if (splits > 1) __threadfence();
    const int m = m0 + tid;
    const bool row_ok = m < d.M;
    long long row = 0;
    int cs = 1;
    if (MODE == GEMM_DGRAD) {
      const int img = row_ok ? m / HW : 0;
      row = img * a.x_sN + (long long)(m - img * HW) * a.x_sP;
      cs = a.x_sC;
    } else {
      row = (long long)m * d.Nc;
    }
    const float* wst = a.ws + (long long)tile * splits * (TC_BM * BN) + (long long)tid * BN;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      if (splits == 1) {
        tmem_ld32(lane_base + c, v);  // warp-collective: executed by all lanes, also for rows beyond M
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        for (int zz = 0; zz < splits; ++zz) {
          const float* p = wst + (long long)zz * (TC_BM * BN) + c;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p + j));
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
        }
      }
      if (!row_ok) continue;
      const int n = n0 + c;
      if (MODE == GEMM_FPROP && a.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + n + j);
      }
      float* op = a.out + row + (long long)n * cs;
      if (cs == 1) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE, int BN>
int launch_tc(const GemmArgs& a, const TcDims& d0, cudaStream_t stream) {
  TcDims d = d0;
  const int tm = ceil_div(d.M, TC_BM), tn = d.Nc / BN;
  const long long tiles = (long long)tm * tn;
  const int ws_tiles_tc = a.ws_tiles * (IG_BM * IG_BN) / (TC_BM * BN);  // workspace is sized in 64x64 tiles
  int splits = a.splits;
  if (splits <= 0) {
    splits = 1;
    if (tiles < kNumSMs) {
      splits = ceil_div(kNumSMs, tiles);
      const int max_by_k = d.total_kblocks / 2 > 0 ? d.total_kblocks / 2 : 1;
      if (splits > max_by_k) splits = max_by_k;
    }
  }
  if (splits > d.total_kblocks) splits = d.total_kblocks;
  if (splits > 1 && (a.ws == nullptr || a.counters == nullptr)) splits = 1;
  if (splits > 1 && tiles * splits > ws_tiles_tc) splits = (int)(ws_tiles_tc / tiles) > 1 ? (int)(ws_tiles_tc / tiles) : 1;
  d.kblocks_per_split = ceil_div(d.total_kblocks, splits);
  splits = ceil_div(d.total_kblocks, d.kblocks_per_split);
  if (tn > 65535 || splits > 65535) { set_error("igemm_tc: grid too large"); return -1; }
  const size_t smem = (size_t)TC_STAGES * (TC_BM + BN) * TC_BK * 4;
  static bool attr_done = false;
  if (!attr_done) {
    BRE_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(tm, tn, splits), block(TC_THREADS);
  igemm_tc_kernel<MODE, BN><<<grid, block, smem, stream>>>(a, d);
  BRE_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool igemm_tc_supported(const GemmArgs& a) {
  const ConvGeom& g = a.g;
  int M, Nc, K;
  gemm_dims(a, M, Nc, K);
  for (int s = 0; s < a.nsrc; ++s)
    if (!aligned16(a.act[s]) || !aligned16(a.wgt[s])) return false;
  if (!aligned16(a.out) || (a.splits == 0 && a.ws == nullptr)) return false;
  const bool x_nhwc = a.x_sC == 1 && a.x_sP % 4 == 0 && a.x_sN % 4 == 0;
  if (Nc % 64 != 0) return false;
  switch (a.mode) {
    case GEMM_FPROP: return x_nhwc && g.Ci % TC_BK == 0;                 // k-block inside one (r, s) cell
    case GEMM_DGRAD: return x_nhwc && g.Co % TC_BK == 0 && g.Ci % 64 == 0;
    case GEMM_WGRAD: return x_nhwc && g.Co % TC_BM == 0 && g.Ci % 4 == 0 && g.Co % 4 == 0;
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
  if (a.mode == GEMM_FPROP) return launch_tc<GEMM_FPROP, 64>(a, d, stream);
  if (a.mode == GEMM_DGRAD) return launch_tc<GEMM_DGRAD, 64>(a, d, stream);
  return launch_tc<GEMM_WGRAD, 64>(a, d, stream);
}

}  // namespace bre
