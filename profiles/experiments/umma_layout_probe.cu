// Probe how tcgen05.mma (kind::tf32, cta_group::1, M=128, N=64, K=8) maps shared-memory bytes to B(n, k) for an
// MN-major / K-major no-swizzle descriptor: put a single 1.0f at byte offset `off` of the B region, use A(m, k) = k + 1
// (K-major, known-good layout), run one MMA and read D[0][:]; D[0][n] = k + 1 identifies (n, k).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_layout_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype = 0) {
  uint64_t d = (uint64_t)(ltype & 7) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void probe(int b_mn, uint32_t lbo, uint32_t sbo, int region_bytes, int step, int* out_n, float* out_v, uint32_t ltype) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sA = (float*)smem;                 // 128 x 8 K-major: byte(row,k) = (k/4)*2048 + row*16 + (k%4)*4
  uint8_t* sB = smem + 4096;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int k = 0; k < 8; ++k) *(float*)((uint8_t*)sA + (k / 4) * 2048 + tid * 16 + (k % 4) * 4) = (float)(k + 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0;
  int idx = 0;
  for (int off = 0; off < region_bytes; off += step, ++idx) {
    for (int i = tid; i < region_bytes / 4; i += 128) ((float*)sB)[i] = 0.f;
    __syncthreads();
    if (tid == 0) *(float*)(sB + off) = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = make_desc(smem_u32(sA), 2048, 128), bd = make_desc(smem_u32(sB), lbo, sbo, ltype);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
          : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
      if (spin > (1u << 22)) __trap();
    }
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
      uint32_t r[32];
      int found_n = -1; float found_v = 0.f; int count = 0;
      for (int c = 0; c < 64; c += 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(r[j]);
          if (v != 0.f) { found_n = c + j; found_v = v; ++count; }
        }
      }
      if (tid == 0) { out_n[idx] = count == 1 ? found_n : (count == 0 ? -1 : -100 - count); out_v[idx] = found_v; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main(int argc, char** argv) {
  int b_mn = argc > 1 ? atoi(argv[1]) : 1;
  uint32_t lbo = argc > 2 ? atoi(argv[2]) : 2048, sbo = argc > 3 ? atoi(argv[3]) : 128;
  int region = argc > 4 ? atoi(argv[4]) : 8192, step = argc > 5 ? atoi(argv[5]) : 4;
  uint32_t ltype = argc > 6 ? atoi(argv[6]) : 0;
  int n = region / step;
  int* dn; float* dv;
  cudaMalloc(&dn, n * sizeof(int)); cudaMalloc(&dv, n * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + region + 1024);
  probe<<<1, 128, 4096 + region>>>(b_mn, lbo, sbo, region, step, dn, dv, ltype);
  cudaError_t e = cudaDeviceSynchronize();
  printf("# b_mn=%d lbo=%u sbo=%u region=%d step=%d ltype=%u -> %s\n", b_mn, lbo, sbo, region, step, ltype, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  int* hn = (int*)malloc(n * sizeof(int)); float* hv = (float*)malloc(n * sizeof(float));
  cudaMemcpy(hn, dn, n * sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(hv, dv, n * sizeof(float), cudaMemcpyDeviceToHost);
  int shown = 0;
  for (int i = 0; i < n; ++i)
    if (hn[i] != -1) { printf("off %5d -> n=%3d k=%g\n", i * step, hn[i], hv[i] - 1.f); if (++shown >= 4096) break; }
  printf("# %d offsets mapped\n", shown);
  return 0;
}
