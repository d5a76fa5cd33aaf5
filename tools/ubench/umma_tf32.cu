// Stand-alone check of the tcgen05 (UMMA) building blocks the DS-CNN pointwise kernel uses: D[128 x 64] = A[128 x 64] * B^T,
// B given as [n = 64][k = 64], fp32 inputs, 3xTF32 split products (hi*hi + hi*lo + lo*hi) accumulated in TMEM, read back with
// tcgen05.ld.  Operands sit in shared memory in the canonical K-major no-swizzle layout
//   element (row, k) at float offset (k / 4) * LBO_f + row * 4 + k % 4       (LBO_f = stride between 4-wide K chunks),
// i.e. 8-row core matrices of 128 contiguous bytes, SBO = 128 B between 8-row groups, LBO free (padded against bank conflicts).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32 umma_tf32.cu ; prints the max error against fp64.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;        // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;        // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                  // descriptor version 1 (sm_100)
  return d;                                                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128, 1) umma_test(const float* A, const float* B, float* D, int lbo_f, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128, N = 64, K = 64;
  float* a_hi = sm;                       // [16 chunks][lbo_f]
  float* a_lo = a_hi + 16 * lbo_f;
  float* b_hi = a_lo + 16 * lbo_f;        // [16 chunks][N*4 + 4]
  const int lbo_b = N * 4 + 4;
  float* b_lo = b_hi + 16 * lbo_b;
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    const float x = A[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = (k >> 2) * lbo_f + m * 4 + (k & 3);
    a_hi[o] = hi;
    a_lo[o] = x - hi;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = (k >> 2) * lbo_b + n * 4 + (k & 3);
    b_hi[o] = hi;
    b_lo[o] = x - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                        // one warp allocates 64 TMEM columns (fp32 accumulator 128 lanes x 64 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of the operands -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    int first = 1;
    for (int pass = 0; pass < 3; ++pass) {                 // lo*hi, hi*lo, hi*hi (small terms first)
      const float* ap = pass == 0 ? a_lo : a_hi;
      const float* bp = pass == 1 ? b_lo : b_hi;
      if (mode == 1 && pass != 2) continue;                // mode 1: single TF32 product only (to see the split's effect)
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = make_desc(smem_u32(ap + 2 * ks * lbo_f), lbo_f * 4, 128);
        const uint64_t db = make_desc(smem_u32(bp + 2 * ks * lbo_b), lbo_b * 4, 128);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32w .. 32w+31 (row m = 32 w + lane), 64 columns in 4 loads of 16
  float v[64];
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c0 + j] = __uint_as_float(r[j]);
  }
  const int m = warp * 32 + (tid & 31);
  for (int n = 0; n < N; ++n) D[(size_t)blockIdx.x * M * N + m * N + n] = v[n];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main(int argc, char** argv) {
  const int M = 128, N = 64, K = 64;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    for (int lbo_f : {M * 4, M * 4 + 4}) {
      const size_t smem = (size_t)(2 * 16 * lbo_f + 2 * 16 * (N * 4 + 4)) * 4;
      cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaMemset(dD, 0xFF, D.size() * 4);
      umma_test<<<1, 128, smem>>>(dA, dB, dD, lbo_f, mode);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double r = 0;
          for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * (double)B[n * K + k];
          maxerr = fmax(maxerr, fabs(r - (double)D[m * N + n]));
          maxref = fmax(maxref, fabs(r));
        }
      printf("mode %d (%s) lbo_f %d: %s  max|err| %.3e  max|ref| %.3f  rel %.3e   D[0][0..3] = %g %g %g %g\n", mode,
             mode == 0 ? "3xTF32" : "1xTF32", lbo_f, cudaGetErrorString(e), maxerr, maxref, maxerr / maxref, D[0], D[1], D[2], D[3]);
    }
  }
  return 0;
}
