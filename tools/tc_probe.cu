// tc_probe.cu -- standalone check of the tcgen05 building blocks used by the bf16 gradient kernel:
//   2-D TMA tensor copies (128B swizzle) of a [KR rows][d] bf16 tile into [d/64][KR][64] shared memory,
//   an MN-major SW128 A descriptor over that tile (A = X^T chunk: 128 features x KR rows),
//   a hand-written no-swizzle K-major B operand (the bf16 pieces of r: N=16 x K=KR),
//   tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), tcgen05.commit, tcgen05.ld.
// Output: G[f][n] = sum_k X[k][f] * R[n][k], compared with the CPU.   nvcc -arch=sm_100a tc_probe.cu -o tc_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int KR = 16;     // rows per tile = MMA K
constexpr int D = 512;     // features
constexpr int NB = D / 64; // 64-feature blocks
constexpr int NCH = D / 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmap, const float *r_pieces /*[16][KR]*/,
                                             float *out /*[D][16]*/) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *tile = smem;                              // [NB][KR][64] bf16, 128B-swizzled: NB * KR * 128 bytes
  unsigned char *b2 = smem + NB * KR * 128;                // 512 B: N=16 x K=16, no-swizzle K-major core matrices
  uint64_t *bars = reinterpret_cast<uint64_t *>(b2 + 512); // [0] = tma full, [1] = mma done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM: NCH chunks x 16 columns, at least 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // B operand: element (n, k) at  (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2   (SBO = 256 B, LBO = 128 B)
  for (int i = tid; i < 16 * KR; i += 128) {
    const int n = i / KR, k = i % KR;
    *reinterpret_cast<__nv_bfloat16 *>(b2 + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) =
        __float2bfloat16_rn(r_pieces[n * KR + k]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(NB * KR * 128) : "memory");
    for (int b = 0; b < NB; ++b) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                   "r"(smem_u32(tile + b * KR * 128)), "l"(&tmap), "r"(b * 64), "r"(0), "r"(smem_u32(&bars[0])) : "memory");
    }
  }
  // everyone waits for the tile
  {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bars[0])), "r"(0) : "memory");
  }
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    // instruction descriptor: c=F32(1)<<4, a=BF16(1)<<7, b=BF16(1)<<10, a_major=MN(1)<<15, b_major=K(0)<<16, N>>3 <<17, M>>4 <<24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    // B descriptor (no swizzle, K-major): start, LBO = 128 B, SBO = 256 B, version 1
    const uint64_t bdesc = (uint64_t)((smem_u32(b2) & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
    for (int c = 0; c < NCH; ++c) {
      // A descriptor (SW128, MN-major): 2 blocks of 64 features; LBO = block stride = KR*128 B, SBO = 8-row group stride = 1024 B
      const uint32_t a_addr = smem_u32(tile + (2 * c) * KR * 128);
      const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)((KR * 128) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                             (1ull << 46) | (2ull << 61);
      const uint32_t taddr = tmem_base + c * 16;
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::
                   "r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bars[1])), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c = 0; c < NCH; ++c) {
    uint32_t v[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int f = c * 128 + warp * 32 + lane;
    for (int n = 0; n < 16; ++n) out[f * 16 + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
}

int main() {
  const int rows = 64;
  __nv_bfloat16 *hX = (__nv_bfloat16 *)malloc(sizeof(__nv_bfloat16) * rows * D);
  float *hXf = (float *)malloc(sizeof(float) * rows * D);
  srand(1);
  for (int i = 0; i < rows * D; ++i) {
    float v = (float)(rand() % 2001 - 1000) / 512.0f;
    hX[i] = __float2bfloat16_rn(v);
    hXf[i] = __bfloat162float(hX[i]);
  }
  float hR[16 * KR];
  for (int n = 0; n < 16; ++n)
    for (int k = 0; k < KR; ++k) hR[n * KR + k] = (n < 3) ? __bfloat162float(__float2bfloat16_rn((float)(rand() % 401 - 200) / 64.0f / (1 << (8 * n)))) : 0.f;
  __nv_bfloat16 *dX; float *dR, *dOut;
  CK(cudaMalloc(&dX, sizeof(__nv_bfloat16) * rows * D));
  CK(cudaMalloc(&dR, sizeof(hR)));
  CK(cudaMalloc(&dOut, sizeof(float) * D * 16));
  CK(cudaMemcpy(dX, hX, sizeof(__nv_bfloat16) * rows * D, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, hR, sizeof(hR), cudaMemcpyHostToDevice));
  // tensor map over X [rows][D] bf16: dims {D, rows}, box {64, KR}, 128B swizzle
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, KR};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
  const int smem_bytes = NB * KR * 128 + 512 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe<<<1, 128, smem_bytes>>>(tmap, dR, dOut);
  CK(cudaDeviceSynchronize());
  float *hOut = (float *)malloc(sizeof(float) * D * 16);
  CK(cudaMemcpy(hOut, dOut, sizeof(float) * D * 16, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  int bad = 0;
  for (int f = 0; f < D; ++f)
    for (int n = 0; n < 16; ++n) {
      double ref = 0;
      for (int k = 0; k < KR; ++k) ref += (double)hXf[k * D + f] * hR[n * KR + k];
      const double err = fabs(ref - hOut[f * 16 + n]);
      if (err > maxerr) maxerr = err;
      if (fabs(ref) > maxref) maxref = fabs(ref);
      if (err > 1e-3 * (1 + fabs(ref)) && bad < 8) { printf("mismatch f=%d n=%d got %g want %g\n", f, n, hOut[f * 16 + n], ref); ++bad; }
    }
  printf("tc_probe: max abs err %.3g (max |ref| %.3g) -> %s\n", maxerr, maxref, maxerr < 1e-3 * (1 + maxref) ? "OK" : "FAIL");
  return maxerr < 1e-3 * (1 + maxref) ? 0 : 1;
}
