// xchg.cu -- K2': the one-shot all-reduce over NVLink peer memory (see agd_common.cuh for the protocol).
#include <cuda_runtime.h>
#include <stdint.h>

#include "agd_common.cuh"

namespace agd {

namespace {

// standalone publish (the CSR path has no slab reduction to fuse it into)
__global__ void __launch_bounds__(256) xchg_publish_kernel(const double *__restrict__ acc, const XchgPub pub) {
  __shared__ bool last;
  const size_t base = ((size_t)pub.buf * pub.world + pub.my_rank) * pub.slot_stride;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < pub.n; c += gridDim.x * 256) {
    const double v = acc[c];
    for (int p = 0; p < pub.world; ++p) pub.peers.slot[p][base + c] = v;
  }
  __syncthreads();   // one cumulative system-scope fence per block (see k1_reduce_kernel)
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(pub.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x < pub.world) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(&pub.peers.flag[threadIdx.x][pub.buf * pub.world + pub.my_rank]) = pub.epoch;
    }
    if (threadIdx.x == 0) *pub.ticket = 0u;
  }
}

// wait for the W flags of this epoch, then add the W slots in rank order (identical bits on every rank)
__global__ void __launch_bounds__(256) xchg_gather_kernel(const double *xbuf, const unsigned long long *flags, int world,
                                                          int buf, int n, int slot_stride, unsigned long long epoch,
                                                          double *acc_out) {
  if (threadIdx.x < world) {
    const volatile unsigned long long *f = flags + buf * world + threadIdx.x;
    while (*f < epoch) __nanosleep(20);
    __threadfence_system();   // acquire: the slot data this flag guards is read (ld.cg, from L2) after the CTA barrier
  }
  __syncthreads();
  for (int c = blockIdx.x * 256 + threadIdx.x; c < n; c += gridDim.x * 256) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcg(xbuf + ((size_t)buf * world + r) * slot_stride + c);  // written remotely: bypass L1
    acc_out[c] = s;
  }
}

}  // namespace

cudaError_t xchg_publish_launch(const double *acc, const XchgPub &pub, cudaStream_t st) {
  int grid = (pub.n + 255) / 256;
  if (grid > 64) grid = 64;
  xchg_publish_kernel<<<grid, 256, 0, st>>>(acc, pub);
  return cudaGetLastError();
}

cudaError_t xchg_gather_launch(const double *xbuf_local, const unsigned long long *flags_local, int world, int buf, int n,
                               int slot_stride, unsigned long long epoch, double *acc_out, cudaStream_t st) {
  int grid = (n + 255) / 256;
  if (grid > 64) grid = 64;
  xchg_gather_kernel<<<grid, 256, 0, st>>>(xbuf_local, flags_local, world, buf, n, slot_stride, epoch, acc_out);
  return cudaGetLastError();
}

}  // namespace agd
