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

// ---- reduce-scatter + all-gather for large payloads (see agd_common.cuh)
// step 1: slice p of this rank's partial sums -> slot my_rank of rank p's rs area; then the "arrived" flag on every peer
__global__ void __launch_bounds__(256) xchg_rs_publish_kernel(const double *__restrict__ acc, const XchgRs x) {
  __shared__ bool last;
  const int W = x.world, S = x.slot_stride;
  const size_t L = ((size_t)S + W - 1) / W;                      // slot capacity
  const int l = (x.n + W - 1) / W;                               // slice length of this sweep
  const size_t off_rs = 2 * (size_t)W * S;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < x.n; c += gridDim.x * 256) {
    const int p = c / l;
    x.peers.slot[p][off_rs + ((size_t)x.buf * W + x.my_rank) * L + (size_t)(c - p * l)] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(x.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x < W) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(&x.peers.flag[threadIdx.x][2 * W + x.buf * W + x.my_rank]) = x.epoch;
    }
    if (threadIdx.x == 0) *x.ticket = 0u;
  }
}

// step 2: wait for the W contributions to MY slice, add them in rank order, store the finished slice into every rank's res area
__global__ void __launch_bounds__(256) xchg_rs_reduce_bcast_kernel(const double *xbuf, const unsigned long long *flags, const XchgRs x) {
  __shared__ bool last;
  const int W = x.world, S = x.slot_stride;
  const size_t L = ((size_t)S + W - 1) / W;
  const int l = (x.n + W - 1) / W;
  const size_t off_rs = 2 * (size_t)W * S, off_res = off_rs + 2 * (size_t)W * L;
  if (threadIdx.x < W) {
    const volatile unsigned long long *f = flags + 2 * W + x.buf * W + threadIdx.x;
    while (*f < x.epoch) __nanosleep(20);
    __threadfence_system();
  }
  __syncthreads();
  const int c0 = x.my_rank * l;
  int len = x.n - c0;
  if (len > l) len = l;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < len; i += gridDim.x * 256) {
    double s = 0.0;
    for (int r = 0; r < W; ++r) s += __ldcg(xbuf + off_rs + ((size_t)x.buf * W + r) * L + i);   // rank order: identical bits everywhere
    const size_t dst = off_res + (size_t)x.buf * S + (size_t)(c0 + i);
    for (int q = 0; q < W; ++q) x.peers.slot[q][dst] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(x.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x < W) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(&x.peers.flag[threadIdx.x][4 * W + x.buf * W + x.my_rank]) = x.epoch;
    }
    if (threadIdx.x == 0) *x.ticket = 0u;
  }
}

// step 3 (stand-alone form; K3 kernels inline it): wait for the W finished slices, copy them out
__global__ void __launch_bounds__(256) xchg_rs_gather_kernel(const double *xbuf, const unsigned long long *flags, int world, int buf,
                                                             int n, int slot_stride, unsigned long long epoch, double *acc_out) {
  if (threadIdx.x < world) {
    const volatile unsigned long long *f = flags + 4 * world + buf * world + threadIdx.x;
    while (*f < epoch) __nanosleep(20);
    __threadfence_system();
  }
  __syncthreads();
  const size_t L = ((size_t)slot_stride + world - 1) / world;
  const double *res = xbuf + 2 * (size_t)world * slot_stride + 2 * (size_t)world * L + (size_t)buf * slot_stride;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < n; c += gridDim.x * 256) acc_out[c] = __ldcg(res + c);
}

}  // namespace

cudaError_t xchg_rs_publish_launch(const double *acc, const XchgRs &x, cudaStream_t st) {
  int grid = (x.n + 255) / 256;
  if (grid > 296) grid = 296;
  xchg_rs_publish_kernel<<<grid, 256, 0, st>>>(acc, x);
  return cudaGetLastError();
}
cudaError_t xchg_rs_reduce_bcast_launch(const double *xbuf_local, const unsigned long long *flags_local, const XchgRs &x, cudaStream_t st) {
  const int l = (x.n + x.world - 1) / x.world;
  int grid = (l + 255) / 256;
  if (grid > 148) grid = 148;
  if (grid < 1) grid = 1;
  xchg_rs_reduce_bcast_kernel<<<grid, 256, 0, st>>>(xbuf_local, flags_local, x);
  return cudaGetLastError();
}
cudaError_t xchg_rs_gather_launch(const double *xbuf_local, const unsigned long long *flags_local, int world, int buf, int n,
                                  int slot_stride, unsigned long long epoch, double *acc_out, cudaStream_t st) {
  int grid = (n + 255) / 256;
  if (grid > 296) grid = 296;
  xchg_rs_gather_kernel<<<grid, 256, 0, st>>>(xbuf_local, flags_local, world, buf, n, slot_stride, epoch, acc_out);
  return cudaGetLastError();
}

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
