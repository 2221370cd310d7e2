// k1_csr_tiles.cu -- K1 for CSR shards in a column-blocked, row-tiled layout (sm_100a).
//
// The row-major CSR kernel (k1_csr.cu) gathers w[col] and scatters into g[col] once per stored entry; with d = 10^6 and
// uniformly spread columns both are random 8-byte accesses into L2-resident vectors, i.e. two 32-byte L2 sector operations
// per entry, and that -- not HBM -- bounds it (ncu, round 1: DRAM 9 %, L2 66 %).  Here the entries of a shard are re-ordered
// once, at load time, into TILES = (row tile of 65536 rows) x (column block of 8192 columns), row-major inside a tile, so
// that the column-indexed side of both contractions lives in SHARED MEMORY:
//   kernel A  (margins, BLAS.dot of the sparse branch of Gradient.compute, call site AGD.scala:198): a CTA stages the tile's
//             w block (64 KB) with one TMA bulk copy, multiplies every entry by w_blk[lcol] out of shared memory and adds the
//             product into m[row] -- rows ascend inside a tile, so these fp64 REDs are coalesced into few L2 sectors;
//   kernel A2 (one thread per row): (loss', loss) from the finished margin, row mask, loss / count sums;
//   kernel B  (BLAS.axpy into cumGradient): g_blk[lcol] += mult[row] * val accumulates in shared memory (mult[row] is a
//             coalesced read of an L2-resident tile), then ONE coalesced flush of the block's non-zero sums per tile.
// Tiles are visited row tile by row tile (all column blocks of a row tile before the next), so the 512 KB slices of m / mult
// a row tile touches stay in L2 while its ~123 column blocks are processed by different CTAs.
// HBM traffic per pass: the 8-byte entries twice (once per contraction) + 24 bytes per row.  Sums are formed in a
// different order than in the row-major kernel (fp64, equal to rounding).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "agd_common.cuh"
#include "k1_device.cuh"

namespace agd {

namespace {

constexpr int kRT = 65536;     // rows per row tile (16-bit local row)
constexpr int kCB = 8192;      // columns per column block (13-bit local column): 64 KB of fp64 in shared memory
constexpr int kCBShift = 13;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- build: tile id of every entry (row tile major, column block minor); the stable sort by this key keeps the row-major
// order of the CSR stream inside each tile
__global__ void __launch_bounds__(256) tile_keys_kernel(const long long *rowptr, const int *idx, long long rows, int nb,
                                                       uint32_t *keys, uint32_t *vals) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * 256LL + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * 256LL) >> 5;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const long long lo = rowptr[r], hi = rowptr[r + 1];
    const uint32_t rt = (uint32_t)(r / kRT);
    for (long long k = lo + lane; k < hi; k += 32) {
      keys[k] = rt * (uint32_t)nb + ((uint32_t)idx[k] >> kCBShift);
      vals[k] = (uint32_t)k;
    }
  }
}

// rows of the entries, expanded once (entry -> row): needed to pack the local row of each sorted entry
__global__ void __launch_bounds__(256) entry_rows_kernel(const long long *rowptr, long long rows, uint32_t *erow) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * 256LL + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * 256LL) >> 5;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const long long lo = rowptr[r], hi = rowptr[r + 1];
    for (long long k = lo + lane; k < hi; k += 32) erow[k] = (uint32_t)(r % kRT);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) tile_gather_kernel(const uint32_t *sorted_keys, const uint32_t *sorted_src, const int *idx,
                                                         const T *val, const uint32_t *erow, long long nnz, long long ntiles,
                                                         uint32_t *pk, T *tval, long long *tile_ptr) {
  for (long long q = blockIdx.x * 256LL + threadIdx.x; q < nnz; q += (long long)gridDim.x * 256LL) {
    const uint32_t src = sorted_src[q];
    pk[q] = (erow[src] << kCBShift) | ((uint32_t)idx[src] & (uint32_t)(kCB - 1));
    tval[q] = val[src];
    // tile boundaries: tile_ptr[t] = first entry whose key is >= t
    const uint32_t key = sorted_keys[q];
    const uint32_t prev = q == 0 ? 0u : sorted_keys[q - 1];
    if (q == 0) for (uint32_t t = 0; t <= key; ++t) tile_ptr[t] = 0;
    else if (key != prev) for (uint32_t t = prev + 1; t <= key; ++t) tile_ptr[t] = q;
    if (q == nnz - 1) for (long long t = (long long)key + 1; t <= ntiles; ++t) tile_ptr[t] = nnz;
  }
}

// ---- kernel A: margins.  m[row] (and m2[row] at the second point) must be zero on entry.
template <typename T, bool DUAL>
__global__ void __launch_bounds__(kThreads) csr_tile_margin_kernel(const uint32_t *__restrict__ pk, const T *__restrict__ tval,
                                                                 const long long *__restrict__ tile_ptr, long long ntiles, int nb,
                                                                 const double *__restrict__ w, const double *__restrict__ w2,
                                                                 int d, double *__restrict__ m, double *__restrict__ m2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *w_s = reinterpret_cast<double *>(smem_raw);          // [kCB]
  double *w2_s = w_s + kCB;                                    // [kCB] (DUAL)
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long lo = tile_ptr[t], hi = tile_ptr[t + 1];
    if (lo == hi) continue;                                    // uniform per CTA
    const long long rt = t / nb;
    const int cb = (int)(t - rt * nb);
    const int c0 = cb * kCB;
    const int ncol = d - c0 < kCB ? d - c0 : kCB;
    __syncthreads();                                           // the previous tile's readers are done with w_s
    if (threadIdx.x == 0) {                                    // TMA-stage this tile's block of w (and w2): one bulk copy each
      const uint32_t bytes = ((uint32_t)ncol * 8u + 15u) & ~15u;   // bulk copies move whole 16-byte units; the weight vectors carry 4 spare doubles
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes * (DUAL ? 2u : 1u)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(w_s)),
                   "l"(w + c0), "r"(bytes), "r"(bar_a) : "memory");
      if (DUAL)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(w2_s)),
                     "l"(w2 + c0), "r"(bytes), "r"(bar_a) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_W:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_W;\n\tbra WAIT_W;\n\tDONE_W:\n\t}" ::"r"(bar_a),
        "r"(phase) : "memory");
    phase ^= 1u;
    double *mrow = m + rt * kRT;
    double *m2row = DUAL ? m2 + rt * kRT : nullptr;
    for (long long e = lo + threadIdx.x; e < hi; e += kThreads) {
      const uint32_t p = pk[e];
      const double v = (double)tval[e];
      const uint32_t lcol = p & (uint32_t)(kCB - 1), lrow = p >> kCBShift;
      atomicAdd(mrow + lrow, v * w_s[lcol]);                   // rows ascend along e: coalesced REDs
      if (DUAL) atomicAdd(m2row + lrow, v * w2_s[lcol]);
    }
  }
}

// ---- kernel A2: per row, the plug-in's scalar part (loss', loss) and the sums
template <bool DUAL>
__global__ void __launch_bounds__(256) csr_rows_kernel(const double *__restrict__ m, const double *__restrict__ m2,
                                                       const double *__restrict__ labels, long long rows, int kind,
                                                       unsigned long long sample_seed, unsigned long long sample_thresh,
                                                       long long row_base, double *__restrict__ mult, double *gacc, int d) {
  __shared__ double red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double lossacc = 0.0, cntacc = 0.0, lossacc2 = 0.0;
  for (long long r = blockIdx.x * 256LL + threadIdx.x; r < rows; r += (long long)gridDim.x * 256LL) {
    const double ylab = labels[r];
    double mu, loss;
    loss_eval(kind, m[r], ylab, mu, loss);
    const bool sel = row_selected(sample_seed, sample_thresh, row_base + r);
    mult[r] = sel ? mu : 0.0;
    if (sel) {
      lossacc += loss;
      cntacc += 1.0;
      if (DUAL) {
        double mu2, loss2;
        loss_eval(kind, m2[r], ylab, mu2, loss2);
        lossacc2 += loss2;
      }
    }
  }
  for (int off = 16; off >= 1; off >>= 1) {
    lossacc += __shfl_xor_sync(0xffffffffu, lossacc, off);
    cntacc += __shfl_xor_sync(0xffffffffu, cntacc, off);
    if (DUAL) lossacc2 += __shfl_xor_sync(0xffffffffu, lossacc2, off);
  }
  if (lane == 0) { red[0][warp] = lossacc; red[1][warp] = cntacc; red[2][warp] = lossacc2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, c = 0.0, s2 = 0.0;
    for (int wi = 0; wi < 8; ++wi) { s += red[0][wi]; c += red[1][wi]; s2 += red[2][wi]; }
    atomicAdd(&gacc[d], s);
    atomicAdd(&gacc[d + 1], c);      // counts are small integers: exact in any order
    if (DUAL) {
      atomicAdd(&gacc[d + 2], s2);
      atomicAdd(&gacc[d + 3], c);
    }
  }
}

// ---- kernel B: gradient
template <typename T>
__global__ void __launch_bounds__(kThreads) csr_tile_grad_kernel(const uint32_t *__restrict__ pk, const T *__restrict__ tval,
                                                               const long long *__restrict__ tile_ptr, long long ntiles, int nb,
                                                               const double *__restrict__ mult, int d, double *gacc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *g_s = reinterpret_cast<double *>(smem_raw);          // [kCB]
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long lo = tile_ptr[t], hi = tile_ptr[t + 1];
    if (lo == hi) continue;
    const long long rt = t / nb;
    const int cb = (int)(t - rt * nb);
    const int c0 = cb * kCB;
    const int ncol = d - c0 < kCB ? d - c0 : kCB;
    for (int c = threadIdx.x; c < ncol; c += kThreads) g_s[c] = 0.0;
    __syncthreads();
    const double *mrow = mult + rt * kRT;
    for (long long e = lo + threadIdx.x; e < hi; e += kThreads) {
      const uint32_t p = pk[e];
      const double mu = mrow[p >> kCBShift];                   // rows ascend along e: coalesced, L2-resident
      if (mu != 0.0) atomicAdd(&g_s[p & (uint32_t)(kCB - 1)], mu * (double)tval[e]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncol; c += kThreads) {       // one coalesced flush per tile, untouched columns skipped
      const double v = g_s[c];
      if (v != 0.0) atomicAdd(&gacc[c0 + c], v);
    }
    __syncthreads();
  }
}

inline unsigned grid_for(long long total, int sm_count) {
  long long g = (total + 255) / 256;
  if (g > 16LL * sm_count) g = 16LL * sm_count;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

int csr_tiles_rows_per_tile() { return kRT; }
int csr_tiles_cols_per_block() { return kCB; }

// Builds the tiled twin of a CSR shard.  All buffers are allocated here (the caller owns and frees them through CsrTiles).
cudaError_t csr_tiles_build(const int64_t *rowptr, const int32_t *idx, const void *val, int elem_bytes, int64_t rows, int64_t nnz,
                            int32_t d, int sm_count, CsrTiles *out, cudaStream_t st) {
  cudaError_t e;
  const int nb = (d + kCB - 1) / kCB;
  const long long nrt = (rows + kRT - 1) / kRT;
  const long long ntiles = nrt * nb;
  if (nnz >= (1LL << 32) || ntiles >= (1LL << 31)) return cudaErrorInvalidValue;
  out->nb = nb; out->nrt = nrt; out->ntiles = ntiles; out->nnz = nnz; out->rows = rows;
  uint32_t *keys = nullptr, *keys_out = nullptr, *src = nullptr, *src_out = nullptr, *erow = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
#define AGD_TRY(x) do { e = (x); if (e != cudaSuccess) goto done; } while (0)
  AGD_TRY(cudaMalloc(&out->tile_ptr, ((size_t)ntiles + 1) * sizeof(long long)));
  AGD_TRY(cudaMemsetAsync(out->tile_ptr, 0, ((size_t)ntiles + 1) * sizeof(long long), st));
  AGD_TRY(cudaMalloc(&out->m, ((size_t)nrt * kRT) * sizeof(double)));
  AGD_TRY(cudaMalloc(&out->m2, ((size_t)nrt * kRT) * sizeof(double)));
  AGD_TRY(cudaMalloc(&out->mult, ((size_t)nrt * kRT) * sizeof(double)));
  if (nnz > 0) {
    AGD_TRY(cudaMalloc(&out->pk, (size_t)nnz * sizeof(uint32_t)));
    AGD_TRY(cudaMalloc(&out->tval, (size_t)nnz * elem_bytes));
    AGD_TRY(cudaMalloc(&keys, (size_t)nnz * 4)); AGD_TRY(cudaMalloc(&keys_out, (size_t)nnz * 4));
    AGD_TRY(cudaMalloc(&src, (size_t)nnz * 4)); AGD_TRY(cudaMalloc(&src_out, (size_t)nnz * 4));
    AGD_TRY(cudaMalloc(&erow, (size_t)nnz * 4));
    tile_keys_kernel<<<grid_for(rows * 32, sm_count), 256, 0, st>>>((const long long *)rowptr, idx, rows, nb, keys, src);
    entry_rows_kernel<<<grid_for(rows * 32, sm_count), 256, 0, st>>>((const long long *)rowptr, rows, erow);
    int end_bit = 1;
    while ((1LL << end_bit) < ntiles) ++end_bit;
    AGD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, src, src_out, (long long)nnz, 0, end_bit, st));
    AGD_TRY(cudaMalloc(&tmp, tmp_bytes));
    AGD_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, src, src_out, (long long)nnz, 0, end_bit, st));  // stable
    if (elem_bytes == 4)
      tile_gather_kernel<float><<<grid_for(nnz, sm_count), 256, 0, st>>>(keys_out, src_out, idx, (const float *)val, erow, nnz, ntiles,
                                                                          out->pk, (float *)out->tval, out->tile_ptr);
    else
      tile_gather_kernel<double><<<grid_for(nnz, sm_count), 256, 0, st>>>(keys_out, src_out, idx, (const double *)val, erow, nnz, ntiles,
                                                                           out->pk, (double *)out->tval, out->tile_ptr);
    AGD_TRY(cudaGetLastError());
  }
  AGD_TRY(cudaStreamSynchronize(st));
done:
#undef AGD_TRY
  cudaFree(keys); cudaFree(keys_out); cudaFree(src); cudaFree(src_out); cudaFree(erow); cudaFree(tmp);
  return e;
}

void csr_tiles_free(CsrTiles *t) {
  cudaFree(t->pk); cudaFree(t->tval); cudaFree(t->tile_ptr); cudaFree(t->m); cudaFree(t->m2); cudaFree(t->mult);
  *t = CsrTiles();
}

// One applySmooth over a tiled shard: gacc[0..d+3] as k1_csr_launch produces it.
cudaError_t k1_csr_tiles_launch(const K1CsrArgs &a, const CsrTiles &t, int elem_bytes, int sm_count, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.gacc, 0, ((size_t)a.d + 4) * sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (a.rows <= 0) return cudaSuccess;
  const bool dual = a.w2 != nullptr;
  e = cudaMemsetAsync(t.m, 0, (size_t)a.rows * sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (dual) { e = cudaMemsetAsync(t.m2, 0, (size_t)a.rows * sizeof(double), st); if (e != cudaSuccess) return e; }
  const int smem_a = (dual ? 2 : 1) * kCB * 8, smem_b = kCB * 8;
  long long grid = (long long)(dual ? 1 : 3) * sm_count;
  if (grid > t.ntiles) grid = t.ntiles;
  if (grid < 1) grid = 1;
  long long grid_b = 3LL * sm_count;
  if (grid_b > t.ntiles) grid_b = t.ntiles;
  if (grid_b < 1) grid_b = 1;
#define AGD_LAUNCH_A(T, D)                                                                                                 \
  do {                                                                                                                     \
    e = cudaFuncSetAttribute(csr_tile_margin_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);           \
    if (e != cudaSuccess) return e;                                                                                        \
    csr_tile_margin_kernel<T, D><<<(unsigned)grid, kThreads, smem_a, st>>>(t.pk, (const T *)t.tval, t.tile_ptr, t.ntiles, t.nb, a.w, \
                                                                            a.w2, a.d, t.m, t.m2);                         \
  } while (0)
  if (t.nnz > 0) {
    if (elem_bytes == 4) { if (dual) AGD_LAUNCH_A(float, true); else AGD_LAUNCH_A(float, false); }
    else { if (dual) AGD_LAUNCH_A(double, true); else AGD_LAUNCH_A(double, false); }
  }
#undef AGD_LAUNCH_A
  const unsigned grid_r = grid_for(a.rows, sm_count);
  if (dual) csr_rows_kernel<true><<<grid_r, 256, 0, st>>>(t.m, t.m2, a.labels, a.rows, a.kind, a.sample_seed, a.sample_thresh, a.row_base, t.mult, a.gacc, a.d);
  else csr_rows_kernel<false><<<grid_r, 256, 0, st>>>(t.m, t.m2, a.labels, a.rows, a.kind, a.sample_seed, a.sample_thresh, a.row_base, t.mult, a.gacc, a.d);
  if (t.nnz > 0) {
    if (elem_bytes == 4) {
      e = cudaFuncSetAttribute(csr_tile_grad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
      if (e != cudaSuccess) return e;
      csr_tile_grad_kernel<float><<<(unsigned)grid_b, kThreads, smem_b, st>>>(t.pk, (const float *)t.tval, t.tile_ptr, t.ntiles, t.nb, t.mult, a.d, a.gacc);
    } else {
      e = cudaFuncSetAttribute(csr_tile_grad_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
      if (e != cudaSuccess) return e;
      csr_tile_grad_kernel<double><<<(unsigned)grid_b, kThreads, smem_b, st>>>(t.pk, (const double *)t.tval, t.tile_ptr, t.ntiles, t.nb, t.mult, a.d, a.gacc);
    }
  }
  return cudaGetLastError();
}

}  // namespace agd
