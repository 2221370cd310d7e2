// agd_common.cuh -- shared declarations of the sm_100a hot-path kernels (internal, not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/agd_b200.h"

namespace agd {

// ---------------------------------------------------------------- K1: fused row-block gradient
// Replaces the seqOp fold of AGD.scala:197-200 + Gradient.compute [mllib-1.3.0] over one shard.
struct K1Args {
  const void *X;          // shard, row-major, ld == d, element type float or double
  const double *labels;   // rows (+ padding)
  const double *w;        // d doubles (device)
  const double *w2;       // optional second point of a fused sweep (AGD.scala:304 riding along with :250): loss only ...
  int32_t dual_full;      // ... unless dual_full: loss AND gradient at w2 (second slab block of d + 4 doubles; ring kernel, fp32/fp64)
  double *slabs;          // [grid][d + 4]: per-block column sums of loss' * x, the loss sum, the row count at w; loss sum, count at w2
  int64_t rows;           // rows in the shard
  int32_t d;
  int32_t kind;           // AGD_GRAD_*
  int32_t stages;         // smem ring depth
  int32_t slab_stride;    // d + 4
  unsigned long long sample_seed, sample_thresh;  // Bernoulli row mask (thresh 0 = every row), see row_selected()
  long long row_base;     // global index of the shard's first row
  int32_t tune_rows;      // 0 = default; rows per tile of the headline ring shape (4|8)
  int32_t tune_ctas;      // 0 = default; resident CTAs per SM (1|2|3)
  int32_t tune_full;      // 0 = default; 1 = keep the column predicates even when every thread owns whole vectors
  int32_t tc_margins_f64; // tcgen05 kernel: 1 = fp64-exact margins on the CUDA cores (option tc_margins=f64), 0 = fp32 (default)
};

// launch helpers (k1_dense.cu); return the number of blocks that wrote a slab
int k1_ring_supported(int32_t d, int elem_bytes);
int k1_ring_dual_supported(int32_t d, int elem_bytes);
cudaError_t k1_ring_launch(const K1Args &a, int elem_bytes, int sm_count, int *blocks_out, cudaStream_t st);
int k1_ring_dual_full_supported(int32_t d, int elem_bytes);
cudaError_t k1_generic_launch(const K1Args &a, int elem_bytes, int sm_count, int max_blocks, int *blocks_out,
                              cudaStream_t st);
int k1_max_blocks(int sm_count);
// bf16 shards: margins on CUDA cores, X^T r on tcgen05 (k1_tc.cu); d % 128 == 0, d <= 4096
int k1_tc_supported(int32_t d, int elem_bytes);
cudaError_t k1_tc_launch(const K1Args &a, int sm_count, int *blocks_out, cudaStream_t st);
// ---------------------------------------------------------------- K2': one-shot all-reduce over NVLink peer memory
// Every rank owns an exchange buffer xbuf[2][W][n] (+ flags[2][W]) that all peers can store into (P2P / CUDA IPC).
// publish: rank r stores its n = d+4 partial sums into slot r of EVERY rank's buffer, fences, then raises the flag
// (epoch) -- fused into the tail of k1_reduce_kernel, or standalone for the CSR path.  gather: each rank waits for the
// W flags of the epoch and adds the W slots in rank order, so every rank gets the same bits (replaces combOp +
// treeAggregate + broadcast, AGD.scala:193-204, without a library call in the loop).  Buffers alternate by epoch parity.
constexpr int kMaxRanks = 16;
struct XchgPeers {
  double *slot[kMaxRanks];               // base of rank p's xbuf as mapped into THIS device
  unsigned long long *flag[kMaxRanks];   // base of rank p's flags
};
struct XchgPub {
  XchgPeers peers;
  int world, my_rank, buf, n;
  int slot_stride;          // doubles between two ranks' slots (the buffers' capacity per slot, NOT this sweep's n: sweeps of
                            // different payloads -- d + 4 or 2 (d + 4) -- must not move the slots of the other parity buffer)
  unsigned long long epoch;
  unsigned int *ticket;
};
// The gather half of the exchange, for kernels that consume the all-reduced sums right away (K3): instead of a launch of its own,
// the consumer waits for the W flags, adds the W slots in rank order where it needs a value, and writes every one of the n sums
// to `acc` for later readers.  world == 0: nothing pending, read `acc` as usual.
struct XchgGather {
  const double *xbuf = nullptr;               // this device's exchange buffer [2][W][slot_stride] (one-shot) or its result area (rs)
  const unsigned long long *flags = nullptr;  // [2][W] of the flag set to wait on
  int world = 0, buf = 0, n = 0, slot_stride = 0;
  int rs = 0;                                 // 1: reduce-scatter form -- the finished sums sit at xbuf[buf * slot_stride + j]
  unsigned long long epoch = 0;
};
// ---- large payloads (n >= kXchgRsMin doubles, e.g. d = 10^6): reduce-scatter + all-gather over the same peer memory.
// Rank r stores slice p of its partial sums into rank p's `rs` area (slot r), rank p adds the W slots of ITS slice in rank order
// and stores the finished slice into every rank's `res` area; 2 n / W doubles leave each rank per sweep instead of n W.
// Layout of one device's exchange allocation (doubles): [one-shot: 2 W S][rs: 2 W L][res: 2 S], S = 2 (d + 4), L = ceil(S / W);
// flags (u64): [one-shot 2 W][rs arrived 2 W][res arrived 2 W].
constexpr int kXchgRsMin = 32768;
struct XchgRs {
  XchgPeers peers;
  int world, my_rank, buf, n;
  int slot_stride;              // S
  unsigned long long epoch;
  unsigned int *ticket;
};
cudaError_t xchg_rs_publish_launch(const double *acc, const XchgRs &x, cudaStream_t st);
cudaError_t xchg_rs_reduce_bcast_launch(const double *xbuf_local, const unsigned long long *flags_local, const XchgRs &x, cudaStream_t st);
// waits for the W finished slices and copies them to acc_out (stand-alone form of the rs gather)
cudaError_t xchg_rs_gather_launch(const double *xbuf_local, const unsigned long long *flags_local, int world, int buf, int n,
                                  int slot_stride, unsigned long long epoch, double *acc_out, cudaStream_t st);
inline size_t xchg_rs_slice(int S, int W) { return ((size_t)S + W - 1) / W; }
inline size_t xchg_off_rs(int S, int W) { return 2 * (size_t)W * S; }
inline size_t xchg_off_res(int S, int W) { return xchg_off_rs(S, W) + 2 * (size_t)W * xchg_rs_slice(S, W); }
inline size_t xchg_total_doubles(int S, int W) { return xchg_off_res(S, W) + 2 * (size_t)S; }
cudaError_t xchg_publish_launch(const double *acc, const XchgPub &pub, cudaStream_t st);
cudaError_t xchg_gather_launch(const double *xbuf_local, const unsigned long long *flags_local, int world, int buf, int n,
                               int slot_stride, unsigned long long epoch, double *acc_out, cudaStream_t st);

// out[c] = sum_b slabs[b][c] for c < n, n = d + 4 or 2 (d + 4) (gradient sums, loss sum, row count, loss sum and count at w2; fixed order =>
// deterministic);
// with pub != nullptr the sums are also stored into every peer's exchange slot and the epoch flag is raised
cudaError_t k1_reduce_launch(const double *slabs, int blocks, int32_t n, double *out, const XchgPub *pub, cudaStream_t st);

// CSR variant (k1_csr.cu)
struct K1CsrArgs {
  const int64_t *rowptr;
  const int32_t *idx;
  const void *val;        // float or double
  const double *labels;
  const double *w;
  const double *w2;       // optional second point (loss only), as in K1Args
  double *gacc;           // d + 4 doubles, zeroed by the launch: gradient sum, loss sum, count; loss sum, count at w2
  int64_t rows;
  int32_t d;
  int32_t kind;
  unsigned long long sample_seed, sample_thresh;
  long long row_base;
  int32_t tune;           // option ring_rows: 1 = the simple (unpipelined) loop
};
cudaError_t k1_csr_launch(const K1CsrArgs &a, int elem_bytes, int sm_count, cudaStream_t st);


// ---------------------------------------------------------------- K3: fused O(d) driver-side vector work
// Replaces the breeze d-vector ops of AGD.scala:249,255,263-264,273,278,315-316,327, the
// normalisation of :207 and Updater.compute [mllib-1.3.0] behind applyProjector (:214-222).
enum { K3_NS = 8 };  // scalars produced per call
struct K3StepArgs {
  const double *acc;      // packed [grad_sum(d) | loss_sum | count] after the all-reduce
  const double *x_old, *z_old, *y;
  double *g_y, *z, *x;
  double *partials;       // [blocks][K3_NS]
  unsigned int *ticket;
  double *scalars;        // [K3_NS]: S(x-y)^2, (x-y).g_y, S x^2, S (x-x_old)^2, g_y.(x-x_old), S|x|, loss_sum, count
  double theta, one_minus_theta, step, reg;
  int32_t d, updater;
  double *y_spec;         // optional: y_spec = x * spec_ca + z * spec_cb, the guessed y of the next iteration (speculative sweep)
  double spec_ca, spec_cb;
  unsigned long long *seq_out = nullptr;   // optional: launch sequence number stored behind the scalars (host polls it)
  unsigned long long seq = 0;
  XchgGather xg;          // pending exchange whose sums this kernel gathers itself (then `acc` is written, not only read)
  double *acc_w = nullptr;
  double *hist_out = nullptr;   // optional (mapped host memory): {loss sum, count} at the second point of the sweep just consumed (AGD.scala:304)
};
cudaError_t k3_step_launch(const K3StepArgs &a, cudaStream_t st);
struct K3GxArgs {
  const double *acc;
  const double *x, *y, *g_y;
  double *g_x;
  double *partials;
  unsigned int *ticket;
  double *scalars;        // [0] = (x-y).(g_x-g_y), [6] = loss_sum, [7] = count
  int32_t d;
  unsigned long long *seq_out = nullptr;
  unsigned long long seq = 0;
  XchgGather xg;          // as in K3StepArgs
  double *acc_w = nullptr;
};
cudaError_t k3_gx_launch(const K3GxArgs &a, cudaStream_t st);
// out = a*ca + b*cb (separate roundings, as breeze does at AGD.scala:249)
cudaError_t k3_combine_launch(double *out, const double *a, double ca, const double *b, double cb, int32_t d,
                              cudaStream_t st);
// x_old = x ; z_old = z ; y = x*ca + z*cb  (AGD.scala:241 + :249 in one launch)
cudaError_t k3_begin_launch(double *x_old, double *z_old, double *y, const double *x, const double *z, double ca,
                            double cb, int32_t d, cudaStream_t st);
// dst0 = src0 ; dst1 = src1 (either pair may be null)
cudaError_t k3_copy2_launch(double *dst0, const double *src0, double *dst1, const double *src1, int32_t d,
                            cudaStream_t st);
// plain prox for agd_prox / the GD comparator: w_out = Updater.compute(w, g[/count], step, reg); scalars[2]=S w'^2, [5]=S|w'|
struct K3ProxArgs {
  const double *w, *g;
  double *w_out;
  double *partials;
  unsigned int *ticket;
  double *scalars;
  double step, reg;
  const double *acc_tail;   // optional {loss_sum, count} on the device: g is divided by count first
                            // (count == 0 leaves w unchanged); scalars[6..7] receive the pair
  int32_t d, updater;
};
cudaError_t k3_prox_launch(const K3ProxArgs &a, cudaStream_t st);
int k3_blocks(int32_t d);

// ---------------------------------------------------------------- K0: synthetic workload (harness)
cudaError_t synth_dense_launch(void *X, int elem_bytes, uint64_t seed, int64_t row0, int64_t rows, int32_t d, int32_t ld,
                               cudaStream_t st);
cudaError_t synth_wtrue_launch(double *w, uint64_t seed, int32_t d, cudaStream_t st);
cudaError_t synth_labels_launch(const void *X, int elem_bytes, const double *w_true, double *labels, uint64_t seed,
                                int kind, int64_t row0, int64_t rows, int32_t d, int32_t ld, cudaStream_t st);

cudaError_t synth_csr_launch(int64_t *rowptr, int32_t *idx, void *val, int elem_bytes, const double *w_true,
                             double *labels, uint64_t seed, int kind, int64_t row0, int64_t rows, int32_t d, int32_t k,
                             cudaStream_t st);

// dst[i] = src[i] + shift (rebasing an appended partition's rowptr onto the resident CSR shard)
cudaError_t csr_shift_rowptr_launch(int64_t *dst, const int64_t *src, int64_t n, int64_t shift, cudaStream_t st);

// rowptr (as the caller passed it: rows + 1 entries from 0) and idx of an appended CSR partition, checked on the device:
// *flag = 0 ok, 1 = rowptr not monotone / not ending at nnz, 2 = a column id outside [0, d)
cudaError_t csr_validate_launch(const int64_t *rowptr_host_order, int64_t rows, const int32_t *idx, int64_t nnz, int32_t d,
                                int *flag, cudaStream_t st);

// records `msg` as the handle's agd_last_error (for translation units other than agd_api.cu)
void set_last_error(agd_handle *h, const char *msg);

// ---------------------------------------------------------------- load path
// dst (store dtype, rows x dst_ld, columns >= d zero) <- src (src dtype, leading dimension ld), rows x d
cudaError_t convert_rows_launch(void *dst, int dst_bytes, const void *src, int src_bytes, int64_t rows, int32_t d,
                                int64_t ld, int32_t dst_ld, cudaStream_t st);

}  // namespace agd
