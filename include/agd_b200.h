/*
 * agd_b200.h -- C-ABI of the B200-native accelerated-gradient-descent hot path.
 *
 * This is the drop-in boundary a JVM binding (JNI) for staple/spark-agd would bind: plain
 * pointers and sizes, no C++/torch types.  Reference paths are relative to /root/reference:
 *   AGD.scala   = src/main/scala/org/apache/spark/mllib/optimization/AcceleratedGradientDescent.scala
 *   Suite.scala = src/test/scala/org/apache/spark/mllib/optimization/AcceleratedGradientDescentSuite.scala
 *
 * Model: one agd_handle per process owns one or more local B200s.  Each GPU pins one row-shard of
 * the (n x d) design matrix in HBM (the analogue of `dataRDD.cache()`, Suite.scala:51).  A "pass" is
 * one applySmooth (AGD.scala:192-208): fused row-block gradient kernel over the shard, one
 * all-reduce of the packed [grad(d) | loss | count] fp64 buffer, and the fused O(d) update kernel.
 * All entry points return 0 on success, nonzero on error (see agd_last_error).  There is no CPU
 * fallback anywhere: without a usable sm_100 GPU every compute entry point fails.
 *
 * Threading: agd_reserve / agd_load_* may be called concurrently for DIFFERENT local devices (Spark
 * task threads); everything else is single-caller, like the driver thread of AGD.scala:177.
 * Ownership: the caller owns every host buffer; the library owns device memory until agd_destroy.
 */
#ifndef AGD_B200_H
#define AGD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGD_B200_ABI_VERSION 2

typedef struct agd_handle agd_handle;

/* Closed enum of the Gradient plug-ins the reference can be given (AGD.scala:41,198):
 * LogisticGradient (binary), LeastSquaresGradient, HingeGradient of spark-mllib 1.3.0.
 * AGD_GRAD_LEAST_SQUARES_HALF is the Spark >= 1.4 definition (loss diff^2/2, gradient diff*x). */
enum { AGD_GRAD_LOGISTIC = 0, AGD_GRAD_LEAST_SQUARES = 1, AGD_GRAD_HINGE = 2, AGD_GRAD_LEAST_SQUARES_HALF = 3 };
/* Closed enum of the Updater plug-ins (AGD.scala:41,215): SimpleUpdater, SquaredL2Updater, L1Updater. */
enum { AGD_UPD_SIMPLE = 0, AGD_UPD_SQUARED_L2 = 1, AGD_UPD_L1 = 2 };
/* Element types: host source type and HBM storage type of the design matrix. */
enum { AGD_F64 = 0, AGD_F32 = 1, AGD_BF16 = 2 };
/* agd_params.flags */
enum {
  AGD_FLAG_MEMOIZE_FX = 1, /* reuse (f_x, g_x) of AGD.scala:269 for the history pass at :304 when x is
                              unchanged (bit-identical result, 3 -> 2 passes per iteration).  Where the shard's kernel has a
                              two-gradient form, applySmooth(x) then shares its sweep with applySmooth(y) of the NEXT
                              iteration (y guessed from "accepted, no restart"; a wrong guess is discarded, a restart reuses
                              (f_x, g_x) because then y = x): an accepted iteration reads X once. */
  AGD_FLAG_NO_FUSE = 2,    /* by default the history evaluation applySmooth(x) of AGD.scala:304 rides along with
                              applySmooth(y) of the NEXT iteration (AGD.scala:250) in one sweep over the shards: the same
                              evaluations and the same results bit for bit (dense shards), one read of X fewer per iteration.
                              This flag runs every evaluation as a sweep of its own. */
};

/* The constructor arguments + eight hyper-parameters of AGD.scala:41-51 (defaults: agd_default_params). */
typedef struct {
  double convergence_tol; /* AGD.scala:44  default 1e-4 */
  int32_t num_iterations; /* AGD.scala:45  default 100  */
  double reg_param;       /* AGD.scala:46  default 0.0  */
  double L0;              /* AGD.scala:47  default 1.0  */
  double Lexact;          /* AGD.scala:48  default +inf */
  double beta;            /* AGD.scala:49  default 0.5  */
  double alpha;           /* AGD.scala:50  default 0.9  */
  int32_t may_restart;    /* AGD.scala:51  default 1    */
  int32_t gradient;       /* AGD_GRAD_*  (the `gradient` delegate, AGD.scala:41) */
  int32_t updater;        /* AGD_UPD_*   (the `updater` delegate,  AGD.scala:41) */
  int32_t flags;          /* AGD_FLAG_*; 0 reproduces the reference's pass structure exactly */
} agd_params;

/* What `run` learned; the reference only logs (AGD.scala:310,334). */
typedef struct {
  int32_t iterations;     /* = length of the loss history */
  int32_t passes;         /* applySmooth evaluations executed */
  int32_t backtracks;     /* times AGD.scala:292 raised L */
  int32_t restarts;       /* times AGD.scala:327-331 fired */
  int32_t converged;      /* left through AGD.scala:319 or :323 */
  int32_t stopped_nan;    /* left through AGD.scala:309-312 */
  int32_t nonterminating; /* L became NaN: the reference would spin forever in :246-293; we stop */
  int32_t collective_kind; /* 0 = none / NCCL all-reduce, 1 = one-shot exchange over NVLink peer memory */
  double final_L;
  double final_theta;
  double seconds_total;   /* host wall time inside agd_run */
  double k1_ms_total;     /* CUDA-event time of the gradient kernel (device 0), all passes */
  int64_t k1_launches;    /* launches of the gradient kernel per device */
  int64_t gpu_launches;   /* this library's own kernel launches issued by the call, per device */
  double allreduce_ms_total; /* CUDA-event time of the all-reduce (0 when world == 1) */
  double device_ms_total;    /* CUDA events on device 0's stream around the whole call */
  int64_t collective_calls;  /* all-reduces enqueued per device */
  int32_t wasted_passes;     /* speculative applySmooth(x) passes discarded because ||x-y||^2 == 0 (AGD.scala:265) */
  int32_t fused_passes;      /* applySmooth evaluations that shared a sweep over X with another one (sweeps = passes - fused_passes) */
} agd_stats;

/* ---- lifecycle ---- */
int agd_abi_version(void);
/* sizeof(agd_params) / sizeof(agd_stats) as compiled, so a foreign binding can verify its struct layout */
int agd_sizeof_params(void);
int agd_sizeof_stats(void);
void agd_default_params(agd_params *p);
/* Opens n_dev local GPUs (device ordinals in device_ids).  Fails when a device is not sm_100. */
int agd_create(const int32_t *device_ids, int32_t n_dev, agd_handle **out);
int agd_destroy(agd_handle *h);
/* Message of the last failure on this handle (h may be NULL: last agd_create failure). */
const char *agd_last_error(const agd_handle *h);

/* ---- the collective (replaces treeAggregate + broadcast, AGD.scala:193,196-204) ----
 * One rank per GPU.  With a single process owning all GPUs nothing needs to be called (the local GPUs
 * are a complete world: direct peer pointers, NCCL only as a fallback); calling agd_comm_init on such a
 * handle replaces that default, e.g. two processes with two GPUs each = world 4.  With one process per GPU (torchrun / one executor per GPU): rank 0
 * calls agd_comm_unique_id, ships the 128 bytes to every process, and every process calls
 * agd_comm_init(h, id, world_ranks, first_rank) where its local GPUs take ranks
 * first_rank .. first_rank + n_dev - 1. */
int agd_comm_unique_id(void *out128);
int agd_comm_init(agd_handle *h, const void *id128, int32_t world_ranks, int32_t first_rank);
/* The same world WITHOUT NCCL.  The per-pass exchange never uses a library (P2P stores + epoch flags into peer HBM,
 * csrc/xchg.cu); only its setup needs every rank to learn every other rank's CUDA IPC handles, and the host language can
 * ship those itself (a Spark driver collecting one small blob per executor; torch.distributed / gloo in the tests):
 *   1. every process:          agd_comm_init_ipc(h, world_ranks, first_rank)
 *   2. after loading its shards (the feature dimension sizes the buffers):
 *                              agd_xchg_export(h, blob, capacity, &n)   -> n = local GPUs * AGD_XCHG_HANDLE_BYTES
 *   3. host: concatenate every process's blob in rank order, hand the whole to every process:
 *                              agd_xchg_import(h, all, world_ranks * AGD_XCHG_HANDLE_BYTES)
 * Works between processes on different GPUs (NVLink / PCIe P2P) and between processes sharing ONE GPU (CUDA IPC on the
 * same device), which is how the 1-GPU test box exercises this path.  There is no fallback in such a world: if a pair
 * of ranks cannot map each other, agd_xchg_import fails.  Steps 2-3 are repeated after agd_clear + a load with another d. */
#define AGD_XCHG_HANDLE_BYTES 192
int agd_comm_init_ipc(agd_handle *h, int32_t world_ranks, int32_t first_rank);
int agd_xchg_export(agd_handle *h, void *out, int64_t capacity_bytes, int64_t *bytes_written);
int agd_xchg_import(agd_handle *h, const void *all_ranks, int64_t bytes);

/* ---- shard loading (replaces RDD[(Double, Vector)] partitions cached on executors, AGD.scala:178) ----
 * agd_reserve fixes the shard geometry of local device `dev` and allocates HBM for `rows_capacity`
 * rows stored as `store_dtype` (AGD_F64, AGD_F32 or AGD_BF16; values are rounded to nearest-even).  agd_load_dense APPENDS `rows` rows (row-major,
 * leading dimension ld elements, element type src_dtype AGD_F64|AGD_F32) and their labels.  If the
 * device was not reserved, the first load reserves exactly `rows`. */
int agd_reserve(agd_handle *h, int32_t dev, int64_t rows_capacity, int32_t d, int32_t store_dtype);
int agd_load_dense(agd_handle *h, int32_t dev, const void *X, int32_t src_dtype, const double *labels,
                   int64_t rows, int32_t d, int64_t ld, int32_t store_dtype);
/* SparseVector rows as CSR (values src_dtype AGD_F64|AGD_F32, stored as store_dtype AGD_F64|AGD_F32).
 * APPENDS `rows` rows: rowptr has rows+1 entries starting at 0 and is rebased onto the resident shard. */
int agd_load_csr(agd_handle *h, int32_t dev, const int64_t *rowptr, const int32_t *idx, const void *val,
                 int32_t src_dtype, const double *labels, int64_t rows, int32_t d, int32_t store_dtype);
/* LIBSVM text ingest (MLUtils.loadLibSVMFile of spark-mllib 1.3.0, the usual producer of the RDD handed to optimize):
 * `label index:value ...` per line, one-based ascending indices, '#' comment lines and blank lines skipped,
 * num_features <= 0 infers the dimension.  agd_libsvm_read parses on the host (no GPU needed) into an opaque
 * object with accessors; agd_load_libsvm parses and appends the rows as CSR, split over the local GPUs. */
typedef struct agd_libsvm agd_libsvm;
int agd_libsvm_read(const char *path, int32_t num_features, agd_libsvm **out);
int64_t agd_libsvm_rows(const agd_libsvm *L);
int32_t agd_libsvm_dim(const agd_libsvm *L);
int64_t agd_libsvm_nnz(const agd_libsvm *L);
const int64_t *agd_libsvm_rowptr(const agd_libsvm *L);
const int32_t *agd_libsvm_indices(const agd_libsvm *L);
const double *agd_libsvm_values(const agd_libsvm *L);
const double *agd_libsvm_labels(const agd_libsvm *L);
const char *agd_libsvm_error(const agd_libsvm *L);
void agd_libsvm_free(agd_libsvm *L);
int agd_load_libsvm(agd_handle *h, const char *path, int32_t num_features, int32_t store_dtype);
/* Drops every shard (all local devices). */
int agd_clear(agd_handle *h);
/* Rows currently resident on local device `dev`; feature count (0 when empty). */
int64_t agd_rows(const agd_handle *h, int32_t dev);
int32_t agd_dim(const agd_handle *h);

/* ---- measurement harness (not in the reference): synthetic workload generated in place ----
 * Every GPU rank r of the world fills its shard with global rows [r*total_rows/W, (r+1)*total_rows/W)
 * of the counter-based synthetic design matrix (spec: spark-agd_b200/csrc/synth.cu) and labels for
 * `gradient`.  agd_get_rows downloads rows (as the storage dtype) and labels for checking. */
int agd_generate(agd_handle *h, int64_t total_rows, int32_t d, int32_t store_dtype, uint64_t seed,
                 int32_t gradient);
int agd_get_rows(agd_handle *h, int32_t dev, int64_t row0, int64_t rows, void *X_out, double *labels_out);
int agd_synth_wtrue(agd_handle *h, uint64_t seed, int32_t d, double *w_out);
/* CSR flavour of the synthetic workload: exactly nnz_per_row stored entries per row, strictly increasing
 * column ids (stratified), same label rules.  agd_get_csr_rows downloads a row range (rowptr rebased to 0). */
int agd_generate_csr(agd_handle *h, int64_t total_rows, int32_t d, int32_t nnz_per_row, int32_t store_dtype,
                     uint64_t seed, int32_t gradient);
int agd_get_csr_rows(agd_handle *h, int32_t dev, int64_t row0, int64_t rows, int64_t *rowptr_out, int32_t *idx_out,
                     void *val_out, int64_t nnz_capacity, double *labels_out);

/* ---- plug-in granularity entry points (host buffers in and out) ----
 * agd_smooth = applySmooth (AGD.scala:192-208): loss/count and grad/count over ALL shards of the
 * world; w, grad are d doubles on the host.  Every rank must call it. */
int agd_smooth(agd_handle *h, int32_t gradient, const double *w, double *loss, double *grad, int64_t *count);
/* agd_smooth at w plus the loss (no gradient) at a second point w2, both from ONE sweep over the shards -- the fused form of
 * applySmooth(y) (AGD.scala:250) and the history evaluation applySmooth(x) (:304) that agd_run uses.  On dense shards every
 * output equals, bit for bit, what two agd_smooth calls return.  Fails on shards whose kernel has no two-point form
 * (tcgen05 bf16 path, d below one 16-row tile); agd_run then simply does not fuse. */
int agd_smooth_pair(agd_handle *h, int32_t gradient, const double *w, const double *w2, double *loss, double *grad,
                    int64_t *count, double *loss2);
/* Two complete applySmooth evaluations (loss and gradient at w AND at w2) from ONE sweep over the shards -- what agd_run's
 * memoised pass structure uses to evaluate applySmooth(x) of the backtracking test (AGD.scala:269) together with
 * applySmooth(y) of the next iteration (:250).  Bit for bit what two agd_smooth calls return.  Dense fp32 / fp64 shards with
 * at most 512 16-byte vectors per row (d <= 2048 fp32, <= 1024 fp64) and bf16 shards on the tcgen05 kernel; fails elsewhere. */
int agd_smooth_two(agd_handle *h, int32_t gradient, const double *w, const double *w2, double *loss, double *grad,
                   int64_t *count, double *loss2, double *grad2);
/* agd_prox = applyProjector (AGD.scala:214-222): Updater.compute(w, g, step, iter = 1, reg). */
int agd_prox(agd_handle *h, int32_t updater, const double *w, const double *g, double step, double reg,
             int32_t d, double *w_out, double *reg_val);

/* ---- the whole loop, natively: AcceleratedGradientDescent.run (AGD.scala:177-338) ----
 * w0, w_out: d doubles.  loss_hist: capacity >= max(num_iterations, 1); *n_hist receives its length.
 * Every rank must call it with identical arguments; every rank receives identical results. */
int agd_run(agd_handle *h, const agd_params *p, const double *w0, double *w_out, double *loss_hist,
            int32_t *n_hist, agd_stats *stats);

/* GradientDescent.runMiniBatchSGD of spark-mllib 1.3.0 with miniBatchFraction = 1.0 (the comparator
 * the reference's tests run beside AGD, Suite.scala:78,118,225), on the same kernels. */
int agd_gd_run(agd_handle *h, int32_t gradient, int32_t updater, double step_size, int32_t num_iterations,
               double reg_param, const double *w0, double *w_out, double *loss_hist, int32_t *n_hist,
               agd_stats *stats);

/* The mini-batch form: iteration i uses the rows kept by `data.sample(false, miniBatchFraction, 42 + i)`, realised as a
 * counter-based Bernoulli mask (Philox keyed by 42 + i and the global row index; Spark's own sampler is seeded per
 * partition and is not reproducible across partitionings either).  fraction >= 1 is the full batch. */
int agd_gd_run_minibatch(agd_handle *h, int32_t gradient, int32_t updater, double step_size, int32_t num_iterations,
                         double reg_param, double mini_batch_fraction, const double *w0, double *w_out,
                         double *loss_hist, int32_t *n_hist, agd_stats *stats);

/* Name of the gradient kernel the shard on local device `dev` dispatches to (for reports), "" when empty. */
const char *agd_kernel_name(const agd_handle *h, int32_t dev);

/* Options: "k1_variant" = auto|ring|generic|tc, "collective" = auto|nccl|p2p, ring tuning knobs. */
int agd_set_option(agd_handle *h, const char *key, const char *value);

#ifdef __cplusplus
}
#endif
#endif /* AGD_B200_H */
