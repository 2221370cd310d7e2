/* agd_oracle.h -- CPU ORACLE interface (test infrastructure, not product code).
 * See agd_oracle.c for the reference file:line each function restates. */
#ifndef AGD_ORACLE_H
#define AGD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_DENSE_F64 = 0, ORACLE_DENSE_F32 = 1, ORACLE_CSR = 2 };
enum { ORACLE_GRAD_LOGISTIC = 0, ORACLE_GRAD_LEAST_SQUARES = 1, ORACLE_GRAD_HINGE = 2,
       ORACLE_GRAD_LEAST_SQUARES_HALF = 3 };
enum { ORACLE_UPD_SIMPLE = 0, ORACLE_UPD_SQUARED_L2 = 1, ORACLE_UPD_L1 = 2 };

/* RDD[(Double, Vector)] held on the host: labels + rows (AGD.scala:178). */
typedef struct {
  int64_t n;
  int32_t d;
  int32_t storage;            /* ORACLE_DENSE_F64 | ORACLE_DENSE_F32 | ORACLE_CSR */
  const double *Xd;           /* dense f64, row-major, leading dimension ld */
  const float *Xf;            /* dense f32 (upcast to f64 on use) */
  int64_t ld;
  const int64_t *rowptr;      /* CSR (SparseVector rows): n+1 */
  const int32_t *csr_idx;
  const double *csr_val;      /* f64 values, or NULL when csr_val_f32 is set */
  const float *csr_val_f32;
  const double *labels;       /* n */
  uint64_t sample_seed;       /* mini-batch mask of the current pass; sample_thresh == 0 keeps every row */
  uint64_t sample_thresh;
} oracle_data;

/* The eight hyper-parameters of AGD.scala:44-51 + the treeAggregate shape. */
typedef struct {
  double convergence_tol;
  int32_t num_iterations;
  double reg_param;
  double L0;
  double Lexact;
  double beta;
  double alpha;
  int32_t may_restart;
  int32_t partitions;         /* RDD partitions (Suite.scala:51 uses 2) */
  int32_t threads;            /* executor threads folding partitions concurrently */
} oracle_params;

typedef struct {
  int32_t iterations;         /* = len(lossHistory) */
  int32_t passes;             /* applySmooth evaluations */
  int32_t backtracks;
  int32_t restarts;
  int32_t converged;
  int32_t stopped_nan;
  int32_t nonterminating;     /* the reference would loop forever (NaN L) */
  double final_L;
  double final_theta;
} oracle_stats;

typedef struct { uint64_t seed; int have_next; double next_gaussian; } oracle_jrandom;

void oracle_default_params(oracle_params *p);
int oracle_smooth(const oracle_data *D, int grad_kind, const double *w, int partitions, int threads,
                  double *loss_out, double *grad_out, int64_t *count_out);
int oracle_prox(int upd_kind, const double *w, const double *g, double step, double reg, int32_t d,
                double *w_out, double *reg_val);
int oracle_agd_run(const oracle_data *D, int grad_kind, int upd_kind, const oracle_params *p,
                   const double *w0, double *w_out, double *loss_hist, int32_t *n_hist, oracle_stats *st);
int oracle_gd_run(const oracle_data *D, int grad_kind, int upd_kind, double step_size, int num_iterations,
                  double reg_param, int partitions, int threads, const double *w0, double *w_out,
                  double *loss_hist, int32_t *n_hist);

void oracle_jrandom_seed(oracle_jrandom *r, int64_t seed);
double oracle_jrandom_next_double(oracle_jrandom *r);
double oracle_jrandom_next_gaussian(oracle_jrandom *r);
void oracle_jrandom_fill_double(int64_t seed, int64_t n, double *out);
int oracle_jrandom_continue_fill_double(oracle_jrandom *r, int64_t n, double *out);
void oracle_generate_gd_input(double offset, double scale, int32_t n_points, int32_t seed, double *x1, double *y);
int oracle_max_threads(void);
/* CPU-timing hygiene for bench.py's reference arm (thread pinning, first-touch placement) */
int oracle_bind_threads(int threads);
void oracle_unbind_threads(void);
void oracle_first_touch(void *base, int64_t rows, int64_t row_bytes, int partitions, int threads);
void oracle_set_threads(int n);

/* synthetic-workload twin of the product's on-device generator (oracle/synth_oracle.c) */
void oracle_synth_dense_f32(uint64_t seed, int64_t row0, int64_t rows, int32_t d, float *X);
void oracle_synth_dense_f32_placed(uint64_t seed, int64_t row0, int64_t rows, int32_t d, float *X, int partitions,
                                   int threads);
void oracle_synth_wtrue(uint64_t seed, int32_t d, double *w);
void oracle_synth_labels(uint64_t seed, int kind, int64_t row0, int64_t rows, int32_t d, const float *X,
                         const double *w_true, double *labels);

int oracle_row_selected(uint64_t seed, uint64_t thresh, int64_t grow);
int oracle_gd_run_minibatch(oracle_data *D, int grad_kind, int upd_kind, double step_size, int num_iterations,
                            double reg_param, double fraction, int partitions, int threads, const double *w0,
                            double *w_out, double *loss_hist, int32_t *n_hist);
void oracle_synth_csr_f32(uint64_t seed, int64_t row0, int64_t rows, int32_t d, int32_t k, int64_t *rowptr,
                          int32_t *idx, float *val);

#ifdef __cplusplus
}
#endif
#endif
