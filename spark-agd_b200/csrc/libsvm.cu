// libsvm.cu -- LIBSVM text -> CSR rows, the ingest format of the reference's callers
// (MLUtils.loadLibSVMFile of spark-mllib 1.3.0 produces the RDD[(Double, Vector)] that optimize() receives).
// Semantics restated from 1.3.0: one example per line, `label index1:value1 index2:value2 ...`, indices are one-based
// and ascending, stored zero-based; blank lines and lines starting with '#' are skipped; the feature count is the
// given numFeatures or, when <= 0, the largest index seen.  Host-only code (no GPU needed to parse).
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <limits.h>

#include "../../include/agd_b200.h"

namespace agd { void set_last_error(agd_handle *h, const char *msg); }  // agd_api.cu (internal, not part of the ABI)

extern "C" {

struct agd_libsvm {
  std::vector<int64_t> rowptr;
  std::vector<int32_t> idx;
  std::vector<double> val;
  std::vector<double> labels;
  int32_t d = 0;
  std::string err;
};

int agd_libsvm_read(const char *path, int32_t num_features, agd_libsvm **out) {
  if (!path || !out) return 1;
  agd_libsvm *L = new agd_libsvm();
  *out = L;
  FILE *f = fopen(path, "rb");
  if (!f) { L->err = std::string("cannot open ") + path + ": " + strerror(errno); return 1; }
  L->rowptr.push_back(0);
  int32_t max_index = 0;
  char *line = nullptr;
  size_t cap = 0;
  long long lineno = 0;
  ssize_t got;
  while ((got = getline(&line, &cap, f)) >= 0) {
    ++lineno;
    char *p = line;
    while (*p == ' ' || *p == '\t') ++p;
    char *end = p + strlen(p);
    while (end > p && (end[-1] == '\n' || end[-1] == '\r' || end[-1] == ' ' || end[-1] == '\t')) *--end = 0;
    if (*p == 0 || *p == '#') continue;
    char *q = nullptr;
    const double label = strtod(p, &q);
    if (q == p) { L->err = "line " + std::to_string(lineno) + ": cannot parse the label"; fclose(f); free(line); return 1; }
    p = q;
    int32_t prev = 0;
    for (;;) {
      while (*p == ' ' || *p == '\t') ++p;
      if (*p == 0) break;
      const long index = strtol(p, &q, 10);
      if (q == p || *q != ':') { L->err = "line " + std::to_string(lineno) + ": expected index:value"; fclose(f); free(line); return 1; }
      p = q + 1;
      const double v = strtod(p, &q);
      if (q == p) { L->err = "line " + std::to_string(lineno) + ": cannot parse a value"; fclose(f); free(line); return 1; }
      p = q;
      if (index > (long)INT32_MAX) {  // must be rejected BEFORE the narrowing cast below
        L->err = "line " + std::to_string(lineno) + ": feature index " + std::to_string(index) + " exceeds the int32 range";
        fclose(f); free(line); return 1;
      }
      if (index < 1 || index <= prev) {  // loadLibSVMFile requires one-based ascending indices
        L->err = "line " + std::to_string(lineno) + ": indices must be one-based and ascending";
        fclose(f); free(line); return 1;
      }
      prev = (int32_t)index;
      L->idx.push_back((int32_t)index - 1);
      L->val.push_back(v);
    }
    if (prev > max_index) max_index = prev;
    L->labels.push_back(label);
    L->rowptr.push_back((int64_t)L->idx.size());
  }
  free(line);
  fclose(f);
  L->d = num_features > 0 ? num_features : max_index;
  if (num_features > 0 && max_index > num_features) {
    L->err = "feature index " + std::to_string(max_index) + " exceeds numFeatures " + std::to_string(num_features);
    return 1;
  }
  return 0;
}

int64_t agd_libsvm_rows(const agd_libsvm *L) { return L ? (int64_t)L->labels.size() : 0; }
int32_t agd_libsvm_dim(const agd_libsvm *L) { return L ? L->d : 0; }
int64_t agd_libsvm_nnz(const agd_libsvm *L) { return L ? (int64_t)L->idx.size() : 0; }
const int64_t *agd_libsvm_rowptr(const agd_libsvm *L) { return L->rowptr.data(); }
const int32_t *agd_libsvm_indices(const agd_libsvm *L) { return L->idx.data(); }
const double *agd_libsvm_values(const agd_libsvm *L) { return L->val.data(); }
const double *agd_libsvm_labels(const agd_libsvm *L) { return L->labels.data(); }
const char *agd_libsvm_error(const agd_libsvm *L) { return L ? L->err.c_str() : "null"; }
void agd_libsvm_free(agd_libsvm *L) { delete L; }

// Parse + shard: rows are split contiguously over the handle's local GPUs (CSR storage `store_dtype`).
int agd_load_libsvm(agd_handle *h, const char *path, int32_t num_features, int32_t store_dtype) {
  if (!h) return 1;
  agd_libsvm *L = nullptr;
  int rc = agd_libsvm_read(path, num_features, &L);
  if (rc) {
    // the parser's message becomes the handle's agd_last_error (what MLUtils.loadLibSVMFile's caller sees as the exception text)
    agd::set_last_error(h, (std::string("agd_load_libsvm: ") + agd_libsvm_error(L)).c_str());
    agd_libsvm_free(L);
    return 1;
  }
  const int64_t n = agd_libsvm_rows(L);
  int nd = 0;
  while (agd_rows(h, nd) >= 0) ++nd;
  for (int i = 0; i < nd && rc == 0; ++i) {
    const int64_t lo = (int64_t)i * n / nd, hi = (int64_t)(i + 1) * n / nd;
    std::vector<int64_t> rp((size_t)(hi - lo) + 1);
    for (int64_t r = lo; r <= hi; ++r) rp[(size_t)(r - lo)] = L->rowptr[(size_t)r] - L->rowptr[(size_t)lo];
    const int64_t a = L->rowptr[(size_t)lo];
    rc = agd_load_csr(h, i, rp.data(), L->idx.data() + a, L->val.data() + a, AGD_F64, L->labels.data() + lo, hi - lo,
                      L->d, store_dtype);
  }
  agd_libsvm_free(L);
  return rc;
}

}  // extern "C"
