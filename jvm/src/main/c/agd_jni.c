/*
 * agd_jni.c -- JNI shim over include/agd_b200.h for the Scala facade in
 * jvm/src/main/scala/org/apache/spark/mllib/optimization/NativeAGD.scala.
 *
 * Built where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       jvm/src/main/c/agd_jni.c -Lspark-agd_b200 -lagd_b200 -o libagd_jni.so
 * This image has no JDK.  tests/test_jvm_binding.py still puts the file through the C compiler against
 * tests/stubs/jni.h (the handful of JNI declarations it uses) with -fsyntax-only -Wall -Werror, and checks the
 * exported Java_..._NativeAGD_00024_* names and arities against the @native declarations in NativeAGD.scala.
 * Rules: pin arrays only for the duration of one C-ABI call, never retain JVM pointers, check every array length
 * against what the C side will read or write, turn every nonzero return code into a RuntimeException carrying
 * agd_last_error.
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "agd_b200.h"

#define H(ptr) ((agd_handle *)(intptr_t)(ptr))
#define JNI_FN(name) Java_org_apache_spark_mllib_optimization_NativeAGD_00024_##name

static void throw_msg(JNIEnv *env, const char *cls, const char *msg) {
  jclass ex = (*env)->FindClass(env, cls);
  if (ex) (*env)->ThrowNew(env, ex, msg);
}
static void fail(JNIEnv *env, agd_handle *h) { throw_msg(env, "java/lang/RuntimeException", agd_last_error(h)); }
static int bad_length(JNIEnv *env, const char *what, jlong have, jlong need) {
  if (have >= need) return 0;
  char buf[160];
  snprintf(buf, sizeof buf, "%s has %lld elements, the native call needs %lld", what, (long long)have, (long long)need);
  throw_msg(env, "java/lang/IllegalArgumentException", buf);
  return 1;
}

JNIEXPORT jlong JNICALL JNI_FN(create)(JNIEnv *env, jobject self, jintArray devices) {
  (void)self;
  jsize n = (*env)->GetArrayLength(env, devices);
  jint *ids = (*env)->GetIntArrayElements(env, devices, NULL);
  if (!ids) return 0;
  agd_handle *h = NULL;
  int rc = agd_create((const int32_t *)ids, (int32_t)n, &h);
  (*env)->ReleaseIntArrayElements(env, devices, ids, JNI_ABORT);
  if (rc) { fail(env, NULL); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL JNI_FN(destroy)(JNIEnv *env, jobject self, jlong h) {
  (void)env; (void)self;
  agd_destroy(H(h));
}

JNIEXPORT void JNICALL JNI_FN(clear)(JNIEnv *env, jobject self, jlong h) {
  (void)self;
  if (agd_clear(H(h))) fail(env, H(h));
}

/* One chunk of one RDD partition, packed row-major by the task thread that owns the partition (labels: rows doubles,
 * x: rows*d doubles); storeDtype = AGD_F64 | AGD_F32 | AGD_BF16.  Thread-safe per device (include/agd_b200.h). */
JNIEXPORT void JNICALL JNI_FN(loadDense)(JNIEnv *env, jobject self, jlong h, jint dev, jdoubleArray x,
                                         jdoubleArray labels, jlong rows, jint d, jint storeDtype) {
  (void)self;
  if (rows < 0 || d <= 0) { throw_msg(env, "java/lang/IllegalArgumentException", "bad partition geometry"); return; }
  if (bad_length(env, "x", (*env)->GetArrayLength(env, x), rows * (jlong)d)) return;
  if (bad_length(env, "labels", (*env)->GetArrayLength(env, labels), rows)) return;
  double *px = (*env)->GetPrimitiveArrayCritical(env, x, NULL);
  double *pl = px ? (*env)->GetPrimitiveArrayCritical(env, labels, NULL) : NULL;
  int rc = 1;
  if (px && pl) rc = agd_load_dense(H(h), dev, px, AGD_F64, pl, rows, d, d, storeDtype);
  if (pl) (*env)->ReleasePrimitiveArrayCritical(env, labels, pl, JNI_ABORT);
  if (px) (*env)->ReleasePrimitiveArrayCritical(env, x, px, JNI_ABORT);
  if (!px || !pl) { throw_msg(env, "java/lang/OutOfMemoryError", "could not pin the partition arrays"); return; }
  if (rc) fail(env, H(h));
}

/* SparseVector partitions as CSR (rowptr: rows + 1 longs from 0; idx / values: rowptr[rows] entries). */
JNIEXPORT void JNICALL JNI_FN(loadCsr)(JNIEnv *env, jobject self, jlong h, jint dev, jlongArray rowptr, jintArray idx,
                                       jdoubleArray values, jdoubleArray labels, jlong rows, jint d, jint storeDtype) {
  (void)self;
  if (rows < 0 || d <= 0) { throw_msg(env, "java/lang/IllegalArgumentException", "bad partition geometry"); return; }
  if (bad_length(env, "rowptr", (*env)->GetArrayLength(env, rowptr), rows + 1)) return;
  if (bad_length(env, "labels", (*env)->GetArrayLength(env, labels), rows)) return;
  jlong nnz = 0;
  (*env)->GetLongArrayRegion(env, rowptr, (jsize)rows, 1, &nnz);
  if (bad_length(env, "idx", (*env)->GetArrayLength(env, idx), nnz)) return;
  if (bad_length(env, "values", (*env)->GetArrayLength(env, values), nnz)) return;
  jlong *rp = (*env)->GetPrimitiveArrayCritical(env, rowptr, NULL);
  jint *ix = rp ? (*env)->GetPrimitiveArrayCritical(env, idx, NULL) : NULL;
  double *pv = ix ? (*env)->GetPrimitiveArrayCritical(env, values, NULL) : NULL;
  double *pl = pv ? (*env)->GetPrimitiveArrayCritical(env, labels, NULL) : NULL;
  int rc = 1;
  if (pl) rc = agd_load_csr(H(h), dev, (const int64_t *)rp, (const int32_t *)ix, pv, AGD_F64, pl, rows, d, storeDtype);
  if (pl) (*env)->ReleasePrimitiveArrayCritical(env, labels, pl, JNI_ABORT);
  if (pv) (*env)->ReleasePrimitiveArrayCritical(env, values, pv, JNI_ABORT);
  if (ix) (*env)->ReleasePrimitiveArrayCritical(env, idx, ix, JNI_ABORT);
  if (rp) (*env)->ReleasePrimitiveArrayCritical(env, rowptr, rp, JNI_ABORT);
  if (!pl) { throw_msg(env, "java/lang/OutOfMemoryError", "could not pin the partition arrays"); return; }
  if (rc) fail(env, H(h));
}

JNIEXPORT jlong JNICALL JNI_FN(rows)(JNIEnv *env, jobject self, jlong h, jint dev) {
  (void)env; (void)self;
  return (jlong)agd_rows(H(h), dev);
}

/* AcceleratedGradientDescent.run: returns the loss history; `weights` is updated in place; stats (length >= 8)
 * receives {iterations, passes, backtracks, restarts, converged, stopped_nan, final_L, device_ms_total}. */
JNIEXPORT jdoubleArray JNICALL JNI_FN(run)(JNIEnv *env, jobject self, jlong h, jint gradient, jint updater,
                                           jdouble convergenceTol, jint numIterations, jdouble regParam,
                                           jdoubleArray weights, jdouble L0, jdouble Lexact, jdouble beta,
                                           jdouble alpha, jboolean mayRestart, jint flags, jdoubleArray stats) {
  (void)self;
  agd_params p;
  agd_default_params(&p);
  p.convergence_tol = convergenceTol; p.num_iterations = numIterations; p.reg_param = regParam;
  p.L0 = L0; p.Lexact = Lexact; p.beta = beta; p.alpha = alpha; p.may_restart = mayRestart ? 1 : 0;
  p.gradient = gradient; p.updater = updater; p.flags = flags;
  /* agd_run reads and writes agd_dim(h) doubles through `weights` */
  if (bad_length(env, "initialWeights", (*env)->GetArrayLength(env, weights), agd_dim(H(h)))) return NULL;
  if ((*env)->GetArrayLength(env, weights) != agd_dim(H(h))) {
    throw_msg(env, "java/lang/IllegalArgumentException", "initialWeights.size differs from the feature count of the data");
    return NULL;
  }
  jint cap = numIterations > 0 ? numIterations : 1;
  jdoubleArray hist = (*env)->NewDoubleArray(env, cap);
  if (!hist) return NULL;
  double *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
  double *hh = w ? (*env)->GetDoubleArrayElements(env, hist, NULL) : NULL;
  int32_t n_hist = 0;
  agd_stats st;
  memset(&st, 0, sizeof st);
  int rc = 1;
  if (w && hh) rc = agd_run(H(h), &p, w, w, hh, &n_hist, &st);
  if (hh) (*env)->ReleaseDoubleArrayElements(env, hist, hh, 0);
  if (w) (*env)->ReleaseDoubleArrayElements(env, weights, w, rc ? JNI_ABORT : 0);
  if (!w || !hh) return NULL;   /* OutOfMemoryError is already pending */
  if (rc) { fail(env, H(h)); return NULL; }
  if (stats && (*env)->GetArrayLength(env, stats) >= 8) {
    const jdouble s[8] = {st.iterations, st.passes, st.backtracks, st.restarts, st.converged, st.stopped_nan,
                          st.final_L, st.device_ms_total};
    (*env)->SetDoubleArrayRegion(env, stats, 0, 8, s);
  }
  if (n_hist == cap) return hist;
  jdoubleArray out = (*env)->NewDoubleArray(env, n_hist);
  if (out && n_hist > 0) {
    double *src = (*env)->GetDoubleArrayElements(env, hist, NULL);
    if (src) {
      (*env)->SetDoubleArrayRegion(env, out, 0, n_hist, src);
      (*env)->ReleaseDoubleArrayElements(env, hist, src, JNI_ABORT);
    }
  }
  return out;
}

/* applySmooth at plug-in granularity: grad (d doubles) filled in place, returns loss/count. */
JNIEXPORT jdouble JNICALL JNI_FN(smooth)(JNIEnv *env, jobject self, jlong h, jint gradient, jdoubleArray weights,
                                         jdoubleArray grad) {
  (void)self;
  const jlong d = agd_dim(H(h));
  if (bad_length(env, "weights", (*env)->GetArrayLength(env, weights), d)) return 0.0;
  if (bad_length(env, "grad", (*env)->GetArrayLength(env, grad), d)) return 0.0;
  double *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
  double *g = w ? (*env)->GetDoubleArrayElements(env, grad, NULL) : NULL;
  double loss = 0.0;
  int64_t count = 0;
  int rc = 1;
  if (w && g) rc = agd_smooth(H(h), gradient, w, &loss, g, &count);
  if (g) (*env)->ReleaseDoubleArrayElements(env, grad, g, 0);
  if (w) (*env)->ReleaseDoubleArrayElements(env, weights, w, JNI_ABORT);
  if (!w || !g) return 0.0;
  if (rc) fail(env, H(h));
  return loss;
}
