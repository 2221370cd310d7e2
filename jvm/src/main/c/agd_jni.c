/*
 * agd_jni.c -- JNI shim over include/agd_b200.h for the Scala facade in
 * jvm/src/main/scala/org/apache/spark/mllib/optimization/NativeAGD.scala.
 *
 * NOT compiled in this image (no JDK: <jni.h> is absent); shipped as the binding a maintainer of
 * staple/spark-agd would add.  Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       jvm/src/main/c/agd_jni.c -Lspark-agd_b200 -lagd_b200 -o libagd_jni.so
 * Rules: pin arrays only for the duration of one C-ABI call, never retain JVM pointers, turn every
 * nonzero return code into a RuntimeException carrying agd_last_error.
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#include <jni.h>
#define AGD_HAVE_JNI 1
#endif
#endif

#ifdef AGD_HAVE_JNI
#include <stdint.h>
#include <string.h>

#include "agd_b200.h"

#define H(ptr) ((agd_handle *)(intptr_t)(ptr))

static jint fail(JNIEnv *env, agd_handle *h) {
  jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (ex) (*env)->ThrowNew(env, ex, agd_last_error(h));
  return -1;
}

JNIEXPORT jlong JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_create(JNIEnv *env, jobject self,
                                                                                       jintArray devices) {
  jsize n = (*env)->GetArrayLength(env, devices);
  jint *ids = (*env)->GetIntArrayElements(env, devices, NULL);
  agd_handle *h = NULL;
  int rc = agd_create((const int32_t *)ids, (int32_t)n, &h);
  (*env)->ReleaseIntArrayElements(env, devices, ids, JNI_ABORT);
  if (rc) { fail(env, NULL); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_destroy(JNIEnv *env, jobject self,
                                                                                       jlong h) {
  agd_destroy(H(h));
}

/* One RDD partition, already packed row-major by the facade (labels: rows doubles, x: rows*d doubles). */
JNIEXPORT void JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_loadDense(
    JNIEnv *env, jobject self, jlong h, jint dev, jdoubleArray x, jdoubleArray labels, jlong rows, jint d,
    jboolean storeF32) {
  double *px = (*env)->GetPrimitiveArrayCritical(env, x, NULL);
  double *pl = (*env)->GetPrimitiveArrayCritical(env, labels, NULL);
  int rc = agd_load_dense(H(h), dev, px, AGD_F64, pl, rows, d, d, storeF32 ? AGD_F32 : AGD_F64);
  (*env)->ReleasePrimitiveArrayCritical(env, labels, pl, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, x, px, JNI_ABORT);
  if (rc) fail(env, H(h));
}

/* SparseVector partitions as CSR. */
JNIEXPORT void JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_loadCsr(
    JNIEnv *env, jobject self, jlong h, jint dev, jlongArray rowptr, jintArray idx, jdoubleArray val,
    jdoubleArray labels, jlong rows, jint d) {
  jlong *rp = (*env)->GetPrimitiveArrayCritical(env, rowptr, NULL);
  jint *ix = (*env)->GetPrimitiveArrayCritical(env, idx, NULL);
  double *pv = (*env)->GetPrimitiveArrayCritical(env, val, NULL);
  double *pl = (*env)->GetPrimitiveArrayCritical(env, labels, NULL);
  int rc = agd_load_csr(H(h), dev, (const int64_t *)rp, (const int32_t *)ix, pv, AGD_F64, pl, rows, d, AGD_F64);
  (*env)->ReleasePrimitiveArrayCritical(env, labels, pl, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, val, pv, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, idx, ix, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, rowptr, rp, JNI_ABORT);
  if (rc) fail(env, H(h));
}

/* AcceleratedGradientDescent.run: returns the loss history; `weights` is updated in place. */
JNIEXPORT jdoubleArray JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_run(
    JNIEnv *env, jobject self, jlong h, jint gradient, jint updater, jdouble convergenceTol, jint numIterations,
    jdouble regParam, jdoubleArray weights, jdouble L0, jdouble Lexact, jdouble beta, jdouble alpha,
    jboolean mayRestart, jint flags) {
  agd_params p;
  agd_default_params(&p);
  p.convergence_tol = convergenceTol; p.num_iterations = numIterations; p.reg_param = regParam;
  p.L0 = L0; p.Lexact = Lexact; p.beta = beta; p.alpha = alpha; p.may_restart = mayRestart ? 1 : 0;
  p.gradient = gradient; p.updater = updater; p.flags = flags;
  jsize d = (*env)->GetArrayLength(env, weights);
  jint cap = numIterations > 0 ? numIterations : 1;
  jdoubleArray hist = (*env)->NewDoubleArray(env, cap);
  if (!hist) return NULL;
  double *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
  double *hh = (*env)->GetDoubleArrayElements(env, hist, NULL);
  int32_t n_hist = 0;
  agd_stats st;
  int rc = agd_run(H(h), &p, w, w, hh, &n_hist, &st);
  (*env)->ReleaseDoubleArrayElements(env, hist, hh, 0);
  (*env)->ReleaseDoubleArrayElements(env, weights, w, 0);
  (void)d;
  if (rc) { fail(env, H(h)); return NULL; }
  if (n_hist == cap) return hist;
  jdoubleArray out = (*env)->NewDoubleArray(env, n_hist);
  if (out && n_hist > 0) {
    double *src = (*env)->GetDoubleArrayElements(env, hist, NULL);
    (*env)->SetDoubleArrayRegion(env, out, 0, n_hist, src);
    (*env)->ReleaseDoubleArrayElements(env, hist, src, JNI_ABORT);
  }
  return out;
}

/* applySmooth at plug-in granularity: grad (d doubles) filled in place, returns loss/count. */
JNIEXPORT jdouble JNICALL Java_org_apache_spark_mllib_optimization_NativeAGD_00024_smooth(
    JNIEnv *env, jobject self, jlong h, jint gradient, jdoubleArray weights, jdoubleArray grad) {
  double *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
  double *g = (*env)->GetDoubleArrayElements(env, grad, NULL);
  double loss = 0.0;
  int64_t count = 0;
  int rc = agd_smooth(H(h), gradient, w, &loss, g, &count);
  (*env)->ReleaseDoubleArrayElements(env, grad, g, 0);
  (*env)->ReleaseDoubleArrayElements(env, weights, w, JNI_ABORT);
  if (rc) fail(env, H(h));
  return loss;
}
#endif /* AGD_HAVE_JNI */
