package org.apache.spark.mllib.optimization

/** JNI surface of libagd_jni.so (jvm/src/main/c/agd_jni.c) over include/agd_b200.h.
  * Source only: this image has no JVM, so the facade is not compiled or run here. */
private[optimization] object NativeAGD {
  System.loadLibrary("agd_jni")

  @native def create(devices: Array[Int]): Long
  @native def destroy(handle: Long): Unit
  @native def loadDense(handle: Long, dev: Int, x: Array[Double], labels: Array[Double], rows: Long, d: Int,
                        storeF32: Boolean): Unit
  @native def loadCsr(handle: Long, dev: Int, rowptr: Array[Long], idx: Array[Int], values: Array[Double],
                      labels: Array[Double], rows: Long, d: Int): Unit
  @native def run(handle: Long, gradient: Int, updater: Int, convergenceTol: Double, numIterations: Int,
                  regParam: Double, weights: Array[Double], L0: Double, Lexact: Double, beta: Double, alpha: Double,
                  mayRestart: Boolean, flags: Int): Array[Double]
  @native def smooth(handle: Long, gradient: Int, weights: Array[Double], grad: Array[Double]): Double

  /** Closed enums of include/agd_b200.h; anything else has no GPU implementation and is rejected. */
  def gradientId(g: Gradient): Int = g match {
    case _: LogisticGradient => 0
    case _: LeastSquaresGradient => 1
    case _: HingeGradient => 2
    case other => throw new UnsupportedOperationException(
      s"${other.getClass.getName} has no B200 kernel (Logistic/LeastSquares/Hinge only; there is no CPU fallback)")
  }
  def updaterId(u: Updater): Int = u match {
    case _: SimpleUpdater => 0
    case _: SquaredL2Updater => 1
    case _: L1Updater => 2
    case other => throw new UnsupportedOperationException(
      s"${other.getClass.getName} has no B200 kernel (Simple/SquaredL2/L1 only; there is no CPU fallback)")
  }
}
