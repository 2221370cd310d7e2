package org.apache.spark.mllib.optimization

import org.apache.spark.annotation.DeveloperApi
import org.apache.spark.mllib.linalg.{DenseVector, SparseVector, Vector, Vectors}
import org.apache.spark.rdd.RDD

/** Drop-in for staple/spark-agd's optimizer: same package, class, constructor, setters, `optimize`
  * and `run`, but the loop executes natively on the box's B200s through NativeAGD (JNI over
  * include/agd_b200.h).  Source only -- no JVM in the build image.  Data path: each RDD partition
  * is packed into primitive arrays and handed to one GPU once (the analogue of `.cache()`); after
  * that no row ever crosses the JVM boundary again. */
@DeveloperApi
class AcceleratedGradientDescent(private var gradient: Gradient, private var updater: Updater) extends Optimizer {
  private var convergenceTol = 1e-4; private var numIterations = 100; private var regParam = 0.0
  private var L0 = 1.0; private var Lexact = Double.PositiveInfinity
  private var beta = 0.5; private var alpha = 0.9; private var mayRestart = true

  def setConvergenceTol(tol: Double): this.type = { convergenceTol = tol; this }
  def setNumIterations(iters: Int): this.type = { numIterations = iters; this }
  def setRegParam(reg: Double): this.type = { regParam = reg; this }
  def setL0(v: Double): this.type = { L0 = v; this }
  def setLexact(v: Double): this.type = { Lexact = v; this }
  def setBeta(v: Double): this.type = { beta = v; this }
  def setAlpha(v: Double): this.type = { alpha = v; this }
  def setMayRestart(v: Boolean): this.type = { mayRestart = v; this }
  def setGradient(g: Gradient): this.type = { gradient = g; this }
  def setUpdater(u: Updater): this.type = { updater = u; this }

  def optimize(data: RDD[(Double, Vector)], initialWeights: Vector): Vector =
    AcceleratedGradientDescent.run(data, gradient, updater, convergenceTol, numIterations, regParam,
      initialWeights, L0, Lexact, beta, alpha, mayRestart)._1
}

@DeveloperApi
object AcceleratedGradientDescent {
  /** agd_params.flags (include/agd_b200.h): 0 = every applySmooth evaluation of the reference is executed, the history
    * evaluation sharing one sweep over the shards with the next iteration's first one (bit-identical results);
    * -Dagd.flags=1 (AGD_FLAG_MEMOIZE_FX) / 2 (AGD_FLAG_NO_FUSE) select the other pass structures. */
  private def flags: Int = sys.props.get("agd.flags").map(_.trim.toInt).getOrElse(0)

  /** GPUs of this box; override with -Dagd.devices=0,1,... */
  private def devices: Array[Int] =
    sys.props.get("agd.devices").map(_.split(',').map(_.trim.toInt)).getOrElse(Array(0))

  def run(data: RDD[(Double, Vector)], gradient: Gradient, updater: Updater, convergenceTol: Double,
          numIterations: Int, regParam: Double, initialWeights: Vector, L0: Double, Lexact: Double, beta: Double,
          alpha: Double, mayRestart: Boolean): (Vector, Array[Double]) = {
    val g = NativeAGD.gradientId(gradient); val u = NativeAGD.updaterId(updater)   // fail before touching data
    val d = initialWeights.size
    val devs = devices
    val handle = NativeAGD.create(devs)
    try {
      // single-box deployment: partitions stream through the driver, partition p goes to GPU p % G
      data.mapPartitionsWithIndex { (p, rows) => Iterator((p, rows.toArray)) }.toLocalIterator.foreach {
        case (p, rows) if rows.nonEmpty =>
          val labels = rows.map(_._1)
          if (rows.forall(_._2.isInstanceOf[DenseVector])) {
            val x = new Array[Double](rows.length * d)
            var i = 0
            while (i < rows.length) { System.arraycopy(rows(i)._2.toArray, 0, x, i * d, d); i += 1 }
            NativeAGD.loadDense(handle, p % devs.length, x, labels, rows.length, d, storeF32 = false)
          } else {
            val sv = rows.map(_._2 match { case s: SparseVector => s; case v => Vectors.dense(v.toArray).toSparse })
            val rowptr = sv.scanLeft(0L)(_ + _.indices.length)
            NativeAGD.loadCsr(handle, p % devs.length, rowptr, sv.flatMap(_.indices), sv.flatMap(_.values), labels,
              rows.length, d)
          }
        case _ =>
      }
      val w = initialWeights.toArray.clone()
      val history = NativeAGD.run(handle, g, u, convergenceTol, numIterations, regParam, w, L0, Lexact, beta, alpha,
        mayRestart, flags)
      (Vectors.dense(w), history)
    } finally NativeAGD.destroy(handle)
  }
}
