"""spark-agd_b200: B200-native accelerated (proximal) gradient descent -- the hot path of
staple/spark-agd (AcceleratedGradientDescent.optimize) behind the reference's own operator API.

Layout: csrc/ holds the sm_100a CUDA kernels and the C-ABI (include/agd_b200.h); optimization.py is
the host-side mirror of the reference interface.  The directory name carries a hyphen (the repo's
naming contract); import it as `spark_agd_b200` through the loader module at the repo root.
"""
from . import _native
from ._native import NativeError, build, exported_symbols
from .glm import (GeneralizedLinearAlgorithm, LinearRegressionWithAGD, LogisticRegressionWithAGD, SVMWithAGD, append_bias,
                  column_std)
from .optimization import (AcceleratedGradientDescent, Context, DeviceDataset, Gradient, GradientDescent,
                           HingeGradient, L1Updater, LeastSquaresGradient, LogisticGradient, MLUtils, RunStats,
                           SimpleUpdater, SquaredL2Updater, Updater, bf16_to_f32, run_with_stats)

__all__ = ["GeneralizedLinearAlgorithm", "LinearRegressionWithAGD", "LogisticRegressionWithAGD", "SVMWithAGD",
           "append_bias", "column_std", "AcceleratedGradientDescent", "Context", "DeviceDataset", "Gradient", "GradientDescent",
           "HingeGradient", "L1Updater", "LeastSquaresGradient", "LogisticGradient", "MLUtils", "NativeError", "RunStats",
           "SimpleUpdater", "SquaredL2Updater", "Updater", "bf16_to_f32", "build", "exported_symbols", "run_with_stats"]
