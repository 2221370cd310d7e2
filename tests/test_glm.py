"""GeneralizedLinearAlgorithm-style callers: host-side recipe on CPU, end-to-end on the GPU."""
import numpy as np
import pytest


def test_prepare_matches_mllib_recipe(agd):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((50, 4)) * np.array([1.0, 10.0, 0.1, 5.0]) + 3.0
    X[:, 2] = 7.0                                                    # zero-variance column
    alg = agd.LogisticRegressionWithAGD().setIntercept(True).setFeatureScaling(True)
    Xt, scale = alg.prepare(X)
    std = X.std(axis=0, ddof=1)
    assert Xt.shape == (50, 5) and np.all(Xt[:, -1] == 1.0)          # appendBias: 1.0 as the LAST feature
    np.testing.assert_allclose(scale[[0, 1, 3]], 1.0 / std[[0, 1, 3]], rtol=1e-14)
    assert scale[2] == 0.0 and np.all(Xt[:, 2] == 0.0)
    np.testing.assert_allclose(Xt[:, [0, 1, 3]].std(axis=0, ddof=1), 1.0, rtol=1e-12)
    plain, ones = agd.LinearRegressionWithAGD().prepare(X)
    assert plain.shape == X.shape and np.all(ones == 1.0)


@pytest.mark.gpu
def test_logistic_regression_with_agd(agd, ctx, oracle):
    rng = np.random.default_rng(1)
    n, d = 4000, 12
    X = rng.standard_normal((n, d)) * rng.uniform(0.5, 20.0, size=d)
    wt = rng.standard_normal(d) / np.sqrt(d)
    y = ((X / X.std(axis=0)) @ wt - 0.7 + rng.logistic(size=n) > 0).astype(np.float64)
    alg = agd.LogisticRegressionWithAGD(numIterations=40, regParam=0.01, convergenceTol=0.0)
    alg.setIntercept(True).setFeatureScaling(True)
    model = alg.run(ctx, y, X)
    # the same recipe with the oracle as the optimizer
    Xt, scale = alg.prepare(X)
    w0 = np.concatenate([np.zeros(d), [1.0]])       # appendBias(initialWeights) [mllib-1.3.0]: the intercept starts at 1.0
    ref = oracle.agd_run(oracle.Data(y, X=Xt), "logistic", "squared_l2", w0, convergence_tol=0.0,
                         num_iterations=40, reg_param=0.01)
    np.testing.assert_allclose(model.weights, ref.weights[:d] * scale, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(model.intercept, ref.weights[-1], rtol=1e-7)
    acc = (model.predict(X) == y).mean()
    assert acc > 0.6 and abs(model.intercept) > 0.05
