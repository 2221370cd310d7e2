"""LIBSVM ingest (MLUtils.loadLibSVMFile [mllib-1.3.0]): host-side parser on CPU, CSR load + run on the GPU."""
import numpy as np
import pytest

TEXT = """# a comment line
1 1:0.5 3:-2.25 7:1e-3

0 2:4
-1.5 1:1 2:2 3:3 4:4 5:5 6:6 7:7
1
"""


def test_parse_libsvm(agd, tmp_path):
    p = tmp_path / "a.libsvm"
    p.write_text(TEXT)
    y, rowptr, idx, val, d = agd.MLUtils.parseLibSVMFile(str(p))
    assert d == 7
    np.testing.assert_array_equal(y, [1.0, 0.0, -1.5, 1.0])
    np.testing.assert_array_equal(rowptr, [0, 3, 4, 11, 11])
    np.testing.assert_array_equal(idx, [0, 2, 6, 1, 0, 1, 2, 3, 4, 5, 6])
    np.testing.assert_array_equal(val, [0.5, -2.25, 1e-3, 4, 1, 2, 3, 4, 5, 6, 7])
    assert agd.MLUtils.parseLibSVMFile(str(p), numFeatures=10)[4] == 10


@pytest.mark.parametrize("bad", ["1 3:1 2:1\n", "1 0:1\n", "x 1:1\n", "1 1:\n"])
def test_parse_libsvm_rejects_bad_input(agd, tmp_path, bad):
    p = tmp_path / "bad.libsvm"
    p.write_text(bad)
    with pytest.raises(ValueError):
        agd.MLUtils.parseLibSVMFile(str(p))
    with pytest.raises(ValueError):
        agd.MLUtils.parseLibSVMFile(str(tmp_path / "missing.libsvm"))


def test_parse_libsvm_rejects_indices_beyond_int32(agd, tmp_path):
    p = tmp_path / "big.libsvm"
    p.write_text("1 1:1 4294967297:2\n")          # 2^32 + 1 would truncate to 1 and slip past the ordering check
    with pytest.raises(ValueError, match="int32"):
        agd.MLUtils.parseLibSVMFile(str(p))


@pytest.mark.gpu
def test_load_libsvm_reports_parse_errors_through_the_handle(agd, ctx, tmp_path):
    p = tmp_path / "bad.libsvm"
    p.write_text("1 3:1 2:1\n")
    with pytest.raises(agd.NativeError, match="line 1: indices must be one-based and ascending"):
        agd.MLUtils.loadLibSVMFile(ctx, str(p))


@pytest.mark.gpu
def test_load_libsvm_and_run(agd, ctx, oracle, tmp_path):
    rng = np.random.default_rng(0)
    n, d = 400, 60
    lines = []
    for i in range(n):
        k = rng.integers(1, 12)
        cols = np.sort(rng.choice(d, size=k, replace=False)) + 1
        lines.append(f"{int(rng.random() > 0.5)} " + " ".join(f"{c}:{rng.standard_normal():.17g}" for c in cols))
    p = tmp_path / "data.libsvm"
    p.write_text("\n".join(lines) + "\n")
    y, rowptr, idx, val, dd = agd.MLUtils.parseLibSVMFile(str(p), numFeatures=d)
    data = agd.MLUtils.loadLibSVMFile(ctx, str(p), numFeatures=d)
    assert data.d == d and data.local_rows(0) == n
    w0 = np.zeros(d)
    w, hist, st = agd.run_with_stats(data, agd.HingeGradient(), agd.SquaredL2Updater(), 0.0, 10, 0.05, w0)
    ref = oracle.agd_run(oracle.Data(y, csr=(rowptr, idx, val), d=d), "hinge", "squared_l2", w0, convergence_tol=0.0,
                         num_iterations=10, reg_param=0.05)
    np.testing.assert_allclose(hist, ref.loss_history, rtol=1e-10)
    assert np.linalg.norm(w - ref.weights) / np.linalg.norm(ref.weights) < 1e-8
    data.close()
