"""The JVM side of the drop-in boundary cannot run here (no JDK / Scala / Spark in the image), so it is
checked as far as a C compiler and a parser can: the JNI shim compiles warning-free against the JNI declarations it
uses (tests/stubs/jni.h), every C-ABI symbol it needs is exported by libagd_b200.so, and its exported
Java_..._NativeAGD_00024_* functions match the @native declarations of NativeAGD.scala in name and arity."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "jvm", "src", "main", "c", "agd_jni.c")
SCALA_DIR = os.path.join(ROOT, "jvm", "src", "main", "scala", "org", "apache", "spark", "mllib", "optimization")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
PREFIX = "Java_org_apache_spark_mllib_optimization_NativeAGD_00024_"


def _cflags():
    return ["-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include")]


def test_jni_shim_compiles_against_the_jni_declarations_it_uses():
    res = subprocess.run([GCC, "-fsyntax-only", *_cflags(), SHIM], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def _native_decls():
    src = open(os.path.join(SCALA_DIR, "NativeAGD.scala")).read()
    out = {}
    for m in re.finditer(r"@native\s+def\s+(\w+)\s*\(([^)]*)\)", src, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def _shim_functions():
    src = open(SHIM).read()
    out = {}
    for m in re.finditer(r"JNI_FN\((\w+)\)\s*\(([^)]*)\)", src, flags=re.S):
        params = [p for p in m.group(2).split(",") if p.strip()]
        assert params[0].strip().startswith("JNIEnv") and params[1].strip().startswith("jobject")
        out[m.group(1)] = len(params) - 2
    return out


def test_jni_symbols_match_the_native_declarations(tmp_path):
    decl, impl = _native_decls(), _shim_functions()
    assert decl and decl == impl, f"@native declarations {decl} != JNI functions {impl}"
    # the object file really exports the mangled names (object NativeAGD -> NativeAGD$ -> _00024)
    obj = str(tmp_path / "agd_jni.o")
    res = subprocess.run([GCC, "-c", "-fPIC", *_cflags(), SHIM, "-o", obj], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    nm = subprocess.run(["nm", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    defined = {ln.split()[-1] for ln in nm if " T " in ln}
    undefined = {ln.split()[-1] for ln in nm if ln.strip().startswith("U ")}
    assert defined == {PREFIX + n for n in decl}
    # everything the shim needs from the product is a C-ABI symbol the library exports
    import spark_agd_b200
    need = {u for u in undefined if u.startswith("agd_")}
    assert need and need <= set(spark_agd_b200._native.exported_symbols())
    assert undefined - need <= {"snprintf", "memset", "memcpy", "__stack_chk_fail", "_GLOBAL_OFFSET_TABLE_"}


def test_facade_keeps_the_reference_surface():
    """Same class, constructor, ten setters, optimize and run as AGD.scala:41-143,177-189; executor-side packing; the
    reference's two log lines."""
    src = open(os.path.join(SCALA_DIR, "AcceleratedGradientDescent.scala")).read()
    assert "class AcceleratedGradientDescent(private var gradient: Gradient, private var updater: Updater) extends Optimizer" in src
    for setter in ("setConvergenceTol", "setNumIterations", "setRegParam", "setL0", "setLexact", "setBeta", "setAlpha",
                   "setMayRestart", "setGradient", "setUpdater"):
        assert f"def {setter}(" in src
    assert "def optimize(data: RDD[(Double, Vector)], initialWeights: Vector): Vector" in src
    assert re.search(r"def run\(data: RDD\[\(Double, Vector\)\], gradient: Gradient, updater: Updater, convergenceTol: Double,\s+"
                     r"numIterations: Int, regParam: Double, initialWeights: Vector, L0: Double, Lexact: Double, beta: Double,\s+"
                     r"alpha: Double, mayRestart: Boolean\): \(Vector, Array\[Double\]\)", src)
    assert "mapPartitionsWithIndex" in src and "toLocalIterator" not in src and "agd.store" in src
    assert 'logWarning("Unable to compute loss function.")' in src and "run finished. Last 10 losses" in src
