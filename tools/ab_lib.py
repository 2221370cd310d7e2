"""A/B of two builds of libagd_b200.so on the same box: mean K1 time per launch on the headline shard.
usage: python tools/ab_lib.py libA.so libB.so[@key=value,...] ... [rows] [d]   (alternates the builds twice; AGD_FLAG_NO_FUSE so all run one-point sweeps)"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S


def time_lib(spec, rows, d, iters=6):
    path, _, opts = spec.partition("@")          # lib.so@key=value,key=value
    L = C.CDLL(path)
    L.agd_last_error.restype = C.c_char_p
    h = C.c_void_p()
    dev = (C.c_int32 * 1)(0)
    assert L.agd_create(dev, 1, C.byref(h)) == 0
    L.agd_generate.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_int32]
    assert L.agd_generate(h, rows, d, 1, 42, 0) == 0, L.agd_last_error(h)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        assert L.agd_set_option(h, k.encode(), v.encode()) == 0, L.agd_last_error(h)
    P, St = S._native.Params, S._native.Stats
    p = P(0.0, iters, 0.0, 1.0, float("inf"), 0.5, 0.9, 1, 0, 0, 2)
    w0, w, hist = np.zeros(d), np.zeros(d), np.zeros(iters)
    nh, st = C.c_int32(), St()
    out = []
    for rep in range(3):
        rc = L.agd_run(h, C.byref(p), w0.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                       hist.ctypes.data_as(C.c_void_p), C.byref(nh), C.byref(st))
        assert rc == 0, L.agd_last_error(h)
        out.append(st.k1_ms_total / st.k1_launches)
    one = sum(out[1:]) / 2
    p.flags = 0                                   # fused run: some sweeps evaluate two points
    two = []
    for rep in range(2):
        rc = L.agd_run(h, C.byref(p), w0.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                       hist.ctypes.data_as(C.c_void_p), C.byref(nh), C.byref(st))
        assert rc == 0, L.agd_last_error(h)
        if st.fused_passes:
            two.append((st.k1_ms_total - (st.k1_launches - st.fused_passes) * one) / st.fused_passes)
    L.agd_destroy(h)
    return out[1:] + ([-two[-1]] if two else [])    # negative entry = two-point sweep (ms)


if __name__ == "__main__":
    specs = [x for x in sys.argv[1:] if not x.isdigit()]
    nums = [int(x) for x in sys.argv[1:] if x.isdigit()]
    rows = nums[0] if nums else 10_000_000
    d = nums[1] if len(nums) > 1 else 1024
    res = {sp: [] for sp in specs}
    for _ in range(2):
        for sp in specs:
            res[sp] += [round(x, 4) for x in time_lib(sp, rows, d)]
    print(json.dumps({"rows": rows, "d": d, "k1_ms": res}))
