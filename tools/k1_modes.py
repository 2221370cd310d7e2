"""A/B of builds of libagd_b200.so on the same box, all three sweep forms of the ring kernel on the headline shard:
one point (AGD_FLAG_NO_FUSE), + loss at a second point (default run), two gradients (memoised run).
usage: python tools/k1_modes.py libA.so libB.so[@key=value,...] ... [rows] [d] [iters]
Each library is loaded through ctypes on its own handle; builds alternate twice.  Prints one JSON line."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S


def time_lib(spec, rows, d, iters):
    path, _, opts = spec.partition("@")          # lib.so@key=value,key=value
    L = C.CDLL(os.path.abspath(path))
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
    w0, w, hist = np.zeros(d), np.zeros(d), np.zeros(iters)
    nh, st = C.c_int32(), St()

    def run(flags):
        p = P(0.0, iters, 0.0, 1.0, float("inf"), 0.5, 0.9, 1, 0, 0, flags)
        rc = L.agd_run(h, C.byref(p), w0.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                       hist.ctypes.data_as(C.c_void_p), C.byref(nh), C.byref(st))
        assert rc == 0, L.agd_last_error(h)
        return dict(ms_per_launch=st.k1_ms_total / st.k1_launches, launches=int(st.k1_launches), fused=int(st.fused_passes),
                    iters_per_s=st.iterations / (st.device_ms_total / 1e3), passes=int(st.passes), hist_last=float(hist[nh.value - 1]),
                    w_sum=float(np.sum(w)))

    run(2)                                        # warm-up
    one = run(2)
    fused = run(0)
    memo = run(1)
    L.agd_destroy(h)
    nf = fused["fused"]
    two_ms = (fused["ms_per_launch"] * fused["launches"] - (fused["launches"] - nf) * one["ms_per_launch"]) / nf if nf else None
    nm = memo["fused"]
    twog_ms = (memo["ms_per_launch"] * memo["launches"] - (memo["launches"] - nm) * one["ms_per_launch"]) / nm if nm else None
    same = one["hist_last"] == fused["hist_last"] == memo["hist_last"] and one["w_sum"] == fused["w_sum"] == memo["w_sum"]
    return dict(one_ms=round(one["ms_per_launch"], 4), two_ms=round(two_ms, 4) if two_ms else None,
                two_grad_ms=round(twog_ms, 4) if twog_ms else None, iters_s_unfused=round(one["iters_per_s"], 2),
                iters_s_default=round(fused["iters_per_s"], 2), iters_s_memo=round(memo["iters_per_s"], 2),
                memo_sweeps=memo["launches"], memo_fused=nm, bit_identical_across_modes=same)


if __name__ == "__main__":
    specs = [x for x in sys.argv[1:] if not x.isdigit()]
    nums = [int(x) for x in sys.argv[1:] if x.isdigit()]
    rows = nums[0] if nums else 10_000_000
    d = nums[1] if len(nums) > 1 else 1024
    iters = nums[2] if len(nums) > 2 else 12
    res = {sp: [] for sp in specs}
    for _ in range(2):
        for sp in specs:
            res[sp].append(time_lib(sp, rows, d, iters))
    print(json.dumps({"rows": rows, "d": d, "iters": iters, "results": res}))
