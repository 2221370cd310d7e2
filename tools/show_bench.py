"""One-line summaries of bench.py JSON logs.  usage: python tools/show_bench.py file.log ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    if j.get("impl") == "reference":
        cb = j["cpu_baseline"]
        print(f"{f}: REF {j['value']:.3e} ex/s cores {cb['cores']} calib {cb.get('thread_calibration_examples_per_sec')} quota {cb.get('cgroup_cpu_quota')}")
        continue
    r = j["roofline"]
    step = j["ms_per_step"] * j["steps"]; k1 = r["ms_per_launch"] * r["launches"]
    print(f"{f}: N={j['n_gpus']} {j['config']['rows']}x{j['config']['d']} {j['config']['store']} value {j['value']:.3e} iters/s {j['iters_per_sec']:.2f} "
          f"| K1 {r['kernel']} frac {r['frac']:.3f} one {r['ms_per_launch_one_point']:.3f} two {r['ms_per_launch_two_point']} share {r['k1_share_of_step']:.4f} "
          f"nonK1/sweep {(step - k1) / j['sweeps'] * 1e3:.1f}us ar/pass {j['allreduce_ms_per_pass'] * 1e3:.1f}us "
          f"| memo {j['memoized']['iters_per_sec']:.2f} ({j['memoized']['sweeps']} sweeps) unfused {j['unfused']['iters_per_sec']:.2f} "
          f"| e2e {j['e2e']['value'] if j.get('e2e') else None} | {j['collective']} | clocks {j['clocks']['sm_mhz'] if j.get('clocks') else None}")
    if j.get("parity"):
        p = j["parity"]
        print(f"    parity pass={p['pass']} w_rel_err {p['w_rel_err']:.2e} loss {p['max_loss_rel_err']:.2e} twin {p['shards_equal_cpu_twin']} oracle {p['oracle']['seconds']:.1f}s x{p['oracle']['cores']}")
    if j.get("cpu_baseline"):
        cb = j["cpu_baseline"]
        print(f"    cpu {cb['value']:.3e} cores {cb['cores']} calib {cb.get('thread_calibration_examples_per_sec')}")
