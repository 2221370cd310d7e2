import json,sys
for f in sys.argv[1:]:
    try:
        t=[x for x in open(f).read().strip().splitlines() if x.startswith("{")]
        l=json.loads(t[-1]); p=l["passes"]
        print(f, l["n_gpus"], l["config"]["workload"][:60], "| value %.4g ex/s, %.1f it/s, ms/pass %.3f, k1 %.3f ms, frac %.3f, ar %.3f ms" % (l["value"], l["iters_per_sec"], l["ms_per_step"]*l["steps"]/p, l["roofline"]["ms_per_launch"], l["roofline"]["frac"], l["allreduce_ms_per_pass"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f).read()[-800:])
