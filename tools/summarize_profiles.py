"""Turns the raw ncu outputs in gpurun_out/ into the small tracked summaries under profiles/ (run on the CPU box)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "lts__t_sectors.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw(rep):
    out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[-1]
    return {k: f"{vals[hdr.index(k)]} {units[hdr.index(k)]}".strip() for k in KEYS if k in hdr}, vals[hdr.index("Kernel Name")]

def stalls(rep):
    out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        if len(r) < len(hdr):
            continue
        for s in st:
            try:
                tot[s] += int(r[ix[s]])
            except ValueError:
                pass
    T = sum(tot.values()) or 1
    return {k: round(100 * v / T, 1) for k, v in tot.most_common(8)}

def tobytes(s):
    v, u = s.split()
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]

ROUND = sys.argv[1] if len(sys.argv) > 1 else "r2"
CAPTURES = {
    "r1": [("k1_ring_logistic_10Mx1024_f32", "gpurun_out/k1_r1d_one.ncu-rep", 10_000_000, 1024, 4),
           ("k1_ring_two_point_logistic_10Mx1024_f32", "gpurun_out/k1_r1d_two.ncu-rep", 10_000_000, 1024, 4),
           ("k1_tc_ls_3Mx4096_bf16", "gpurun_out/k1tc_r1c.ncu-rep", 3_000_000, 4096, 2),
           ("k1_csr_hinge_4Mx1M_64nnz_f32", "gpurun_out/k1csr_r1.ncu-rep", 4_000_000, 1_000_000, 4)],
    "r2": [("k1_ring_logistic_10Mx1024_f32", "gpurun_out/k1_r2_one.ncu-rep", 10_000_000, 1024, 4),
           ("k1_ring_two_point_logistic_10Mx1024_f32", "gpurun_out/k1_r2_two.ncu-rep", 10_000_000, 1024, 4),
           ("k1_ring_two_gradient_logistic_10Mx1024_f32", "gpurun_out/k1_r2_twograd.ncu-rep", 10_000_000, 1024, 4),
           ("k1_tc_f32margins_ls_6.25Mx4096_bf16", "gpurun_out/k1tc_r2.ncu-rep", 6_250_000, 4096, 2)],
}
LAUNCHES = {"r1": "profiles/launches_r1_bench_n1.csv", "r2": "profiles/r2/launches_r2_bench_n1.csv"}
OUT = {"r1": "profiles/r1_summary.json", "r2": "profiles/r2/r2_summary.json"}
summary = {}
for tag, rep, rows, d, eb in CAPTURES[ROUND]:
    if not os.path.exists(rep):
        continue
    m, name = raw(rep)
    csr = tag.startswith("k1_csr")
    summary[tag] = {"kernel": name, "rows": rows, "d": d, "metrics": m, "warp_stall_pct": stalls(rep),
                    "algorithmic_bytes": rows * (64 * (4 + eb) + 16) if csr else rows * (d * eb + 8)}
    if tag == "k1_ring_logistic_10Mx1024_f32":
        json.dump({"kernel": "k1_ring_kernel<float,256,256,1,8,2,0>" if ROUND != "r1" else "k1_ring_kernel<float,256,256,1,8,2,false>", "rows": rows, "d": d,
                   "dram_bytes_read": tobytes(m["dram__bytes_read.sum"]), "dram_bytes_write": tobytes(m["dram__bytes_write.sum"]),
                   "algorithmic_bytes": rows * (d * eb + 8), "gpu_time_ms_under_ncu": float(m["gpu__time_duration.sum"].split()[0]),
                   "source": f"ncu --set full --clock-control none, tools/k1_prof.py logistic 10000000 (round {ROUND[1:]})"},
                  open("profiles/k1_traffic.json", "w"), indent=1)
# launch list of the bench command
rows = [r for r in csv.reader(open(LAUNCHES[ROUND])) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("agd::<", "")
    tot[name][0] += 1
    tot[name][1] += float(r[-1])
total = sum(v[1] for v in tot.values())
timed = {k: v for k, v in tot.items() if not k.startswith("synth")}
ttimed = sum(v[1] for v in timed.values())
summary["launch_list_bench_n1"] = {
    "command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 800 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --parity-iters 0",
    "kernels": {k: {"launches": v[0], "ms_total": round(v[1] / 1e6, 4), "share_of_all": round(v[1] / total, 4),
                    "share_of_timed_region_kernels": round(v[1] / ttimed, 4) if k in timed else None}
                for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
json.dump(summary, open(OUT[ROUND], "w"), indent=1)
print(json.dumps(summary, indent=1)[:3500])
