"""Copy the outputs of tools/profile_round.sh from gpurun_out/ into profiles/ (curated CSV, traffic.json, bench lines)."""
import csv, json, shutil, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r01s2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
rows = list(csv.reader(open(f"{go}/{tag}_full_raw.csv")))
h, u = rows[0], rows[1]
prefixes = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct', 'lts__throughput.avg.pct',
            'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct', 'sm__mem_tensor_cycles_active.avg.pct',
            'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct', 'smsp__issue_active.avg.pct',
            'sm__inst_executed_pipe_xu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct',
            'sm__inst_executed_pipe_tmem', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit',
            'sm__warps_active.avg.pct', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'smsp__pcsamp_warps_issue_stalled',
            'smsp__average_warp', 'launch__shared_mem', 'sm__cycles_elapsed.max')
keep = [c for c in h if c in ('ID', 'Kernel Name', 'Block Size', 'Grid Size') or any(c.startswith(p) for p in prefixes)]
idx = [h.index(c) for c in keep]
with open(f"{pr}/{tag}_ncu_full_hot_kernels.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        if len(r) == len(h):
            w.writerow([r[i] for i in idx])
mul = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
per = {}
for r in rows[2:]:
    k = r[h.index('Kernel Name')]
    per[k] = float(r[h.index('dram__bytes_read.sum')]) * mul[u[h.index('dram__bytes_read.sum')]] + \
        float(r[h.index('dram__bytes_write.sum')]) * mul[u[h.index('dram__bytes_write.sum')]]
    t = h.index('sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed')
    print(f"{float(r[h.index('gpu__time_duration.sum')]):8.1f} us  dram {per[k] / 1e6:8.1f} MB  tensor {r[t][:5]}%  {k[:80]}")
g = lambda *names: int(sum(v for k, v in per.items() if any(n in k for n in names)))
json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) from the ncu --set full capture of `python bench.py "
           "--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-breakdown --no-graph` (profiles/%s_ncu_full_hot_kernels.csv, "
           "tools/profile_round.sh); keyed like bench.py's kernel groups (group = sum of its kernels)" % tag,
           "wsum_fwd": g('wsum_fwd'), "wsum_bwd": g('wsum_bwd'), "vq_fwd": g('Sweep1', 'Sweep2', 'vq_select', 'vq_colsum'),
           "vq_bwd": g('Sweep3', 'SweepT', 'StoreEpi<2>'), "nce_fwd_bwd": g('Nce')}, open(f"{pr}/traffic.json", "w"), indent=2)
shutil.copy(f"{go}/{tag}_bench.json", f"{pr}/{tag}_bench.json")
shutil.copy(f"{go}/{tag}_launches.csv", f"{pr}/{tag}_launches_bench_steps2.csv")
for extra in ("aux_kernels",):
    if os.path.exists(f"{go}/{tag}_{extra}.json"):
        shutil.copy(f"{go}/{tag}_{extra}.json", f"{pr}/{tag}_{extra}.json")
d = json.load(open(f"{pr}/{tag}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["clocks"]["reasons"])
for k, v in d["kernels"].items():
    print(" ", k, round(v["ms"], 4), None if v["frac"] is None else round(v["frac"], 3))
print(" e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in (d.get("configs") or {}).items():
    print(" ", k, v.get("ms_per_step"), v.get("error"))
