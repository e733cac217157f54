"""Turn ncu outputs into the small text artefacts kept under profiles/ (development aid).

  python tools/summarize_ncu.py launches <ncu --csv log> <out.csv> "<command line that was profiled>"
  python tools/summarize_ncu.py full <ncu --page raw --csv export> <metrics.txt> <traffic.json> <particles>
"""
import csv, json, sys, collections

def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0].split("<")[0]

def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]; ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict(); cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")); u = r[ui]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        k = short(r[ki]); tot[k] = tot.get(k, 0.0) + ms; cnt[k] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: `{cmd}`\n# (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES with bench.py's `kernels`)\n")
        f.write(f"# captured launches: {sum(cnt.values())}, total {total:.2f} ms\nkernel,launches,total_ms,share\n")
        for k, v in sorted(tot.items(), key=lambda t: -t[1]):
            f.write(f"\"{k}\",{cnt[k]},{v:.3f},{v / total:.3f}\n")

WANT = ["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
KEYS = {"step_perceive_pairs_kernel": "step_perceive", "step_perceive_pairs_fired_kernel": "step_perceive", "step_mlp_tc_kernel": "step_mlp", "step_bwd_tc_kernel": "step_bwd",
        "mlp_wgrad_tc_kernel": "mlp_wgrad", "step_adjoint_pairs_kernel": "step_adjoint", "fire_fill_kernel": "fire_list", "fire_list_onepass_kernel": "fire_list",
        "step_bwd_unfired_kernel": "step_bwd_unfired"}

def full(src, txt, js, particles):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, traffic = set(), {}
    with open(txt, "w") as f:
        f.write(f"# per-kernel ncu --set full metrics (one launch each, {particles} particles)\n")
        for r in rows[2:]:
            k = short(r[ki])
            if k in seen:
                continue
            seen.add(k)
            f.write(f"## {r[ki][:110]}\n")
            vals = dict(zip(hdr, zip(r, units)))
            for m in WANT:
                if m in vals:
                    f.write(f"  {m} = {vals[m][0]} {vals[m][1]}\n")
            base = k.split("::")[-1]
            if base in KEYS and "dram__bytes_read.sum" in vals:
                def b(name):
                    v, u = vals[name]; return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                traffic[KEYS[base]] = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
    json.dump({"source": f"{txt} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, {particles} particles)",
               "particles": int(particles), "bytes_per_launch": traffic}, open(js, "w"), indent=1)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
