"""Turns the raw ncu artefacts in gpurun_out/ into the tracked summaries under profiles/ (round 1)."""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, Pf = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def launches():
    rows = list(csv.reader(open(os.path.join(G, "launches_step.csv"))))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv: continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn]).strip())
        agg[name][0] += 1; agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    micro = sum(v[1] for k, v in agg.items() if "gather4_kernel" in k)
    with open(os.path.join(Pf, "r1_launches_step_summary.md"), "w") as f:
        f.write("# Round 1 - ncu launch list of `bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv "
                "python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n\n"
                "4 train steps (1 warm-up + 1 timed + 1 e2e + 1 instrumented) followed by 13 launches of the 2^24-point "
                "voxel-gather micro-benchmark. Per-launch times are cold-cache and serialised: compare SHARES only.\n\n"
                f"{len(data)} launches, {tot / 1e6:.1f} ms in total under ncu.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
            f.write(f"| `{k[:100]}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% |\n")
        tc = sum(v[1] for k, v in agg.items() if "gemm_tc_kernel" in k)
        ff = sum(v[1] for k, v in agg.items() if "gemm::gemm_kernel" in k or "skinny" in k)
        f.write(f"\nMLP product kernels: tcgen05 `gemm_tc_kernel` {100 * tc / tot:.1f}% + FFMA / skinny kernels "
                f"{100 * ff / tot:.1f}% of the profiled time. Without the micro-benchmark gathers ({micro / 1e6:.1f} ms) the "
                f"product kernels are {100 * (tc + ff) / (tot - micro):.0f}% of the steps, in agreement with the CUDA-event "
                f"share bench.py reports (`roofline.share_of_step`).\n")

def full(rep, out, keep_extra=()):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"] + list(keep_extra)
    idx = [hdr.index(k) for k in keep if k in hdr]
    with open(os.path.join(Pf, out), "w") as f:
        w = csv.writer(f); w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]: w.writerow([r[i] for i in idx])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out_rows = []
    for r in rows[2:]:
        d = {hdr[i]: r[i] for i in idx}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d[k + ".bytes"] = float(r[hdr.index(k)]) * scale.get(units[hdr.index(k)], 1.0)
        d["time_unit"] = units[hdr.index("gpu__time_duration.sum")]
        out_rows.append(d)
    return out_rows

launches()
g = full("prof_gemm_tc.ncu-rep", "r1_gemm_tc_ncu_full.csv")
v = full("prof_gather4.ncu-rep", "r1_voxel_gather_ncu_full.csv")
# traffic of the dominant product shape (forward hidden layer, 262144 x 256 x 256): DRAM bytes per launch
tr = [r["dram__bytes_read.sum.bytes"] + r["dram__bytes_write.sum.bytes"] for r in g if "<1>" in r["Kernel Name"]]
vt = [r["dram__bytes_read.sum.bytes"] + r["dram__bytes_write.sum.bytes"] for r in v]
json.dump({"gemm_tc_kernel_fwd_hidden_layer_dram_bytes_per_launch": max(tr), "gemm_tc_all_captured": tr,
           "gather4_kernel_dram_bytes_per_launch": sum(vt) / len(vt),
           "gather4_kernel_ms_under_ncu": [float(r["gpu__time_duration.sum"]) for r in v],
           "source": "ncu --set full, profiles/r1_gemm_tc_ncu_full.csv and r1_voxel_gather_ncu_full.csv"},
          open(os.path.join(Pf, "r1_traffic.json"), "w"), indent=1)
print(open(os.path.join(Pf, "r1_traffic.json")).read())
for r in g: print(r["Kernel Name"][:40], r["gpu__time_duration.sum"], r["dram__bytes_read.sum"], r["dram__bytes_write.sum"], r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
for r in v: print(r["Kernel Name"][:40], r["gpu__time_duration.sum"], r["dram__bytes_read.sum"], r["dram__bytes_write.sum"], r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), r.get("l1tex__t_sector_hit_rate.pct"))
