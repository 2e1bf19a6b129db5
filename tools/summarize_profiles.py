"""Turn the ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/."""
import csv, io, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("HVS_PROFILE_OUT") or os.path.join(ROOT, "profiles")      # on the GPU box: gpurun_out/profiles (only gpurun_out/ travels back)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarize(rep, out_name, note):
    hdr, units, rows = raw_rows(rep)
    lines = [f"# {os.path.basename(rep)} -- {note}", "# ncu --set full --clock-control none (cold-cache, serialised launches)"]
    res = []
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].split("::")[-1]
        lines.append(f"\n## {short}")
        d = {}
        for k in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                lines.append(f"{k:90s} {v} {u}")
                d[k] = (v, u)
        res.append((short, d))
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")
    return res


def launch_list(csv_path, out_name):
    rows = list(csv.reader(open(csv_path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].split("::")[-1][:60]
        val = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)      # -> microseconds
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += val * scale
    total = sum(v[1] for v in tot.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none [-c 600]  python bench.py --steps K --warmup 3 --e2e-steps 1 --no-cpu-baseline [--skip-hybrid]",
             "# (every launch of the command incl. input generation, warm-up and the e2e leg; compare SHARES, not absolutes)",
             f"{'kernel':62s} {'launches':>8s} {'total_us':>12s} {'share':>7s}"]
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:62s} {n:8d} {us:12.1f} {us / total:7.2%}")
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    reps = {"r1_fwd_v1.ncu-rep": ("r01_fwd_v1_thread_per_token_coef.txt", "forward v1: one coefficient warp, accurate expf/div (T=262144)"),
            "r1_fwd_v3.ncu-rep": ("r01_fwd_v3.txt", "forward v3 (shipped): 2 coefficient warps, 4 lanes/token (T=262144)"),
            "r1_bwd_v1.ncu-rep": ("r01_bwd_v1_smem_resident.txt", "backward v1: tiles resident in shared memory, 1 coefficient warp (T=262144)"),
            "r1_bwd_v3.ncu-rep": ("r01_bwd_v3_two_pass_l2.txt", "backward v3: two passes per tile, second re-loaded through L2 (T=262144)"),
            "r1_bwd_v5.ncu-rep": ("r01_bwd_v5_tmem.txt", "backward v5 (shipped): tiles parked in tensor memory, 3 coefficient warps (T=262144)"),
            "r1_bench_kernels.ncu-rep": ("r01_bench_kernels_T1M.txt", "two-kernel backward era: bench.py kernels at the benchmark size T=2^20 (one launch each)"),
            "r1_fused_v1.ncu-rep": ("r01_fused_v1_front_thread_bound.txt", "fused backward v1: 64 small tcgen05.mma per tile issued by one divergent thread, 4-lane coefficient chain (T=262144)"),
            "r1_fused_v5.ncu-rep": ("r01_fused_v5_bank_conflicts.txt", "fused backward v5: G on the warp MMA path with 4-way conflicting ldmatrix rows + unpadded M-pair rows (T=262144)"),
            "r1_fused_v6.ncu-rep": ("r01_fused_v6.txt", "fused backward v6: conflicts fixed, scaling-form reverse sweep, early forward (T=262144)"),
            "r1b_bench_kernels.ncu-rep": ("r01b_bench_kernels_T1M.txt", "fused backward with 64 small tcgen05.mma per tile (3.63 ms): forward (saving statistics) + fused backward + finalize at T=2^20"),
            "r1c_fused_planB.ncu-rep": ("r01c_fused_mma_contention.txt", "fused backward, 64 small MMAs per tile, G issued by the coefficient warps: mma.sync of the dx pass 'math'-throttled while tcgen05.mma are in flight (T=262144)"),
            "r1c_bench_kernels.ncu-rep": ("r01c_bench_kernels_T1M.txt", "fused backward with 32 big tcgen05.mma per tile, 4-lane coefficient warps (2.99 ms), forward before the tensor-memory parking (1.69 ms), finalize; T=2^20"),
            "r1d_bench_kernels.ncu-rep": ("r01d_bench_kernels_T1M.txt", "one step earlier than r01e: before the per-stream dx stores and the coalesced finalize; T=2^20"),
            "r1e_bench_kernels.ncu-rep": ("r01e_bench_kernels_T1M.txt", "one step earlier than r01f (forward before the half-tile y stores, finalize before the 4-group split); T=2^20"),
            "r1f_bench_kernels.ncu-rep": ("r01f_bench_kernels_T1M.txt", "one step earlier than r01g (forward with one shuffle step per Sinkhorn iteration, 1.60 ms); T=2^20"),
            "r1g_bench_kernels.ncu-rep": ("r01g_bench_kernels_T1M.txt", "one step earlier than r01h (backward before the lean reverse sweep and the uniform warp index, 2.86 ms); T=2^20"),
            "r1h_bench_kernels.ncu-rep": ("r01h_bench_kernels_T1M.txt", "SHIPPED training path at T=2^20: forward (saving statistics; shuffle-free Sinkhorn on scalings, fragments parked in tensor memory, y in half tiles), fused single-pass backward, finalize")}
    reps.update({
        "r2_train.ncu-rep": ("r02_k2_train_kernels.txt", "round 2 training kernels: ManifoldHyperConnection(512,4) T=2^15 and (64,4) T=2^17 forward+backward on _K2TokenPathFn "
                             "(k2_gemm_kernel in all its modes: K-major forward, B MN-major data gradients with the GELU' epilogue, A+B MN-major split-K weight gradients; "
                             "column sums, partial reduction, LayerNorm backward) + the general-shape K1 backward (n=2, C=256, T=2^16); every launch of one forward+backward, in launch order"),
        "r2_bench_kernels.ncu-rep": ("r02_bench_kernels_T1M.txt", "round 2 re-capture of the SHIPPED K1 training path at T=2^20 (kernels unchanged since r01h): forward saving statistics, fused single-pass backward, finalize"),
    })
    traffic = {}
    for rep, (out, note) in reps.items():
        path = os.path.join(g, rep)
        if not os.path.exists(path):
            continue
        res = summarize(path, out, note)
        if rep in ("r1h_bench_kernels.ncu-rep", "r2_bench_kernels.ncu-rep"):
            for short, d in res:
                def gb(k):
                    v, u = d[k]
                    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                import re
                traffic[re.sub(r"<[^>]*>", "", short) + "_bytes_per_launch"] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    if traffic:
        traffic["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, T=2^20 (profiles/r02_bench_kernels_T1M.txt, the default <kAdaptive = false> instantiations)"
        json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    if os.path.exists(os.path.join(g, "r1_launches.csv")):
        launch_list(os.path.join(g, "r1_launches.csv"), "r01_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1b_launches.csv")):
        launch_list(os.path.join(g, "r1b_launches.csv"), "r01b_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1c_launches.csv")):
        launch_list(os.path.join(g, "r1c_launches.csv"), "r01c_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1d_launches.csv")):
        launch_list(os.path.join(g, "r1d_launches.csv"), "r01d_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1e_launches.csv")):
        launch_list(os.path.join(g, "r1e_launches.csv"), "r01e_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1f_launches.csv")):
        launch_list(os.path.join(g, "r1f_launches.csv"), "r01f_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1g_launches.csv")):
        launch_list(os.path.join(g, "r1g_launches.csv"), "r01g_launch_list.txt")
    if os.path.exists(os.path.join(g, "r1h_launches.csv")):
        launch_list(os.path.join(g, "r1h_launches.csv"), "r01h_launch_list.txt")
    if os.path.exists(os.path.join(g, "r2_launches.csv")):
        launch_list(os.path.join(g, "r2_launches.csv"), "r02_launch_list.txt")
    print(open(os.path.join(OUT, "traffic.json")).read() if traffic else "no traffic")
