"""Print the key metrics of every launch in an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 rd sectors(tex)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ smem"),
        ("smsp__inst_executed.sum", "inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("sm__cycles_elapsed.avg", "cycles")]
for r in rows[2:]:
    print("-" * 60)
    for key, label in want:
        for i, h in enumerate(hdr):
            if h == key:
                print("  %-22s %s %s" % (label, r[i][:90], units[i]))

# ---- optional: profiles/ncu_traffic.json for bench.py's roofline.traffic ----------------------------------------------------
# usage: python tools/ncu_summary.py file.ncu-rep --traffic-json profiles/ncu_traffic.json [--commit ID]
# The capture must come from tools/prof_kernels.py (every kernel at 24x64x64, fixed order per repetition: forward 3x3 128->128,
# 1x1 128->256, 1x1 256->128, then the three data gradients, then the three weight gradients); the LAST repetition is used
# (cold inputs: rotating buffers > L2).
if "--traffic-json" in sys.argv:
    import json, os, re
    outp = sys.argv[sys.argv.index("--traffic-json") + 1]
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else "unknown"
    col = {h: i for i, h in enumerate(hdr)}
    order = [("fwd", 3, [24, 64, 64, 128, 128]), ("fwd", 1, [24, 64, 64, 128, 256]), ("fwd", 1, [24, 64, 64, 256, 128]),
             ("dgrad", 3, [24, 64, 64, 128, 128]), ("dgrad", 1, [24, 64, 64, 256, 128]), ("dgrad", 1, [24, 64, 64, 128, 256]),
             ("wgrad", 3, [24, 64, 64, 128, 128]), ("wgrad", 1, [24, 64, 64, 128, 256]), ("wgrad", 1, [24, 64, 64, 256, 128])]
    launches = [r for r in rows[2:] if re.search(r"conv_tc\d?_kernel|wgrad\d?_tc_kernel", r[col["Kernel Name"]])]
    last = launches[-len(order):]
    def to_bytes(v, u):
        f = float(v.replace(",", ""))
        return f * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    ents = []
    for (direction, k, shape), r in zip(order, last):
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        ents.append({"direction": direction, "ksize": k, "shape": shape, "kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", ""),
                     "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                     "time_us": r[col["gpu__time_duration.sum"]], "file": os.path.basename(sys.argv[1]), "commit": commit})
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of tools/prof_kernels.py", "entries": ents},
              open(outp, "w"), indent=1)
    print("wrote", outp, len(ents), "entries")
