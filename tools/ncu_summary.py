"""Summarise an `ncu --set full` report of `bench.py --quick` into profiles/<tag>_ncu_full_summary.csv and
profiles/ncu_traffic.json (DRAM bytes per launch per bench slot, read by bench.py for `roofline.traffic`).

    python tools/ncu_summary.py gpurun_out/prof_r1u.ncu-rep r1u
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = [i for i, c in enumerate(hdr) if c in ("Kernel Name", "Block Size", "Grid Size") or c in METRICS]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in keep]); w.writerow([units[i] for i in keep])
    for r in data:
        w.writerow([r[i] for i in keep])
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
# launch order of one step: gb_emit, proj_l0, agg_l0, proj_l1, agg_l1, ... (the capture filters these kernels)
seq = {"proj": 0, "agg": 0}


def tmpl_args(name):
    """template arguments of proj_umma_kernel<N, PMODE, EMODE, NH> as strings"""
    import re
    m = re.search(r"proj_umma_kernel<([^>]*)>", name)
    return [a.replace("(int)", "").strip() for a in m.group(1).split(",")] if m else []


kern = {}
for r in data:
    name = r[ki]
    tot = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    if "gat_aggregate" in name:
        slot = f"agg_l{seq['agg']}"; seq["agg"] += 1
    elif "proj_umma" in name and tmpl_args(name)[2:3] == ["1"]:      # EMODE == EM_HEADS
        slot = "heads_gemm"
    elif "proj_umma" in name or "proj_pair" in name:      # (proj_pair_kernel<MODE>: the cta_group::2 pair, layers 0..2)
        slot = f"proj_l{seq['proj']}"; seq["proj"] += 1
    elif "gb_emit" in name:
        slot = "gb_emit"
    else:
        continue
    kern[slot] = {"dram_bytes": int(tot), "kernel": name[:120]}
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump({"capture": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none, bench.py --quick, 2048^2 tile)",
               "kernels": kern}, f, indent=1)
print(json.dumps(kern, indent=1))
