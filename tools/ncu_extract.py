"""Box-side: condenses an .ncu-rep (ncu --set full) into a small JSON of the metrics the profiles/ tables quote, then deletes it
(gpurun copies at most 64 MiB back).  usage: python tools/ncu_extract.py gpurun_out/r02_x.ncu-rep"""
import csv, io, json, os, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
PREFIXES = ("smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "sm__inst_executed_pipe_", "smsp__issue_active",
            "sm__pipe_fma", "sm__pipe_alu", "l1tex__lsu_writeback", "l1tex__data_bank_conflicts", "smsp__inst_executed_op_",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_st", "l1tex__t_requests_pipe_lsu_mem_global_op_st", "lts__t_sectors_op_write",
            "l1tex__data_pipe_lsu_wavefronts", "smsp__warps_eligible", "smsp__pcsamp_warps_issue_stalled")
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
for h, u, v in zip(hdr, units, vals):
    if h in WANT or h.startswith(PREFIXES):
        out[h] = {"value": v, "unit": u}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    isrc, isamp = sh.index("Source"), sh.index("# Samples")
    data = srows[2:]
    tot = sum(int(r[isamp]) for r in data)
    top = sorted(data, key=lambda r: -int(r[isamp]))[:40]
    out["warp_samples_total"] = tot
    out["top_sampled_instructions"] = [{"sass": r[isrc].strip()[:90], "samples": int(r[isamp])} for r in top]
    out["source_columns"] = sh
    keep = [i for i, h in enumerate(sh) if h.lower().startswith(("stall", "warp stall", "address", "#")) or "stall" in h.lower()]
    out["top_sampled_detail"] = [{sh[i]: r[i] for i in keep if r[i] not in ("", "0")} for r in top[:25]]
    # position in the program: index of each top instruction in the listing (tells main loop / epilogue apart)
    pos = {id(r): k for k, r in enumerate(data)}
    for d_, r in zip(out["top_sampled_instructions"], top):
        d_["index"] = pos[id(r)]
    out["instructions_total"] = len(data)
json.dump(out, open(rep.replace(".ncu-rep", ".json"), "w"), indent=1)
os.remove(rep)
print(rep, "->", out.get("gpu__time_duration.sum"))
