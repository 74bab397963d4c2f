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
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
for h, u, v in zip(hdr, units, vals):
    if h in WANT:
        out[h] = {"value": v, "unit": u}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    isrc, isamp = sh.index("Source"), sh.index("# Samples")
    data = srows[2:]
    tot = sum(int(r[isamp]) for r in data)
    top = sorted(data, key=lambda r: -int(r[isamp]))[:12]
    out["warp_samples_total"] = tot
    out["top_sampled_instructions"] = [{"sass": r[isrc].strip()[:90], "samples": int(r[isamp])} for r in top]
json.dump(out, open(rep.replace(".ncu-rep", ".json"), "w"), indent=1)
os.remove(rep)
print(rep, "->", out.get("gpu__time_duration.sum"))
