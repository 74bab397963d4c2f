"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_summary.md "<command that was profiled>" """
import collections
import csv
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = []
with open(src) as f:
    for line in f:
        if line.startswith('"ID"'):
            break
    for r in csv.reader(f):
        if len(r) >= 15:
            rows.append((r[4], r[7], r[8], float(r[14])))
# one optimiser step = the launches between the last two groups of Adam launches
adam = [i for i, r in enumerate(rows) if "adam" in r[0]]
groups = []
for i in adam:
    if not groups or i - groups[-1][-1] > 50:
        groups.append([i])
    else:
        groups[-1].append(i)
if len(groups) >= 2:
    a, b = groups[-2][-1] + 1, groups[-1][-1] + 1
    scope = f"launches {a}..{b} = the last complete optimiser step of the run"
else:
    a, b = 0, len(rows)
    scope = "all captured launches (fewer than two optimiser steps in the capture)"
step = rows[a:b]
tot = sum(r[3] for r in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, bs, gs, t in step:
    k = re.sub(r"\(.*", "", n).replace("void ", "")[:100]
    agg[k][0] += 1
    agg[k][1] += t
ours = sum(v[1] for k, v in agg.items() if k.startswith("pvg::"))
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary\n\ncommand: `{cmd}`\n\nscope: {scope}; {len(step)} launches, {tot / 1e6:.1f} ms of kernel time "
            f"(cold-cache, serialised: compare SHARES, not absolutes); hand-written `pvg::` kernels = {100 * ours / tot:.1f} % of it.\n\n")
    f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f} % |\n")
    f.write("\n## largest individual launches of the tensor-core kernels (grid -> avg us)\n\n")
    for key in ("conv_umma_kernel", "conv_wgrad_umma_kernel", "conv_fwd_simt", "conv_wgrad_simt"):
        g = collections.defaultdict(lambda: [0, 0.0])
        for n, bs, gs, t in step:
            if key in n:
                m = re.search(key + r"<[^>]*>", n)
                g[(m.group(0) if m else key, gs)][0] += 1
                g[(m.group(0) if m else key, gs)][1] += t
        for (kn, gs), v in sorted(g.items(), key=lambda kv: -kv[1][1])[:8]:
            f.write(f"- `{kn}` grid {gs}: n={v[0]}, total {v[1] / 1e6:.2f} ms, avg {v[1] / v[0] / 1e3:.1f} us\n")
print(open(dst).read()[:3000])
