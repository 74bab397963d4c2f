"""SASS opcode histogram of the tensor-core objects (CPU only: cuobjdump on build/*.o) -> profiles/r02_sass_histogram.md
usage: python tools/sass_histogram.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [("conv_h3", "forward / data-gradient implicit GEMM (all-fp16 split product, halo reuse, CTA pairs, persistent)"),
         ("conv_wgrad_umma", "weight gradient (MN-major operands, split-K)"),
         ("conv_umma", "TF32-main variants (fallback modes) and dispatch")]
WHAT = [("UTCHMMA", "tcgen05.mma (5th-gen tensor core, accumulator in TMEM)"), ("LDTM", "tcgen05.ld (TMEM -> registers)"),
        ("UTMALDG", "TMA tensor load (cp.async.bulk.tensor) global -> shared"), ("UTMASTG", "TMA tensor store"),
        ("UTCBAR", "tcgen05.commit -> mbarrier"), ("SYNCS", "mbarrier operations"), ("ELECT", "elect.sync (single-lane issue)"),
        ("UCGABAR_ARV", "cluster barrier arrive (CTA pairs)"), ("UCGABAR_WAIT", "cluster barrier wait"), ("REDG", "global reduction"),
        ("F2FP", "packed fp32 -> fp16 conversion (operand planes)"), ("STL", "local-memory store"), ("LDL", "local-memory load")]
out = ["# SASS opcode histogram of the tensor-core kernels (round 2, final kernels)\n",
       "command: `python tools/sass_histogram.py` = `cuobjdump -sass build/<file>.o` on the objects `__graft_entry__.build()` compiles for "
       "sm_100a (`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`); opcode = first token of each instruction, modifiers kept "
       "for the Blackwell-specific ones.\n"]
for name, desc in FILES:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", name + ".o")], capture_output=True, text=True).stdout
    kernels = sass.count("Function : ")
    ops = collections.Counter()
    total = 0
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", sass):
        op = m.group(1)
        total += 1
        special = any(op.startswith(k) for k, _ in WHAT)
        ops[op if special else op.split(".")[0]] += 1
    out.append(f"\n## `build/{name}.o` - {name}.cu - {desc}\n\n{kernels} kernels (template instantiations), {total} SASS instructions.\n")
    out.append("| opcode | count | what it is |\n|---|---:|---|")
    spec = [(o, c) for o, c in ops.items() if any(o.startswith(k) for k, _ in WHAT)]
    for o, c in sorted(spec, key=lambda kv: -kv[1]):
        out.append(f"| `{o}` | {c} | {next(w for k, w in WHAT if o.startswith(k))} |")
    for o, c in [(o, c) for o, c in ops.most_common(40) if not any(o.startswith(k) for k, _ in WHAT)][:12]:
        out.append(f"| `{o}` | {c} | |")
    notes = []
    notes.append("no `HMMA` / `HGMMA` (mma.sync / wgmma) instructions" if not any(o.startswith(("HMMA", "HGMMA")) for o in ops) else "HMMA present")
    notes.append("no `UTMASTG`: the epilogues store with plain `STG` (each thread owns one output pixel and writes its channels contiguously)"
                 if not any(o.startswith("UTMASTG") for o in ops) else "UTMASTG present")
    notes.append("no local-memory traffic (`STL` / `LDL`)" if not any(o.startswith(("STL", "LDL")) for o in ops) else "local-memory instructions present (see table)")
    out.append("\n" + "; ".join(notes) + ".")
open(os.path.join(ROOT, "profiles", "r02_sass_histogram.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
