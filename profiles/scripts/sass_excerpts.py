"""Write the SASS of the hot loops of the shipped library under profiles/ (the Blackwell tell -- UBLKCP in the
default fill kernel -- had to be looked up by hand).  Usage: python profiles/scripts/sass_excerpts.py [tag]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "moleculekit_b200", "lib", "libmkb200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def functions():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            out[name].append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
    return out


def window(lines, pred, before, after):
    hits = [i for i, l in enumerate(lines) if pred(l)]
    if not hits:
        return []
    lo, hi = max(0, hits[0] - before), min(len(lines), hits[-1] + after)
    return lines[lo:hi]


fn = functions()
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpts.txt"), "w") as f:
    f.write(f"# cuobjdump -sass moleculekit_b200/lib/libmkb200.so (sm_100a), excerpts. Regenerate: python profiles/scripts/sass_excerpts.py {tag}\n")
    for name, lines in fn.items():
        if "occ_fill_runs_kernelILb1" in name:
            f.write(f"\n## {name}: {len(lines)} instructions\n")
            ub = [l for l in lines if "UBLKCP" in l or "UTMASTG" in l]
            f.write(f"# TMA stores in the DEFAULT fill kernel: {sum('UTMASTG' in l for l in ub)} UTMASTG (cp.async.bulk.tensor.4d: one 4 x 4 x 8-voxel block "
                    f"per instruction), {sum('UBLKCP' in l for l in ub)} UBLKCP (cp.async.bulk row copies: compact / to-host / non-uniform batches)\n")
            f.write("\n".join(ub) + "\n")
            # the hot loop (v10): packed FFMA2 / FMUL2 from the first pair of LDS.128 up to the end of the two-level run-end flush
            hot = [i for i, l in enumerate(lines) if "FFMA2" in l and "HI_LO, R" in l]
            hot = [i for i in hot if i - hot[0] < 80]
            if hot:
                lo = max(0, hot[0] - 14)
                hi = min(len(lines), hot[-1] + 95)
                f.write("# hot loop (two candidates per trip; per candidate LDS.128 x2, FFMA2 x3, FMUL2 x3, FADD, ISETP on the tag; FMNMX3 x4 per pair --\n"
                        "# the 5 A gate is the float overflow of FFMA2, no predicate) and the two-level run-end flush (low nibble: 16 predicated FMNMX + 4\n"
                        "# folds; high nibble once per group)\n")
                f.write("\n".join(lines[lo:hi]) + "\n")
            ld = [i for i, l in enumerate(lines) if "LDGSTS" in l]
            if ld:
                f.write("# record pass: cp.async (LDGSTS) gathers of the per-atom data straight into the sorted slots\n")
                f.write("\n".join(lines[max(0, ld[0] - 6):ld[-1] + 6]) + "\n")
        if "dist_kernelILi0ELb0" in name:
            f.write(f"\n## {name}: {len(lines)} instructions (K3 distances, bit-exact minimum image: FMUL/FADD/FSUB, no FFMA in the wrap)\n")
            f.write("\n".join(window(lines, lambda l: "MUFU.RSQ" in l, 45, 12)[:110]) + "\n")
        if "occ_band_kernel" in name or "occ_prep_kernel" in name:
            f.write(f"\n## {name}: {len(lines)} instructions\n")
            f.write("\n".join(window(lines, lambda l: "ATOM" in l or "RED" in l, 12, 6)[:60]) + "\n")
print("wrote", f.name)
