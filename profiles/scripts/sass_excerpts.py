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
            ub = [l for l in lines if "UBLKCP" in l]
            f.write(f"# TMA bulk stores (cp.async.bulk.global.shared::cta) in the DEFAULT fill kernel: {len(ub)} UBLKCP\n")
            f.write("\n".join(ub) + "\n")
            # the hot loop: the two LDS.128 of the ping-pong records up to the run-end flush
            # (v9: FSETP compares against |gate| -- the sign of the gate flags the end of a high-nibble group)
            hot = [i for i, l in enumerate(lines) if "FSETP.GEU.AND" in l and ", |R" in l]
            hot = [i for i in hot if hot[-1] - i < 120]
            if hot:
                lo = hot[0] - 16
                hi = hot[-1] + 75
                f.write("# hot loop (two candidates per trip; per candidate FFMA / FMUL x8, FSETP x5, predicated FMNMX x4, LDS.128 x2) and the\n"
                        "# two-level run-end flush (low nibble: 16 predicated FMNMX + 4 folds; high nibble once per group)\n")
                f.write("\n".join(lines[lo:hi]) + "\n")
        if "dist_kernelILi0ELb0" in name:
            f.write(f"\n## {name}: {len(lines)} instructions (K3 distances, bit-exact minimum image: FMUL/FADD/FSUB, no FFMA in the wrap)\n")
            f.write("\n".join(window(lines, lambda l: "MUFU.RSQ" in l, 45, 12)[:110]) + "\n")
        if "occ_band_kernel" in name or "occ_prep_kernel" in name:
            f.write(f"\n## {name}: {len(lines)} instructions\n")
            f.write("\n".join(window(lines, lambda l: "ATOM" in l or "RED" in l, 12, 6)[:60]) + "\n")
print("wrote", f.name)
