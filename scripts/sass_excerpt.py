"""SASS excerpts of the two hot loops for profiles/ ("do this" 2): the test-free fast path of
fdk_backproject_kernel<16,8,4> (16 voxel updates of one column and view) and the STEP phase of
mc_transport_kernel_v3<false,5,3,2> (one Woodcock step of two parked histories).  Read from the built library with
cuobjdump / nvdisasm (no GPU needed):  python scripts/sass_excerpt.py"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte_b200", "lib", "libmonte_gpu.so")


def kernel_sass(cubin, mangled_part):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(out) if l.startswith(".text.") and mangled_part in l][0]
    end = next((i for i in range(start + 1, len(out)) if out[i].startswith(".text.")), len(out))
    ln, seq = None, []
    for l in out[start:end]:
        m = re.search(r'//## File ".*", line (\d+)', l)
        if m:
            ln = int(m.group(1))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4}\*/", l):
            seq.append((ln, re.sub(r"\s+", " ", l.strip())))
    return seq


def opcode(t):
    w = t.split()
    return (w[2] if w[1].startswith("@") else w[1]).rstrip(";")


def excerpt(seq, lo, hi, title, note, path):
    idx = [k for k, (l, _) in enumerate(seq) if l is not None and lo <= l <= hi]
    a, b = idx[0], idx[-1]
    ops = Counter(opcode(t) for _, t in seq[a:b + 1])
    with open(path, "w") as f:
        f.write(title + "\n" + note + "\n")
        f.write("%d SASS instructions, offsets %s .. %s\n" % (b - a + 1, seq[a][1].split("*/")[0] + "*/", seq[b][1].split("*/")[0] + "*/"))
        f.write("opcode histogram: " + ", ".join("%s %d" % kv for kv in ops.most_common()) + "\n\n")
        for l, t in seq[a:b + 1]:
            f.write("%5s  %s\n" % (l if l is not None else "", t))
    return b - a + 1, ops


def main():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, stdout=subprocess.DEVNULL, check=True)
        src = open(os.path.join(ROOT, "monte_b200", "csrc", "fdk.cu")).read().splitlines()
        lo = [i + 1 for i, t in enumerate(src) if "if (u_ok && fmaxf(fabsf(w0v), fabsf(w_last)) <= hv_in)" in t][0] + 1
        hi = [i + 1 for i, t in enumerate(src) if i + 1 > lo and t.strip() == "continue;"][0] - 1
        seq = kernel_sass(os.path.join(d, "fdk.sm_100a.cubin"), "fdk_backproject_kernelILi16ELi8ELi4E")
        n, ops = excerpt(seq, lo, hi, "fdk_backproject_kernel<16,8,4>: the test-free fast path, fdk.cu:%d-%d -- 16 voxel updates (one (s,t) column, one view, 16 z-slices)" % (lo, hi),
                         "per update: 1 FFMA (axial coordinate) + F2I + I2F + FADD (cell, fraction) + 1 IMAD.WIDE (address) + 2 LDG.E.64 (the two row-pair texels) + "
                         "2 FFMA + FADD + FFMA (bilinear) + 1 FFMA (accumulate) = 12; no LDG.32, no clamp, no branch",
                         os.path.join(ROOT, "profiles", "r02_sass_fdk_fastpath.txt"))
        print("fdk fast path: %d instructions for 16 updates = %.2f per update; LDG.E.64: %d, 32-bit LDG: %d" %
              (n, n / 16.0, sum(v for k, v in ops.items() if k.startswith("LDG.E.64")), sum(v for k, v in ops.items() if k.startswith("LDG") and ".64" not in k)))
        src = open(os.path.join(ROOT, "monte_b200", "csrc", "mc.cu")).read().splitlines()
        lo = [i + 1 for i, t in enumerate(src) if "// ======== STEP" in t][0]
        hi = [i + 1 for i, t in enumerate(src) if "// ======== COLLIDE" in t][0] - 1
        seq = kernel_sass(os.path.join(d, "mc.sm_100a.cubin"), "mc_transport_kernel_v3ILb0ELi5ELi3ELi2ELb0ELb0ELb0E")
        # the Philox rounds and u01() are inlined from lines above the kernel: take the contiguous address range
        n, ops = excerpt(seq, lo, hi, "mc_transport_kernel_v3<false,5,3,2>: the STEP phase, mc.cu:%d-%d -- one Woodcock step of TWO parked histories of the lane" % (lo, hi),
                         "slot select, 3 x LDS.128 per slot, Philox2x32-10 (IMAD.WIDE.U32 + LOP3 per round, round keys from the constant bank), lg2, "
                         "3 FFMA, clip test, voxel index by the magic-number trick, one LDG.E.U8 of the label, acceptance, STS of the advanced slot",
                         os.path.join(ROOT, "profiles", "r02_sass_mc_step.txt"))
        print("mc STEP: %d instructions per two-slot visit; MUFU.LG2 %d, LDG.E.U8 %d, IMAD.WIDE.U32 %d" %
              (n, ops.get("MUFU.LG2", 0), sum(v for k, v in ops.items() if k.startswith("LDG.E.U8")), ops.get("IMAD.WIDE.U32", 0)))


if __name__ == "__main__":
    main()
