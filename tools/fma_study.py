"""What would north_star's 1-ULP budget buy in the restoration-filter chain? (experiment, not part of the build)

The shipped kernels compute every a*b+c as the reference's generic path does (multiply, round, add, round) and match the
oracle bit for bit. This script makes a COPY of the column-strip filter kernel's phase functions
(jxl_oxide_b200/csrc/kernels/filter_strip.cuh) in which the multiply-adds of Gaborish, the EPF distances / weights /
weighted sums and the colour stage are contracted to one fused multiply-add each (divisions stay IEEE divisions), and reports

  1. the error that costs: the copy is compiled for the host (tests/emu harness: every CTA run thread by thread, phase by
     phase; std::fmaf is the same correctly rounded operation as the device's FFMA) and its output is compared with the
     oracle's, sample by sample, in units in the last place of the oracle's value;
  2. the instructions it saves: static SASS size of strip_filter_kernel<2, sRGB> with and without contraction (nvcc, sm_100a).

Nothing here touches the product sources; the copies live under tools/_fma_build/ (git-ignored).
    python tools/fma_study.py            # writes profiles/r02_fma_study.md
"""
import collections
import ctypes
import os
import re
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
BUILD = os.path.join(ROOT, "tools", "_fma_build")

HELPER_OLD = "JXLB_FS float fs_absdiff(float a, float b) { return fabsf(fs_sub(a, b)); }"
HELPER_NEW = HELPER_OLD + "\nJXLB_FS float fs_mad(float a, float b, float c) { return fs_fma(a, b, c); }  // contracted (experiment copy)"
REPLACEMENTS = [
    ("o[i * kWX] = fs_mul(fs_add(fs_add(mc, fs_mul(sum_side, w0)), fs_mul(sum_diag, w1)), g);",
     "o[i * kWX] = fs_mul(fs_mad(sum_diag, w1, fs_mad(sum_side, w0, mc)), g);"),
    ("        d01[i] = fs_add(d01[i], t01);\n        d10[i] = fs_add(d10[i], t10);",
     "        d01[i] = fs_fma(sc, p01, d01[i]);\n        d10[i] = fs_fma(sc, p10, d10[i]);"),
    ("const float wu = fmaxf(fs_add(1.0f, fs_mul(du, nis)), 0.0f);", "const float wu = fmaxf(fs_mad(du, nis, 1.0f), 0.0f);"),
    ("const float wd = fmaxf(fs_add(1.0f, fs_mul(dd, nis)), 0.0f);", "const float wd = fmaxf(fs_mad(dd, nis, 1.0f), 0.0f);"),
    ("const float wl = fmaxf(fs_add(1.0f, fs_mul(dl, nis)), 0.0f);", "const float wl = fmaxf(fs_mad(dl, nis, 1.0f), 0.0f);"),
    ("const float wr = fmaxf(fs_add(1.0f, fs_mul(dr, nis)), 0.0f);", "const float wr = fmaxf(fs_mad(dr, nis, 1.0f), 0.0f);"),
    ("float sum = fs_add(ce[c], fs_mul(wu, up[c]));", "float sum = fs_mad(wu, up[c], ce[c]);"),
    ("sum = fs_add(sum, fs_mul(wd, dn[c]));", "sum = fs_mad(wd, dn[c], sum);"),
    ("sum = fs_add(sum, fs_mul(wl, le[c]));", "sum = fs_mad(wl, le[c], sum);"),
    ("sum = fs_add(sum, fs_mul(wr, ri[c]));", "sum = fs_mad(wr, ri[c], sum);"),
    ("pw = fs_sub(fs_mul(pw, v_adj), 0.10889456f);", "pw = fs_mad(pw, v_adj, -0.10889456f);"),
    ("pw = fs_add(fs_mul(pw, v_adj), 0.107963754f);", "pw = fs_mad(pw, v_adj, 0.107963754f);"),
    ("pw = fs_add(fs_mul(pw, v_adj), 0.018092343f);", "pw = fs_mad(pw, v_adj, 0.018092343f);"),
    ("const float acc = fs_sub(fs_mul(pw, mul), 0.055f);", "const float acc = fs_mad(pw, mul, -0.055f);"),
    ("o[0] = fs_add(fs_add(fs_mul(m[0], a), fs_mul(m[1], b)), fs_mul(m[2], c));", "o[0] = fs_mad(m[2], c, fs_mad(m[1], b, fs_mul(m[0], a)));"),
    ("o[1] = fs_add(fs_add(fs_mul(m[3], a), fs_mul(m[4], b)), fs_mul(m[5], c));", "o[1] = fs_mad(m[5], c, fs_mad(m[4], b, fs_mul(m[3], a)));"),
    ("o[2] = fs_add(fs_add(fs_mul(m[6], a), fs_mul(m[7], b)), fs_mul(m[8], c));", "o[2] = fs_mad(m[8], c, fs_mad(m[7], b, fs_mul(m[6], a)));"),
]
DIST2 = re.compile(r"fs_add\(fs_add\(fs_mul\(s0, (fs_absdiff\([^)]*\))\), fs_mul\(s1, (fs_absdiff\([^)]*\))\)\), fs_mul\(s2, (fs_absdiff\([^)]*\))\)\)")


def contracted_header(text):
    assert HELPER_OLD in text
    text = text.replace(HELPER_OLD, HELPER_NEW, 1)
    for old, new in REPLACEMENTS:
        assert old in text, old
        text = text.replace(old, new, 1)
    text, n = DIST2.subn(r"fs_mad(s2, \3, fs_mad(s1, \2, fs_mul(s0, \1)))", text)
    assert n == 4
    return text


def prepare():
    """A private copy of csrc/, tests/emu/ and oracle/ whose filter_strip.cuh is the contracted one."""
    shutil.rmtree(BUILD, ignore_errors=True)
    os.makedirs(BUILD)
    shutil.copytree(os.path.join(ROOT, "jxl_oxide_b200", "csrc"), os.path.join(BUILD, "jxl_oxide_b200", "csrc"))
    shutil.copytree(os.path.join(ROOT, "include"), os.path.join(BUILD, "include"))
    shutil.copytree(os.path.join(ROOT, "oracle"), os.path.join(BUILD, "oracle"), ignore=shutil.ignore_patterns("_build", "_ref"))
    os.makedirs(os.path.join(BUILD, "tests"))
    shutil.copytree(os.path.join(ROOT, "tests", "emu"), os.path.join(BUILD, "tests", "emu"), ignore=shutil.ignore_patterns("_build"))
    hdr = os.path.join(BUILD, "jxl_oxide_b200", "csrc", "kernels", "filter_strip.cuh")
    with open(hdr) as f:
        text = f.read()
    with open(hdr, "w") as f:
        f.write(contracted_header(text))


def sass_size(obj, kernel_tag):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    keep, counts = False, collections.Counter()
    for line in out.splitlines():
        if "Function :" in line:
            keep = kernel_tag in line
            continue
        m = re.match(r"^\s*/\*[0-9a-f]{4,5}\*/\s*(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if keep and m:
            counts[m.group(1)] += 1
    return counts


def nvcc_object(csrc, out):
    from jxl_oxide_b200 import build as b
    flags = [f for f in b.NVCC_FLAGS]
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + flags + ["-x", "cu", "-c", os.path.join(csrc, "kernels", "filters_fused.cu"), "-o", out],
                          cwd=csrc, stderr=subprocess.DEVNULL)


def ulp_distance(a, b):
    """|a - b| in units in the last place of b (b: the oracle), computed on the ordered-integer image of the floats."""
    def ordered(x):
        i = x.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7fffffff), i)
    return np.abs(ordered(a) - ordered(b))


def main():
    import bench
    import oracle_lib
    prepare()
    subprocess.check_call(["make", "-s", "-C", os.path.join(BUILD, "tests", "emu")])
    fused = oracle_lib._load(os.path.join(BUILD, "tests", "emu", "_build", "libjxlemu.so"), None)
    frames = [("synthetic 2000x1500 d1.0 (seed 3)", bench.synth_frame(2000, 1500, 3)),
              ("synthetic 1000x600 d1.0 (seed 7)", bench.synth_frame(1000, 600, 7)),
              ("libjxl starrail.d1-e6 2560x1440 (one EPF step)", open(os.path.join(ROOT, "tests", "golden", "benchmark-data", "starrail.d1-e6.jxl"), "rb").read())]
    rows, total = [], collections.Counter()
    real_emu = oracle_lib.emu_lib
    for name, data in frames:
        want = oracle_lib.OracleImage(data, threads=8).frame(0)[0]
        oracle_lib.emu_lib = lambda: fused          # the contracted phases
        try:
            got = oracle_lib.OracleImage(data, threads=8, emu=True).frame(0)[0]
        finally:
            oracle_lib.emu_lib = real_emu
        h, w = want.shape[1:]
        inner = (slice(None), slice(64, h - 64), slice(64, w - 64))   # the strip kernel's territory (border tiles are untouched)
        d = ulp_distance(got[inner].ravel(), want[inner].ravel())
        hist = collections.Counter(np.minimum(d, 9).tolist())
        total.update(hist)
        absd = np.abs(got[inner].astype(np.float64) - want[inner].astype(np.float64))
        rows.append((name, d.size, hist, int(d.max()), float(absd.max())))
    # static instruction counts
    exact_obj = os.path.join(ROOT, "jxl_oxide_b200", "_obj", "kernels_filters_fused.cu.o")
    if not os.path.exists(exact_obj):
        from jxl_oxide_b200 import build as b
        b.build()
    fused_obj = os.path.join(BUILD, "filters_fused_contracted.o")
    nvcc_object(os.path.join(BUILD, "jxl_oxide_b200", "csrc"), fused_obj)
    tag = "strip_filter_kernelILi2ELi1"
    ce, cf = sass_size(exact_obj, tag), sass_size(fused_obj, tag)

    def fp(c):
        return c["FADD"] + c["FMUL"] + c["FFMA"]
    lines = ["# FMA contraction in the strip filter kernel: what it costs and what it saves (tools/fma_study.py)", "",
             "The shipped kernel rounds every product before adding, like the reference's generic path, and matches the oracle bit for",
             "bit. This experiment contracts the multiply-adds of Gaborish, the EPF distances / weights / weighted sums and the colour",
             "stage (matrix, sRGB polynomial) into fused multiply-adds in a COPY of `kernels/filter_strip.cuh` - divisions stay IEEE -",
             "runs the copy's phases on the host (the tests/emu harness; `std::fmaf` and the device's FFMA are the same correctly",
             "rounded operation) and compares the final samples with the oracle's. Interior of the frame only (64 samples in from",
             "every edge: the strip kernel's territory).", "",
             "| frame | samples | 0 ULP | 1 | 2 | 3 | 4-8 | >= 9 | max ULP | max abs diff |", "|---|---|---|---|---|---|---|---|---|---|"]
    for name, n, hist, mx, amax in rows:
        pct = lambda k: "%.2f %%" % (100.0 * hist.get(k, 0) / n)
        mid = sum(hist.get(k, 0) for k in range(4, 9))
        lines.append(f"| {name} | {n} | {pct(0)} | {pct(1)} | {pct(2)} | {pct(3)} | {100.0 * mid / n:.2f} % | {pct(9)} | {mx} | {amax:.3g} |")
    lines += ["", "Large ULP counts sit on samples near zero (a difference of 1e-7 is many units in the last place of 1e-6): the last column",
              "is the largest absolute difference, to be read against sample values in [0, 1].", "",
              f"Static size of `strip_filter_kernel<2, sRGB>` (sm_100a): {sum(ce.values())} instructions as shipped "
              f"(FADD {ce['FADD']}, FMUL {ce['FMUL']}, FFMA {ce['FFMA']}: {fp(ce)} fp32 arithmetic), {sum(cf.values())} contracted "
              f"(FADD {cf['FADD']}, FMUL {cf['FMUL']}, FFMA {cf['FFMA']}: {fp(cf)}): {100.0 * (1 - sum(cf.values()) / sum(ce.values())):.0f} % fewer "
              "instructions in a kernel that is bound by instruction issue (ncu: 73 % of the issue slots). The contracted build is not",
              "within 1 ULP of the reference everywhere (table above), so it does not meet north_star's parity bar as it stands and is",
              "not offered as a run-time option; the shipped kernels keep bit equality."]
    out = os.path.join(ROOT, "profiles", "r02_fma_study.md")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
