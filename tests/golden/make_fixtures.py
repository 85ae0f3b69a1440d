"""Copies the reference's own test inputs / golden outputs used by this repo's tests.

Run in the build container (where /root/reference exists); the GPU box only sees the copies.
Sources (read-only, data files — no source code is copied):
  crates/jxl-oxide-tests/decode/<case>/{input.jxl,output.buf.zst}      (tests/decode/mod.rs:122-130)
  crates/jxl-oxide-tests/conformance/testcases/<case>/{input.jxl,ref.png}
  crates/jxl-oxide-tests/decode/benchmark-data/*.jxl                    (benches/decode.rs:10-70)
"""
import os
import shutil

REF = "/root/reference/crates/jxl-oxide-tests"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (source dir, files)
    "grayalpha": ("decode/grayalpha", ["input.jxl", "output.buf.zst"]),
    "squeeze_edge": ("decode/squeeze_edge", ["input.jxl", "output.buf.zst"]),
    "issue_311": ("decode/issue_311", ["input.jxl", "output.buf.zst"]),
    "issue_24": ("decode/issue_24", ["input.jxl", "output.buf.zst"]),  # 1x1 VarDCT animation, 9 keyframes
    "minecraft_vardct_e7": ("decode/minecraft_vardct_e7", ["input.jxl"]),
    "opsin_inverse": ("conformance/testcases/opsin_inverse", ["input.jxl", "ref.png"]),
    "alpha_premultiplied": ("conformance/testcases/alpha_premultiplied", ["input.jxl"]),
    "alpha_triangles": ("conformance/testcases/alpha_triangles", ["input.jxl", "ref.png"]),
    "bicycles": ("conformance/testcases/bicycles", ["input.jxl"]),
    "lz77_flower": ("conformance/testcases/lz77_flower", ["input.jxl", "ref.png"]),
    "upsampling": ("conformance/testcases/upsampling", ["input.jxl", "ref.png"]),
    "noise": ("conformance/testcases/noise", ["input.jxl", "ref.png"]),
    "patches_lossless": ("conformance/testcases/patches_lossless", ["input.jxl", "ref.png"]),
    "bike": ("conformance/testcases/bike", ["input.jxl"]),
    "sunset_logo": ("conformance/testcases/sunset_logo", ["input.jxl"]),
    "grayscale_public_university": ("conformance/testcases/grayscale_public_university", ["input.jxl"]),
    "blendmodes": ("conformance/testcases/blendmodes", ["input.jxl", "ref.png"]),
    "animation_icos4d": ("conformance/testcases/animation_icos4d", ["input.jxl"]),
    "animation_spline": ("conformance/testcases/animation_spline", ["input.jxl"]),
    "bench_oriented_brg": ("conformance/testcases/bench_oriented_brg", ["input.jxl", "ref.png"]),
    "grayscale_jpeg": ("conformance/testcases/grayscale_jpeg", ["input.jxl"]),
    "cafe": ("conformance/testcases/cafe", ["input.jxl"]),
    "issue_425": ("decode/issue_425", ["input.jxl", "ref.jpg"]),
    "genshin_ycbcr_420": ("decode/genshin_ycbcr_420", ["input.jxl"]),  # 2560x1440 4:2:0 transcode, 60 groups
    "spot": ("conformance/testcases/spot", ["input.jxl"]),
    "grayscale": ("conformance/testcases/grayscale", ["input.jxl", "ref.png"]),
    # these three need a CMS to match their ref.png (tabulated-curve / chrm / CMYK ICC profiles): kept for the ICC
    # digests, the stream layout and GPU-vs-oracle parity
    "cmyk_layers": ("conformance/testcases/cmyk_layers", ["input.jxl"]),
    "patches": ("conformance/testcases/patches", ["input.jxl"]),
    "progressive": ("conformance/testcases/progressive", ["input.jxl"]),
    "lossless_pfm": ("conformance/testcases/lossless_pfm", ["input.jxl"]),
    "alpha_nonpremultiplied": ("conformance/testcases/alpha_nonpremultiplied", ["input.jxl", "ref.png"]),
    "animation_newtons_cradle": ("conformance/testcases/animation_newtons_cradle", ["input.jxl"]),
    "delta_palette": ("conformance/testcases/delta_palette", ["input.jxl"]),
}
BENCH = ["starrail.d1-e6.jxl", "nahida-motion.d1-e7.jxl", "srgb.d0-e1.jxl", "minecraft.d0-e6.jxl"]

for name, (src, files) in CASES.items():
    os.makedirs(os.path.join(HERE, name), exist_ok=True)
    for f in files:
        shutil.copyfile(os.path.join(REF, src, f), os.path.join(HERE, name, f))
os.makedirs(os.path.join(HERE, "benchmark-data"), exist_ok=True)
for f in BENCH:
    shutil.copyfile(os.path.join(REF, "decode/benchmark-data", f), os.path.join(HERE, "benchmark-data", f))
# bike's reference rendering is 6.7 MB: keep a 640 x 640 crop at (700, 900) (BIKE_CROP in the tests)
from PIL import Image
Image.open(os.path.join(REF, "conformance/testcases/bike/ref.png")).crop((700, 900, 1340, 1540)).save(
    os.path.join(HERE, "bike", "ref_crop_700_900.png"), optimize=True)
# sunset_logo's reference is 0.9 MB: a 512 x 512 crop at (200, 400) of the oriented image
Image.open(os.path.join(REF, "conformance/testcases/sunset_logo/ref.png")).crop((200, 400, 712, 912)).save(
    os.path.join(HERE, "sunset_logo", "ref_crop_200_400.png"), optimize=True)
Image.open(os.path.join(REF, "conformance/testcases/grayscale_public_university/ref.png")).crop((1000, 500, 1512, 1012)).save(
    os.path.join(HERE, "grayscale_public_university", "ref_crop_1000_500.png"), optimize=True)
# cafe (4:2:0 JPEG transcode): a 512 x 512 crop at (600, 800) and the 256 x 256 bottom-right corner
Image.open(os.path.join(REF, "conformance/testcases/cafe/ref.png")).crop((600, 800, 1112, 1312)).save(
    os.path.join(HERE, "cafe", "ref_crop_600_800.png"), optimize=True)
Image.open(os.path.join(REF, "conformance/testcases/cafe/ref.png")).crop((1024, 1344, 1280, 1600)).save(
    os.path.join(HERE, "cafe", "ref_crop_corner.png"), optimize=True)
# genshin_ycbcr_420: a 256 x 256 crop at (1000, 600) of the source JPEG as libjpeg decodes it (ref.jpg is 0.7 MB)
Image.open(os.path.join(REF, "decode/genshin_ycbcr_420/ref.jpg")).convert("RGB").crop((1000, 600, 1256, 856)).save(
    os.path.join(HERE, "genshin_ycbcr_420", "refjpg_crop_1000_600.png"), optimize=True)
# spot colours: a 300 x 300 crop at (150, 50) of the 8-bit rendering
Image.open(os.path.join(REF, "conformance/testcases/spot/ref.png")).crop((150, 50, 450, 350)).save(
    os.path.join(HERE, "spot", "ref_crop_150_50.png"), optimize=True)
Image.open(os.path.join(REF, "conformance/testcases/delta_palette/ref.png")).crop((100, 200, 400, 500)).save(
    os.path.join(HERE, "delta_palette", "ref_crop_100_200.png"), optimize=True)
# lossless_pfm: the float reference (3 MB) is kept as the digest of its top-down interleaved little-endian f32 samples
import hashlib
import numpy as np
_raw = open(os.path.join(REF, "conformance/testcases/lossless_pfm/ref.pfm"), "rb").read().split(b"\n", 3)
_w, _h = map(int, _raw[1].split())
_pix = np.frombuffer(_raw[3], dtype="<f4" if float(_raw[2]) < 0 else ">f4").reshape(_h, _w, 3)[::-1]
with open(os.path.join(HERE, "lossless_pfm", "ref_f32_sha256.txt"), "w") as _f:
    _f.write(hashlib.sha256(np.ascontiguousarray(_pix, dtype="<f4").tobytes()).hexdigest() + "\n")
Image.open(os.path.join(REF, "conformance/testcases/bicycles/ref.png")).crop((300, 100, 812, 612)).save(
    os.path.join(HERE, "bicycles", "ref_crop_300_100.png"), optimize=True)
Image.open(os.path.join(REF, "conformance/testcases/alpha_premultiplied/ref.png")).crop((256, 256, 768, 768)).save(
    os.path.join(HERE, "alpha_premultiplied", "ref_crop_256_256.png"), optimize=True)
_g = Image.open(os.path.join(REF, "conformance/testcases/animation_newtons_cradle/ref.gif"))
for _k in (0, 10):
    _g.seek(_k)
    _g.convert("RGBA").save(os.path.join(HERE, "animation_newtons_cradle", "ref_frame_%02d.png" % _k), optimize=True)
# three frames of the animation's reference APNG
_ap = Image.open(os.path.join(REF, "conformance/testcases/animation_icos4d/ref.apng"))
for _k in (0, 17, 47):
    _ap.seek(_k)
    _ap.convert("RGBA").save(os.path.join(HERE, "animation_icos4d", "ref_frame_%02d.png" % _k), optimize=True)
_ap = Image.open(os.path.join(REF, "conformance/testcases/animation_spline/ref.apng"))
for _k in (0, 23, 59):
    _ap.seek(_k)
    _ap.convert("RGB").save(os.path.join(HERE, "animation_spline", "ref_frame_%02d.png" % _k), optimize=True)
# the reference's own ICC test profiles (crates/jxl-color/src/icc/test-profiles, expectations in icc/parse.rs:566-671)
import glob
os.makedirs(os.path.join(HERE, "icc"), exist_ok=True)
for f in sorted(glob.glob("/root/reference/crates/jxl-color/src/icc/test-profiles/*.icc")):
    shutil.copyfile(f, os.path.join(HERE, "icc", os.path.basename(f)))
# malformed inputs found by the reference's fuzzers (crates/jxl-oxide-tests/tests/fuzz_findings): expectation =
# no crash, a clean error value (or a successful decode)
import glob
os.makedirs(os.path.join(HERE, "fuzz_findings"), exist_ok=True)
for f in sorted(glob.glob(os.path.join(REF, "tests/fuzz_findings/*.fuzz"))):
    shutil.copyfile(f, os.path.join(HERE, "fuzz_findings", os.path.basename(f)))
print("fixtures copied")
