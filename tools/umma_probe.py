"""Hardware probe: which UMMA descriptor / TMA layout semantics hold on B200.

Run on the GPU box:  python tools/umma_probe.py
Each test prints PASS/FAIL; the kernels in cmfpy_b200/csrc/tc_*.cuh use only
layouts that pass here (results recorded in DESIGN.md).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libumma_probe.so"))
lib.probe_umma.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_uint, C.c_uint, C.c_int,
                           C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.probe_tma.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_void_p, C.c_int, C.c_void_p]

SW_NONE, SW128, SW64, SW32 = 0, 2, 4, 6
rng = np.random.default_rng(0)


def rnd(shape):
    """small values exactly representable in TF32; sums exact in fp32"""
    return (rng.integers(0, 16, size=shape) / 8.0).astype(np.float32)


def desc(off, lbo, sbo, layout, base_offset=0):
    d = (off >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= (layout & 7) << 61
    return d


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def swz128(byte):
    """absolute-address 128B swizzle: 16B-chunk index ^= (address >> 7) & 7"""
    return byte ^ (((byte >> 7) & 7) << 4)


class Image:
    def __init__(self, nbytes):
        self.buf = np.zeros(nbytes // 4, dtype=np.float32)

    def put(self, byte, v):
        self.buf[byte // 4] = v


def k_major_sw128(img, off, mat):
    """mat: R x 32 floats, rows of 128 B, K-major swizzle-128B atoms (8 rows x 128 B)"""
    R, Cc = mat.shape
    assert Cc == 32
    for r in range(R):
        for c in range(32):
            img.put(swz128(off + r * 128 + c * 4), mat[r, c])


def k_major_panel(img, off, mat, rows_total):
    """mat: R x Kc floats -> P[kc][row][4], no swizzle"""
    R, Kc = mat.shape
    for r in range(R):
        for c in range(Kc):
            img.put(off + ((c // 4) * rows_total + r) * 16 + (c % 4) * 4, mat[r, c])


def mn_major_sw128(img, off, mat, region_bytes):
    """mat: M x Kk (element (m, kk)); regions of 32 m: [kk][32 m] rows of 128 B, swizzled"""
    M, Kk = mat.shape
    for m in range(M):
        for kk in range(Kk):
            img.put(swz128(off + (m // 32) * region_bytes + kk * 128 + (m % 32) * 4), mat[m, kk])


def swz128_32(byte):
    """SWIZZLE_128B_BASE32B: 32B-chunk index ^= (address >> 7) & 3"""
    return byte ^ (((byte >> 7) & 3) << 5)


def timing(name, img, ad, bd, a_adv, b_adv, nsteps, idsc, N, repeat=512):
    out = np.zeros((128, N), dtype=np.float32)
    st = np.zeros(3, dtype=np.int32)
    lib.probe_umma(img.buf.ctypes.data, img.buf.nbytes, ad, bd, a_adv, b_adv, nsteps, idsc, N, 0, repeat,
                   out.ctypes.data, st.ctypes.data)
    print("%-58s %.1f cycles/MMA (%d MMAs)" % (name, st[2] / float(repeat * nsteps), repeat * nsteps))
    sys.stdout.flush()


def run(name, img, ad, bd, a_adv, b_adv, nsteps, idsc, N, expect, fix=0):
    out = np.zeros((128, N), dtype=np.float32)
    st = np.zeros(3, dtype=np.int32)
    rc = lib.probe_umma(img.buf.ctypes.data, img.buf.nbytes, ad, bd, a_adv, b_adv, nsteps, idsc, N, fix, 1,
                        out.ctypes.data, st.ctypes.data)
    ok = rc == 0 and st[0] == 1 and np.array_equal(out, expect.astype(np.float32))
    err = float(np.abs(out - expect).max()) if rc == 0 else -1
    print("%-58s %s  (status %d, smem base 0x%x, max|err| %.3g)" % (name, "PASS" if ok else "FAIL", st[0], st[1], err))
    sys.stdout.flush()
    return ok


results = {}

# ---- T1: canonical K-major SW128 for A and B ------------------------------------
A = rnd((128, 32)); B = rnd((64, 32))
img = Image(32768)
k_major_sw128(img, 0, A); k_major_sw128(img, 16384, B)
results["T1 A,B K-major SW128 canonical"] = run(
    "T1 A,B K-major SW128 canonical (N=64)", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128),
    2, 2, 4, idesc(128, 64, 0, 0), 64, A.astype(np.float64) @ B.astype(np.float64).T)

B256 = rnd((256, 32))
img = Image(16384 + 32768)
k_major_sw128(img, 0, A); k_major_sw128(img, 16384, B256)
results["T1b N=256"] = run("T1b same, N=256", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128),
                           2, 2, 4, idesc(128, 256, 0, 0), 256, A.astype(np.float64) @ B256.astype(np.float64).T)

# ---- T2: B as K-major no-swizzle panel with a row shift (recon / H-terms windows) --
ROWS = 256 + 24
Bwin = rnd((ROWS, 32))
for s in (0, 1, 5, 19):
    img = Image(16384 + ROWS * 128)
    k_major_sw128(img, 0, A); k_major_panel(img, 16384, Bwin, ROWS)
    exp = A.astype(np.float64) @ Bwin[s:s + 256].astype(np.float64).T
    results["T2 shift %d" % s] = run(
        "T2 B K-major INTERLEAVE panel, row shift %d, N=256" % s, img, desc(0, 16, 1024, SW128),
        desc(16384 + s * 16, ROWS * 16, 128, SW_NONE), 2, 2 * ROWS, 4, idesc(128, 256, 0, 0), 256, exp)

# ---- T2c: A as K-major no-swizzle panel with row shift -----------------------------
AROWS = 128 + 24
Awin = rnd((AROWS, 32))
for s in (0, 7):
    img = Image(20480 + 32768)
    k_major_panel(img, 0, Awin, AROWS); k_major_sw128(img, 20480, B256)
    exp = Awin[s:s + 128].astype(np.float64) @ B256.astype(np.float64).T
    results["T2c shift %d" % s] = run(
        "T2c A K-major INTERLEAVE panel, row shift %d" % s, img, desc(s * 16, AROWS * 16, 128, SW_NONE),
        desc(20480, 16, 1024, SW128), 2 * AROWS, 2, 4, idesc(128, 256, 0, 0), 256, exp)

# ---- T3: A MN-major, SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands) ----
# A element (m, kk): 4 regions of 32 m, each [kk rows][128 B]; 32B-granule swizzle; 4 MMAs of 8 kk
SW128_32B = 1


def mn_major_sw128_32(img, off, mat, region_bytes):
    M, Kk = mat.shape
    for m in range(M):
        for kk in range(Kk):
            img.put(swz128_32(off + (m // 32) * region_bytes + kk * 128 + (m % 32) * 4), mat[m, kk])


Amn = rnd((128, 32))
REG = 32 * 128
img = Image(4 * REG + 32768)
mn_major_sw128_32(img, 0, Amn, REG); k_major_sw128(img, 4 * REG, B256)
exp = Amn.astype(np.float64) @ B256.astype(np.float64).T
for (lbo, sbo) in ((REG, 512), (512, REG), (REG, 1024), (1024, REG)):
    results["T3 %d %d" % (lbo, sbo)] = run(
        "T3 A MN-major SW128_32B LBO=%d SBO=%d, B K-major SW128" % (lbo, sbo), img,
        desc(0, lbo, sbo, SW128_32B), desc(4 * REG, 16, 1024, SW128), 64, 2, 4, idesc(128, 256, 1, 0), 256, exp)
img_t3 = img

# ---- T5: B MN-major SW128_32B, overlapping atoms: 8 lags x 32 k, next lag = next row -------
HR = 32 + 8 + 16
Hw = rnd((HR, 32))
for r0 in (0, 4, 3):
    img = Image(16384 + 8192)
    k_major_sw128(img, 0, A)
    for r in range(HR):
        for c in range(32):
            img.put(swz128_32(16384 + r * 128 + c * 4), Hw[r, c])
    exp = np.zeros((128, 256))
    for a in range(8):
        for k in range(32):
            exp[:, a * 32 + k] = A.astype(np.float64) @ Hw[r0 + a:r0 + a + 32, k].astype(np.float64)
    for (lbo, sbo) in ((128, 512), (512, 128)):
        results["T5 r0=%d %d %d" % (r0, lbo, sbo)] = run(
            "T5 B MN-major SW128_32B overlapping atoms r0=%d LBO=%d SBO=%d" % (r0, lbo, sbo),
            img, desc(0, 16, 1024, SW128), desc(16384 + r0 * 128, lbo, sbo, SW128_32B), 2, 64, 4,
            idesc(128, 256, 0, 1), 256, exp)
img_t5 = img

# ---- T6: B K-major SW128 with a row shift through the start address ----------------
Bw = rnd((256 + 24, 32))
for s in (0, 8, 1, 5):
    img = Image(16384 + (256 + 24) * 128)
    k_major_sw128(img, 0, A); k_major_sw128(img, 16384, Bw)
    exp = A.astype(np.float64) @ Bw[s:s + 256].astype(np.float64).T
    for fix in (0,):
        results["T6 s=%d fix=%d" % (s, fix)] = run(
            "T6 B K-major SW128, row shift %d via start address, base_offset %s" % (s, "computed" if fix else "0"),
            img, desc(0, 16, 1024, SW128), desc(16384 + s * 128, 16, 1024, SW128), 2, 2, 4,
            idesc(128, 256, 0, 0), 256, exp, fix)

# ---- T7: TMA layouts ------------------------------------------------------------------
G = rng.random((64, 64)).astype(np.float32)


def tma(name, box_cols, box_rows, swz, c0, c1, dst_off, expect_fn, dump_bytes=16384):
    dump = np.zeros(dump_bytes, dtype=np.uint8)
    st = np.zeros(2, dtype=np.int32)
    rc = lib.probe_tma(G.ctypes.data, 64, 64, box_cols, box_rows, swz, c0, c1, dst_off, dump.ctypes.data, dump_bytes,
                       st.ctypes.data)
    got = dump.view(np.float32)
    exp = np.full(dump_bytes // 4, np.nan, dtype=np.float32)
    expect_fn(exp)
    mask = ~np.isnan(exp)
    ok = rc == 0 and st[0] == 1 and np.array_equal(got[mask], exp[mask])
    print("%-58s %s (status %d)" % (name, "PASS" if ok else "FAIL", st[0]))
    sys.stdout.flush()
    return ok


def exp_sw128(dst_off, c0, c1, rows):
    def f(exp):
        for r in range(rows):
            for c in range(32):
                exp[swz128(dst_off + r * 128 + c * 4) // 4] = G[c1 + r, c0 + c]
    return f


def exp_panel(dst_off, c0, c1, rows):
    def f(exp):
        for r in range(rows):
            for c in range(4):
                exp[(dst_off + r * 16 + c * 4) // 4] = G[c1 + r, c0 + c]
    return f


results["T7a"] = tma("T7a TMA box 32x40 SWIZZLE_128B == absolute-XOR layout", 32, 40, 3, 32, 5, 0, exp_sw128(0, 32, 5, 40))
results["T7b"] = tma("T7b TMA box 4x48 no swizzle == dense 16B rows", 4, 48, 0, 8, 3, 256, exp_panel(256, 8, 3, 48))
def exp_sw128_32(dst_off, c0, c1, rows):
    def f(exp):
        for r in range(rows):
            for c in range(32):
                exp[swz128_32(dst_off + r * 128 + c * 4) // 4] = G[c1 + r, c0 + c]
    return f


results["T7d"] = tma("T7d TMA box 32x40 SWIZZLE_128B_ATOM_32B == 32B-granule XOR", 32, 40, 4, 32, 5, 0, exp_sw128_32(0, 32, 5, 40))
results["T7e"] = tma("T7e TMA SW128_ATOM_32B to +5*128 destination", 32, 16, 4, 0, 0, 640, exp_sw128_32(640, 0, 0, 16))
results["T7c"] = tma("T7c TMA SW128 to a 1024B-aligned+3*128 destination", 32, 16, 3, 0, 0, 384, exp_sw128(384, 0, 0, 16))

# ---- timing: cycles per M=128 x N=256 x K=8 MMA for the operand layouts in play ---------
img = Image(16384 + 32768)
k_major_sw128(img, 0, A); k_major_sw128(img, 16384, B256)
timing("time: A K-SW128, B K-SW128 (N=256)", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128), 2, 2, 4, idesc(128, 256, 0, 0), 256)
timing("time: A K-SW128, B K-SW128 (N=128)", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128), 2, 2, 4, idesc(128, 128, 0, 0), 128)
timing("time: A K-SW128, B K-SW128 (N=64)", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128), 2, 2, 4, idesc(128, 64, 0, 0), 64)
timing("time: A K-SW128, B K-SW128 (N=32)", img, desc(0, 16, 1024, SW128), desc(16384, 16, 1024, SW128), 2, 2, 4, idesc(128, 32, 0, 0), 32)
img = Image(16384 + ROWS * 128)
k_major_sw128(img, 0, A); k_major_panel(img, 16384, Bwin, ROWS)
timing("time: A K-SW128, B K-INTERLEAVE panel shift 5 (N=256)", img, desc(0, 16, 1024, SW128),
       desc(16384 + 5 * 16, ROWS * 16, 128, SW_NONE), 2, 2 * ROWS, 4, idesc(128, 256, 0, 0), 256)
img = Image(20480 + 32768)
k_major_panel(img, 0, Awin, AROWS); k_major_sw128(img, 20480, B256)
timing("time: A K-INTERLEAVE panel, B K-SW128 (N=256)", img, desc(7 * 16, AROWS * 16, 128, SW_NONE),
       desc(20480, 16, 1024, SW128), 2 * AROWS, 2, 4, idesc(128, 256, 0, 0), 256)
timing("time: A MN-SW128_32B, B K-SW128 (N=256)", img_t3, desc(0, REG, 512, SW128_32B), desc(4 * REG, 16, 1024, SW128),
       64, 2, 4, idesc(128, 256, 1, 0), 256)
timing("time: A K-SW128, B MN-SW128_32B overlapping (N=256)", img_t5, desc(0, 16, 1024, SW128),
       desc(16384 + 3 * 128, 128, 512, SW128_32B), 2, 64, 4, idesc(128, 256, 0, 1), 256)

# shifted K-major SW128 windows (what K1 and K3 read: a lag is +128 B on the start address)
img = Image(16384 + (256 + 24) * 128)
k_major_sw128(img, 0, A); k_major_sw128(img, 16384, Bw)
for sh in (0, 1, 4, 5, 8):
    timing("time: A K-SW128, B K-SW128 window shift %d (N=256)" % sh, img, desc(0, 16, 1024, SW128),
           desc(16384 + sh * 128, 16, 1024, SW128), 2, 2, 4, idesc(128, 256, 0, 0), 256)
img = Image(4 * REG + (256 + 24) * 128)
mn_major_sw128_32(img, 0, Amn, REG); k_major_sw128(img, 4 * REG, Bw)
for sh in (0, 5):
    timing("time: A MN-SW128_32B, B K-SW128 window shift %d (N=256)" % sh, img, desc(0, REG, 512, SW128_32B),
           desc(4 * REG + sh * 128, 16, 1024, SW128), 64, 2, 4, idesc(128, 256, 1, 0), 256)

print()
print("SUMMARY: %d/%d passed" % (sum(results.values()), len(results)))
for k, v in results.items():
    if not v:
        print("  failed:", k)
