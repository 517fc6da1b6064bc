"""Reverse-engineer UMMA operand addressing: which smem word is read for logical (row, k)?"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_gnn_ssl_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return C.c_void_p(t.data_ptr())
def run(aimg, bimg, albo, asbo, blbo, bsbo, amn, bmn, atype=0, btype=0, N=16):
    a = torch.tensor(aimg, dtype=torch.float32, device=dev); b = torch.tensor(bimg, dtype=torch.float32, device=dev)
    D = torch.full((128, N), float("nan"), device=dev)
    _lib.check(L.dcgru_tc_probe(p(a), a.numel() * 4, p(b), b.numel() * 4, albo, asbo, blbo, bsbo, amn, bmn, atype, btype, p(D), N, st), "probe")
    torch.cuda.synchronize()
    return D.cpu().numpy()
def b_onehot_kmajor():
    img = np.zeros(16 * 8, dtype=np.float32)     # [kg(2)][n(16)][4]: LBO 256, SBO 128
    for n in range(8):
        img[(n // 4) * 64 + n * 4 + n % 4] = 1.0
    return img
def a_onehot_kmajor():
    img = np.zeros(128 * 8, dtype=np.float32)    # [kg(2)][m(128)][4]: LBO = 2048, SBO = 128
    for m in range(8):
        img[(m // 4) * 512 + m * 4 + m % 4] = 1.0
    return img
# values must be exact in tf32 (10-bit mantissa): use word index for < 2048 words
aidx = np.arange(2048, dtype=np.float32)
print("== A MN-major, layout types")
for (atype, albo, asbo) in [(1, 1024, 512)]:
    D = run(aimg=aidx, bimg=b_onehot_kmajor(), albo=albo, asbo=asbo, blbo=256, bsbo=128, amn=1, bmn=0, atype=atype)
    print(f"A probe: MN-major type={atype} lbo={albo} sbo={asbo}: word index read for (m, k=0..7):")
    for m in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 24, 31, 32, 33, 63, 64, 96, 127):
        print(f"   m={m:3d}:", [int(D[m, k]) if np.isfinite(D[m, k]) else None for k in range(8)])
print("== B MN-major, layout types")
bidx = np.arange(2048, dtype=np.float32)
for (btype, blbo, bsbo) in [(1, 1024, 512)]:
    D = run(aimg=a_onehot_kmajor(), bimg=bidx, albo=2048, asbo=128, blbo=blbo, bsbo=bsbo, amn=0, bmn=1, btype=btype, N=64)
    print(f"B probe: MN-major type={btype} lbo={blbo} sbo={bsbo}: word index read for (n, k=0..7):")
    for n in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 24, 31, 32, 33, 63):
        print(f"   n={n:3d}:", [int(D[k, n]) if np.isfinite(D[k, n]) else None for k in range(8)])
print("== A K-major, 32B swizzle (type 6): expect (m/8)*64 + (m%8)*8 + ((k/4)^((m%8)>>2))*4 + k%4")
D = run(aimg=aidx, bimg=b_onehot_kmajor(), albo=16, asbo=256, blbo=256, bsbo=128, amn=0, bmn=0, atype=6)
ok = True
for m in range(128):
    for k in range(8):
        exp = (m // 8) * 64 + (m % 8) * 8 + ((k // 4) ^ ((m % 8) >> 2)) * 4 + k % 4
        ok &= int(D[m, k]) == exp
print("K-major SW32 layout as expected:", ok)
for m in (0, 1, 3, 4, 5, 7, 8, 12):
    print(f"   m={m:3d}:", [int(D[m, k]) for k in range(8)])
