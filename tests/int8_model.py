"""Integer model of the INT8 variance engine (bosip.jl_b200/csrc/vargemm_i8.cu + the slice writers in kstar.cu /
i8_slice_linv_kernel), in NumPy int64 / Python integers.  Test infrastructure: it restates OUR kernel's arithmetic (not the
reference's) so that the error bound of the byte-slicing scheme is pinned on the CPU, independent of a GPU run.

Scheme (Ozaki-style error-free splitting, base 256, S slices per operand):
    k(x_i, X_k)          = IA[i,k] * 2^-(8S-1),        IA = rn(k * 2^(8S-1))            (unsigned, kernel values in [0, 1])
    alpha^2 Linv[n,k]    = IB[n,k] * 2^(f_n-(8S-2)),   IB = rn(v * 2^(8S-2-f_n))        (signed, f_n: per-row power-of-two scale)
    A_a = byte (S-1-a) of IA (unsigned), B_b = balanced signed digit (S-1-b) of IB, a = b = 0 most significant
    D_g = sum_{a+b=g} A_a . B_b^T  for g < S (exact int32 on the tensor cores); pairs with a + b >= S are dropped
    v[i,n] = 2^(f_n-13) * sum_g 2^-8g D_g ,   q[i] = sum_n v[i,n]^2
Default of the library (slices = 0): S = 7 slices of alpha^2 L^-1 against SA = 6 slices of K* (IA = rn(k * 2^47)): rows a = 0 .. 5 of the
slice-pair triangle, 27 pairs -- the 28th, (6, 0), would multiply the lowest byte of K* with the MOST significant byte of L^-1.
"""
import math

import numpy as np


def row_exponents(B):
    """f_n = smallest exponent with max_k |B[n,k]| <= 2^f_n (i8_rowscale_kernel)."""
    f = np.zeros(B.shape[0], dtype=np.int64)
    for n, mx in enumerate(np.max(np.abs(B), axis=1)):
        if mx > 0.0:
            m, e = math.frexp(mx)          # mx = m 2^e, m in [0.5, 1)
            f[n] = e - 1 if m == 0.5 else e
    return f


def slice_a(Kstar, S):
    """K* (M x N, values in [0, 1]) -> S unsigned byte planes [S][M][N], most significant first (kstar.cu, slice mode)."""
    IA = np.rint(np.ldexp(Kstar, 8 * S - 1)).astype(np.uint64)        # exact for 8S-1 <= 63
    return [((IA >> np.uint64(8 * (S - 1 - a))) & np.uint64(0xFF)).astype(np.int64) for a in range(S)], IA


def slice_b(B, f, S):
    """alpha^2 L^{-1} (N x N) -> S balanced signed byte planes, most significant first (i8_slice_linv_kernel)."""
    I = np.rint(np.ldexp(B, (8 * S - 2 - f)[:, None])).astype(np.int64)
    bias = sum(0x80 << (8 * b) for b in range(S - 1))
    u = (I + bias).astype(np.uint64)
    planes = []
    for b in range(S):
        s = S - 1 - b                                                   # digit index, 0 = least significant
        byte = ((u >> np.uint64(8 * s)) & np.uint64(0xFF)).astype(np.int64)
        if s < S - 1:
            byte = byte - 128                                           # (byte ^ 0x80) read as int8
        else:
            byte = np.where(byte >= 128, byte - 256, byte)              # the top byte is signed as it stands
        planes.append(byte)
    # the digits reproduce I exactly
    assert np.array_equal(sum(planes[b] * (1 << (8 * (S - 1 - b))) for b in range(S)), I)
    return planes, I


def engine_q(Kstar, B, S=7, SA=None):
    """q[i] = sum_n v[i,n]^2 as the engine computes it; also returns the int32 bound check and the exact integer dot products.
    SA: byte slices of K* (default S; the library's default configuration is S = 7, SA = 6)."""
    SA = S if SA is None else SA
    M, N = Kstar.shape
    f = row_exponents(B)
    A, IA = slice_a(Kstar, SA)          # slice a carries the weight 2^(-8a-7) whatever SA is
    Bp, IB = slice_b(B, f, S)
    D = []
    for g in range(S):
        Dg = np.zeros((M, N), dtype=np.int64)
        for a in range(min(g + 1, SA)):
            Dg += A[a] @ Bp[g - a].T
        assert np.max(np.abs(Dg)) < 2 ** 31                            # the TMEM accumulators are int32
        D.append(Dg)
    if S <= 7:       # one int64 per column: the four low-order groups exactly, rounded by R bits, the rest on top
        R = 11 if S == 7 else 8
        d = lambda k: D[S - 1 - k]                                     # k = 0: least significant group
        acc = d(0) + d(1) * (1 << 8) + d(2) * (1 << 16) + d(3) * (1 << 24)
        acc = (acc + (1 << (R - 1))) >> R
        for k in range(4, S):
            acc = acc + d(k) * (1 << (8 * k - R))
        assert np.max(np.abs(acc)) < 2 ** 63
        x = acc.astype(np.float64) * np.ldexp(1.0, f - 13 + R - 8 * (S - 1))[None, :]
    else:            # S = 8: FP64 Horner chain from the least significant group, fma(w, 2^-8, D_g)
        w = D[S - 1].astype(np.float64)
        for g in range(S - 2, -1, -1):
            w = w * 0.00390625 + D[g].astype(np.float64)
        x = w * np.ldexp(1.0, f - 13)[None, :]
    q = np.zeros(M)
    for n in range(N):                                                  # fma(x, x, rs), column by column
        q = q + x[:, n] * x[:, n]
    return q, x, (IA, IB, f)
