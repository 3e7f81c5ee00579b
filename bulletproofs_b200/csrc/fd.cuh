// Field arithmetic mod p = 2^255 - 19 on the FP64 pipe (DFMA / DADD), for the long squaring chains only.
//
// Why: every field kernel of the engine is bound by the integer-multiply pipe (one IMAD.WIDE per 4 cycles per SM sub-partition,
// DESIGN.md §3), and the largest single consumer is the 249-squaring ladder z^(2^250-1) inside every point decompression
// (curve25519-dalek's FieldElement::pow22523 / invert behind CompressedRistretto::decompress, which the reference calls for
// every proof point: /root/reference/src/range_proof/mod.rs:414-433).  The FP64 pipe of sm_100 is a separate unit that the
// integer kernels leave idle; a ladder that runs there costs the multiply pipe nothing.
//
// Representation: six limbs in mixed radix 2^42.5 (widths 43,42,43,42,43,42 at bit positions 0,43,85,128,170,213; 6 x 42.5 = 255,
// so the wrap-around factor is exactly 19), each limb a non-negative integer held in a double.  "Loosely normalised": limb k is
// below 2^width_k + 2^14.
//
// Exact products from FMAs (the classic double-FMA split): for integers x, y with x*y < 2^94 and a constant C = 2^(52+w),
//     hi = fma_rz(x, y, C)  =  C + 2^w * floor(x*y / 2^w)          (the ulp in [C, 2C) is 2^w; round-toward-zero truncates)
//     lo = fma(x, y, C - hi) =  x*y mod 2^w                         (exact: C - hi is a multiple of 2^w, the result is < 2^w)
// A column's high parts accumulate for free by chaining (hi_next = fma_rz(x', y', hi)), its low parts with one DADD each; every
// intermediate value is an integer below 2^53 or a multiple of 2^w with at most 52 significant bits, i.e. exactly representable —
// the results are bit-exact integers, which tests/test_host_emul.py checks against Python integers with the same code compiled
// for the host (IEEE-754 FMA under FE_TOWARDZERO) and tests/test_gpu_parity.py checks on the device.
//
// Cost: squaring = 120 FP64-pipe instructions (multiplication 177) against 44 (72) wide integer multiplies for the 8x32-bit form of
// fe.cuh.  It is not faster in isolation; it moves the ladder to the other pipe.
#pragma once
#include "fe.cuh"
#include <math.h>
#include <string.h>

struct fd { double v[6]; };

#if defined(__CUDA_ARCH__)
#define FD_FMA_RZ(a, b, c) __fma_rz((a), (b), (c))
#define FD_ADD_RZ(a, b) __dadd_rz((a), (b))
#define FD_FMA(a, b, c) __fma_rn((a), (b), (c))          // only where the result is exactly representable
#define FD_ADD(a, b) __dadd_rn((a), (b))                  // only where the result is exactly representable
#define FD_MUL(a, b) __dmul_rn((a), (b))                  // only where the result is exactly representable
#else
// host build (tests/host_emul only): IEEE-754 double FMA / ADD under FE_TOWARDZERO are what DFMA.RZ / DADD.RZ compute; the other
// operations run in the default mode, as on the device.  The volatiles pin the operation between the two mode switches.
#include <fenv.h>
static inline double fd_host_fma_rz(double a, double b, double c) {
    volatile double va = a, vb = b, vc = c; int old = fegetround(); fesetround(FE_TOWARDZERO);
    volatile double r = fma(va, vb, vc); fesetround(old); return r;
}
static inline double fd_host_add_rz(double a, double b) {
    volatile double va = a, vb = b; int old = fegetround(); fesetround(FE_TOWARDZERO);
    volatile double r = va + vb; fesetround(old); return r;
}
#define FD_FMA_RZ(a, b, c) fd_host_fma_rz((a), (b), (c))
#define FD_ADD_RZ(a, b) fd_host_add_rz((a), (b))
#define FD_FMA(a, b, c) fma((a), (b), (c))
#define FD_ADD(a, b) ((a) + (b))
#define FD_MUL(a, b) ((a) * (b))
#endif

// -DFD_FINISH_INT (experiment, benchmarks/fp64_variants.sh): assemble the columns and carry with 64-bit integer ALU operations on the
// doubles' mantissa fields instead of 31 FP64 operations; the low sums then carry a bias of +2^52 so that their mantissa is the sum.
#ifdef FD_FINISH_INT
#define FD_BIAS0(C) ((C) + 0x1p52)
#define FD_BIASK(C) ((C) + 0x1p52)
#else
#define FD_BIAS0(C) (C)
#define FD_BIASK(C) ((C) - 0x1p52)
#endif
#define FD_C43 0x1p95            // split constant of a 43-bit column: 2^(52+43)
#define FD_C42 0x1p94
#define FD_B52 0x1p52
#define FD_I43 0x1p-43
#define FD_I42 0x1p-42

BP_HD double fd_from_u64(uint64_t x) {          // x < 2^52, exact
    uint64_t bits = 0x4330000000000000ull | x;
#if defined(__CUDA_ARCH__)
    return FD_ADD(__longlong_as_double((long long)bits), -FD_B52);
#else
    double d; memcpy(&d, &bits, 8); return d - FD_B52;
#endif
}
BP_HD uint64_t fd_to_u64(double x) {            // x a non-negative integer < 2^52
    double s = FD_ADD(x, FD_B52);
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(s) & 0xfffffffffffffull;
#else
    uint64_t bits; memcpy(&bits, &s, 8); return bits & 0xfffffffffffffull;
#endif
}

// bits [pos, pos+n) of the 256-bit little-endian word array
BP_HD uint64_t fd_bits(const uint32_t w[8], int pos, int n) {
    int idx = pos >> 5, sh = pos & 31;
    uint64_t x = ((uint64_t)w[idx] | ((idx + 1 < 8) ? (uint64_t)w[idx + 1] << 32 : 0)) >> sh;
    if (sh + n > 64 && idx + 2 < 8) x |= (uint64_t)w[idx + 2] << (64 - sh);
    return x & ((1ull << n) - 1);
}

// any representative below 2^256 -> strictly normalised limbs (limb 5 may reach 2^42 when the folded value is 2^255 .. 2^255+18)
BP_HD fd fd_from_fe(const fe &a) {
    uint32_t w[8];
    uint64_t c = (uint64_t)(a.v[7] >> 31) * 19u;
    for (int i = 0; i < 8; i++) { c += (i == 7) ? (a.v[7] & 0x7fffffffu) : a.v[i]; w[i] = (uint32_t)c; c >>= 32; }
    fd r;
    r.v[0] = fd_from_u64(fd_bits(w, 0, 43));
    r.v[1] = fd_from_u64(fd_bits(w, 43, 42));
    r.v[2] = fd_from_u64(fd_bits(w, 85, 43));
    r.v[3] = fd_from_u64(fd_bits(w, 128, 42));
    r.v[4] = fd_from_u64(fd_bits(w, 170, 43));
    r.v[5] = fd_from_u64(fd_bits(w, 213, 43));
    return r;
}
// loosely normalised limbs -> a representative below 2^256
BP_HD fe fd_to_fe(const fd &a) {
    const int pos[6] = {0, 43, 85, 128, 170, 213};
    uint64_t t[9];
    for (int i = 0; i < 9; i++) t[i] = 0;
    for (int k = 0; k < 6; k++) {
        uint64_t u = fd_to_u64(a.v[k]);
        int idx = pos[k] >> 5, sh = pos[k] & 31;
        uint64_t lo = u << sh, hi = sh ? (u >> (64 - sh)) : 0;
        t[idx] += lo & 0xffffffffull; t[idx + 1] += lo >> 32; t[idx + 2] += hi;
    }
    fe r; uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += t[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return r;
}

// one product into a column: acc is the chained high part (C + 2^w * sum of floors), ls the running sum of the low parts
BP_HD void fd_col_first(double &acc, double &ls, double x, double y, double C, double C_minus_bias) {
    acc = FD_FMA_RZ(x, y, C);
    ls = FD_FMA(x, y, FD_ADD(C_minus_bias, -acc));
}
BP_HD void fd_col_next(double &acc, double &ls, double x, double y) {
    double h = FD_FMA_RZ(x, y, acc);
    ls = FD_ADD(ls, FD_FMA(x, y, FD_ADD(acc, -h)));
    acc = h;
}
// columns -> loosely normalised limbs.  acc[k], ls[k] as above, ls[1..5] carry a bias of -2^52 that cancels the 2^52 of acc[k-1] * 2^-w.
#ifdef FD_FINISH_INT
BP_HD uint64_t fd_mant(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x) & 0xfffffffffffffull;
#else
    uint64_t b; memcpy(&b, &x, 8); return b & 0xfffffffffffffull;
#endif
}
BP_HD fd fd_finish(const double acc[6], const double ls[6]) {
    uint64_t H[6], V[6];
    for (int k = 0; k < 6; k++) { H[k] = fd_mant(acc[k]); V[k] = fd_mant(ls[k]); }
    V[0] += 19u * H[5];
    for (int k = 1; k < 6; k++) V[k] += H[k - 1];
    uint64_t c[6], r[6];
    for (int k = 0; k < 6; k++) { const int w = (k & 1) ? 42 : 43; c[k] = V[k] >> w; r[k] = V[k] & ((1ull << w) - 1); }
    fd o;
    o.v[0] = fd_from_u64(r[0] + 19u * c[5]);
    for (int k = 1; k < 6; k++) o.v[k] = fd_from_u64(r[k] + c[k - 1]);
    return o;
}
#else
BP_HD fd fd_finish(const double acc[6], const double ls[6]) {
    double V[6];
    V[0] = FD_FMA(FD_ADD(acc[5], -FD_C42), 19.0 * FD_I42, ls[0]);        // column 5's high part wraps to column 0 with factor 19
    V[1] = FD_FMA(acc[0], FD_I43, ls[1]);
    V[2] = FD_FMA(acc[1], FD_I42, ls[2]);
    V[3] = FD_FMA(acc[2], FD_I43, ls[3]);
    V[4] = FD_FMA(acc[3], FD_I42, ls[4]);
    V[5] = FD_FMA(acc[4], FD_I43, ls[5]);
    // one parallel carry pass: V[k] < 2^51.1, so every carry is below 2^9.1
    double t[6], r[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 6; k++) {
        const double C = (k & 1) ? FD_C42 : FD_C43;
        double h = FD_ADD_RZ(V[k], C);
        t[k] = FD_ADD(C, -h);                                             // -2^w * carry
        r[k] = FD_ADD(V[k], t[k]);
    }
    fd o;
    o.v[0] = FD_FMA(t[5], -19.0 * FD_I42, r[0]);
    o.v[1] = FD_FMA(t[0], -FD_I43, r[1]);
    o.v[2] = FD_FMA(t[1], -FD_I42, r[2]);
    o.v[3] = FD_FMA(t[2], -FD_I43, r[3]);
    o.v[4] = FD_FMA(t[3], -FD_I42, r[4]);
    o.v[5] = FD_FMA(t[4], -FD_I43, r[5]);
    return o;
}
#endif

// a^2: 21 products.  Cross terms carry 2, odd-odd pairs another 2 (their positions add up one bit past the column's), wrapped
// columns (i + j >= 6) carry 19.
BP_HD fd fd_sq(const fd &a) {
    const double a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4], a5 = a.v[5];
    const double d0 = FD_ADD(a0, a0), d1 = FD_ADD(a1, a1), d2 = FD_ADD(a2, a2), d3 = FD_ADD(a3, a3), d4 = FD_ADD(a4, a4), d5 = FD_ADD(a5, a5);
    const double q1 = FD_ADD(d1, d1), q3 = FD_ADD(d3, d3);
    const double n3 = FD_MUL(a3, 19.0), n4 = FD_MUL(a4, 19.0), n5 = FD_MUL(a5, 19.0);
    double acc[6], ls[6];
    fd_col_first(acc[0], ls[0], a0, a0, FD_C43, FD_BIAS0(FD_C43));
    fd_col_next(acc[0], ls[0], q1, n5); fd_col_next(acc[0], ls[0], d2, n4); fd_col_next(acc[0], ls[0], d3, n3);
    fd_col_first(acc[1], ls[1], d0, a1, FD_C42, FD_BIASK(FD_C42));
    fd_col_next(acc[1], ls[1], d2, n5); fd_col_next(acc[1], ls[1], d3, n4);
    fd_col_first(acc[2], ls[2], d0, a2, FD_C43, FD_BIASK(FD_C43));
    fd_col_next(acc[2], ls[2], d1, a1); fd_col_next(acc[2], ls[2], q3, n5); fd_col_next(acc[2], ls[2], a4, n4);
    fd_col_first(acc[3], ls[3], d0, a3, FD_C42, FD_BIASK(FD_C42));
    fd_col_next(acc[3], ls[3], d1, a2); fd_col_next(acc[3], ls[3], d4, n5);
    fd_col_first(acc[4], ls[4], d0, a4, FD_C43, FD_BIASK(FD_C43));
    fd_col_next(acc[4], ls[4], q1, a3); fd_col_next(acc[4], ls[4], a2, a2); fd_col_next(acc[4], ls[4], d5, n5);
    fd_col_first(acc[5], ls[5], d0, a5, FD_C42, FD_BIASK(FD_C42));
    fd_col_next(acc[5], ls[5], d1, a4); fd_col_next(acc[5], ls[5], d2, a3);
    return fd_finish(acc, ls);
}

// a * b: 36 products
BP_HDN fd fd_mul(const fd &a, const fd &b) {
    double ad[6], bn[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; i++) { ad[i] = (i & 1) ? FD_ADD(a.v[i], a.v[i]) : a.v[i]; bn[i] = FD_MUL(b.v[i], 19.0); }
    double acc[6], ls[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 6; k++) {
        const double C = (k & 1) ? FD_C42 : FD_C43;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 6; i++) {
            const int j = (k - i + 6) % 6;
            const double x = ((i & 1) && (j & 1)) ? ad[i] : a.v[i], y = (i + j >= 6) ? bn[j] : b.v[j];
            if (i == 0) fd_col_first(acc[k], ls[k], x, y, C, k == 0 ? FD_BIAS0(C) : FD_BIASK(C));
            else fd_col_next(acc[k], ls[k], x, y);
        }
    }
    return fd_finish(acc, ls);
}
BP_HDN fd fd_sqn(fd a, int n) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < n; i++) a = fd_sq(a);
    return a;
}

// z^(2^250-1) with z^11 on the side: the ladder of fe_pow_2_250_1 (fe.cuh) on the FP64 pipe
BP_HD fd fd_pow_2_250_1(const fd &z, fd &z11) {
    fd t0 = fd_sqn(z, 1);                        // 2
    fd t1 = fd_mul(z, fd_sqn(t0, 2));            // 9
    t0 = fd_mul(t0, t1);                         // 11
    z11 = t0;
    t1 = fd_mul(t1, fd_sqn(t0, 1));              // 2^5-1
    t1 = fd_mul(fd_sqn(t1, 5), t1);              // 2^10-1
    fd t2 = fd_mul(fd_sqn(t1, 10), t1);          // 2^20-1
    t2 = fd_mul(fd_sqn(t2, 20), t2);             // 2^40-1
    t1 = fd_mul(fd_sqn(t2, 10), t1);             // 2^50-1
    t2 = fd_mul(fd_sqn(t1, 50), t1);             // 2^100-1
    t2 = fd_mul(fd_sqn(t2, 100), t2);            // 2^200-1
    return fd_mul(fd_sqn(t2, 50), t1);           // 2^250-1
}
BP_HDN fe fe_invert_fd(const fe &z) { fd x = fd_from_fe(z), z11; fd t = fd_pow_2_250_1(x, z11); return fd_to_fe(fd_mul(fd_sqn(t, 5), z11)); }
BP_HDN fe fe_pow22523_fd(const fe &z) { fd x = fd_from_fe(z), z11; fd t = fd_pow_2_250_1(x, z11); return fd_to_fe(fd_mul(fd_sqn(t, 2), x)); }
