// Per-proof verification scalars of an aggregated range proof, computed on the device.
//
// Restates RangeProof::verify_multiple_with_rng up to (not including) the mega-MSM:
//   /root/reference/src/range_proof/mod.rs:368-419 (transcript replay, scalar assembly),
//   /root/reference/src/inner_product_proof.rs:198-253 (verification_scalars),
//   /root/reference/src/range_proof/mod.rs:587-593 (delta),
// and multiplies every scalar by a per-proof random weight rho so that a batch of proofs can be
// checked with one random-linear-combination MSM.
//
// The work is split in a sequential head (one thread: transcript, one field inversion) and a
// data-parallel tail (one index i of the N = n*m generator pairs per thread).  All scalars are in
// Montgomery form between the byte boundaries.
#pragma once
#include "sc.cuh"
#include "merlin.cuh"

// verdict codes, mirroring ProofError (/root/reference/src/errors.rs:12-54)
#define BP_PROOF_OK 0
#define BP_PROOF_VERIFICATION_ERROR 1
#define BP_PROOF_FORMAT_ERROR 2

#define BP_MAX_LG_N 20          // verify up to N = n*m = 2^20 generator pairs per proof (R1CS-sized)

struct rp_head {
    // challenges and derived values, Montgomery form
    sc y_inv, z, zz, x, w, c, rho;
    sc a, b, t_x, t_x_bl, e_bl;
    sc s0;                                  // allinv = prod u_j^-1
    sc u_sq[BP_MAX_LG_N], u_inv_sq[BP_MAX_LG_N];
    sc y_inv_pow2[BP_MAX_LG_N];             // y^-(2^b)
    sc basepoint_scalar, blinding_scalar;   // weighted by rho
    uint32_t status;
};

BP_HD bool bp_is_zero32(const uint8_t *p) { uint8_t z = 0; for (int i = 0; i < 32; i++) z |= p[i]; return z == 0; }

BP_HD sc rp_challenge(merlin_t &t, const char *label) { uint8_t buf[64]; merlin_challenge(t, label, buf, 64); return sc_mont_from_wide(buf); }
BP_HD void rp_append_scalar(merlin_t &t, const char *label, const uint8_t *b) { merlin_append(t, label, b, 32); }

// sum_{i<n} x^i for n a power of two (util.rs:240-256), Montgomery form
BP_HD sc rp_sum_of_powers_pow2(const sc &x, uint64_t n) {
    if (n == 1) return sc_mont_one();
    sc result = sc_add(sc_mont_one(), x), factor = x;
    for (uint64_t m = n; m > 2; m >>= 1) { factor = sc_mont_mul(factor, factor); result = sc_add(result, sc_mont_mul(factor, result)); }
    return result;
}

// Sequential head.  proof = 32*(9+2k) bytes, V = m*32 bytes, tstate = serialized transcript the caller
// passed in (state after Transcript::new(label) and any application messages), weights = 128 bytes
// of per-proof randomness (c from the first 64, rho from the last 64).
BP_HDN void rp_prep_head(rp_head &h, const uint8_t *proof, uint32_t k, const uint8_t *V, uint32_t n, uint32_t m,
                         const uint8_t *tstate, const uint8_t *weights) {
    h.status = BP_PROOF_OK;
    const uint8_t *A = proof, *S = proof + 32, *T1 = proof + 64, *T2 = proof + 96;
    const uint8_t *LR = proof + 224, *ab = proof + 224 + 64 * k;
    // RangeProof::from_bytes / InnerProductProof::from_bytes canonicity (mod.rs:519-524, inner_product_proof.rs:399-404)
    sc t_x = sc_load(proof + 128), t_x_bl = sc_load(proof + 160), e_bl = sc_load(proof + 192), a = sc_load(ab), b = sc_load(ab + 32);
    if (sc_geq_l(t_x) || sc_geq_l(t_x_bl) || sc_geq_l(e_bl) || sc_geq_l(a) || sc_geq_l(b)) { h.status = BP_PROOF_FORMAT_ERROR; return; }

    merlin_t t; merlin_load(t, tstate);
    merlin_append(t, "dom-sep", (const uint8_t *)"rangeproof v1", 13);            // transcript.rs:44-48
    merlin_append_u64(t, "n", n); merlin_append_u64(t, "m", m);
    for (uint32_t j = 0; j < m; j++) merlin_append(t, "V", V + 32 * j, 32);         // mod.rs:370-374 (identity allowed)
    bool bad = bp_is_zero32(A) || bp_is_zero32(S);                                  // validate_and_append_point
    merlin_append(t, "A", A, 32); merlin_append(t, "S", S, 32);
    sc y = rp_challenge(t, "y"); h.z = rp_challenge(t, "z");
    bad = bad || bp_is_zero32(T1) || bp_is_zero32(T2);
    merlin_append(t, "T_1", T1, 32); merlin_append(t, "T_2", T2, 32);
    h.x = rp_challenge(t, "x");
    rp_append_scalar(t, "t_x", proof + 128); rp_append_scalar(t, "t_x_blinding", proof + 160); rp_append_scalar(t, "e_blinding", proof + 192);
    h.w = rp_challenge(t, "w");
    merlin_append(t, "dom-sep", (const uint8_t *)"ipp v1", 6);                      // transcript.rs:50-53
    merlin_append_u64(t, "n", (uint64_t)n * m);
    sc u[BP_MAX_LG_N];
    for (uint32_t j = 0; j < k; j++) {                                              // inner_product_proof.rs:218-222
        bad = bad || bp_is_zero32(LR + 64 * j) || bp_is_zero32(LR + 64 * j + 32);
        merlin_append(t, "L", LR + 64 * j, 32); merlin_append(t, "R", LR + 64 * j + 32, 32);
        u[j] = rp_challenge(t, "u");
    }
    if (bad) { h.status = BP_PROOF_VERIFICATION_ERROR; return; }

    h.c = sc_mont_from_wide(weights); h.rho = sc_mont_from_wide(weights + 64);
    h.a = sc_to_mont(a); h.b = sc_to_mont(b); h.t_x = sc_to_mont(t_x); h.t_x_bl = sc_to_mont(t_x_bl); h.e_bl = sc_to_mont(e_bl);
    h.zz = sc_mont_mul(h.z, h.z);

    // one inversion for y and all u_j (Montgomery's trick); a zero challenge has probability ~2^-252:
    // dalek's invert() maps 0 to 0, and so does this (the product chain simply stays 0)
    sc pre[BP_MAX_LG_N + 1];
    sc acc = y;
    for (uint32_t j = 0; j < k; j++) { pre[j] = acc; acc = sc_mont_mul(acc, u[j]); }
    sc inv = sc_mont_invert(acc);
    sc allinv = sc_mont_one();
    for (int j = (int)k - 1; j >= 0; j--) {
        sc ui = sc_mont_mul(inv, pre[j]);          // u_j^-1
        inv = sc_mont_mul(inv, u[j]);
        allinv = sc_mont_mul(allinv, ui);
        h.u_sq[j] = sc_mont_mul(u[j], u[j]); h.u_inv_sq[j] = sc_mont_mul(ui, ui);
    }
    h.y_inv = inv; h.s0 = allinv;
    h.y_inv_pow2[0] = h.y_inv;
    for (uint32_t bb = 1; bb < k; bb++) h.y_inv_pow2[bb] = sc_mont_mul(h.y_inv_pow2[bb - 1], h.y_inv_pow2[bb - 1]);

    // delta(y,z) = (z - z^2) sum y^i - z^3 (2^n - 1) sum z^j          (mod.rs:587-593)
    sc sum_y = rp_sum_of_powers_pow2(y, (uint64_t)n * m), sum_z = rp_sum_of_powers_pow2(h.z, m);
    sc sum_2 = sc_mont_from_u64(n == 64 ? ~0ULL : ((1ULL << n) - 1));
    sc delta = sc_sub(sc_mont_mul(sc_sub(h.z, h.zz), sum_y), sc_mont_mul(sc_mont_mul(sc_mont_mul(h.zz, h.z), sum_2), sum_z));
    // basepoint scalar w (t_x - a b) + c (delta - t_x)  (mod.rs:419);  blinding scalar -e~ - c t~  (mod.rs:430)
    sc bs = sc_add(sc_mont_mul(h.w, sc_sub(h.t_x, sc_mont_mul(h.a, h.b))), sc_mont_mul(h.c, sc_sub(delta, h.t_x)));
    sc bl = sc_neg(sc_add(h.e_bl, sc_mont_mul(h.c, h.t_x_bl)));
    h.basepoint_scalar = sc_mont_mul(h.rho, bs); h.blinding_scalar = sc_mont_mul(h.rho, bl);
}

// Data-parallel tail: weighted g_i and h_i for generator pair i in [0, N) (mod.rs:415-417), Montgomery form
BP_HD void rp_prep_gh(const rp_head &h, uint32_t i, uint32_t k, uint32_t n, sc &g, sc &hh) {
    sc s = h.s0, s_rev = h.s0, yi = sc_mont_one();
    for (uint32_t bb = 0; bb < k; bb++) {
        // s_i = s_0 * prod_{bits b set in i} u_sq[(k-1)-b]  (inner_product_proof.rs:241-250);  s_{N-1-i} uses the clear bits
        if ((i >> bb) & 1u) { s = sc_mont_mul(s, h.u_sq[(k - 1) - bb]); yi = sc_mont_mul(yi, h.y_inv_pow2[bb]); }
        else s_rev = sc_mont_mul(s_rev, h.u_sq[(k - 1) - bb]);
    }
    uint32_t j = i / n, ii = i % n;
    sc zj = sc_mont_one();                         // z^j by square-and-multiply
    { sc base = h.z; for (uint32_t e = j; e; e >>= 1) { if (e & 1u) zj = sc_mont_mul(zj, base); base = sc_mont_mul(base, base); } }
    sc two_i = sc_mont_from_u64(1ULL << ii);
    sc gi = sc_sub(sc_neg(h.z), sc_mont_mul(h.a, s));
    sc hi = sc_add(h.z, sc_mont_mul(yi, sc_sub(sc_mont_mul(sc_mont_mul(h.zz, zj), two_i), sc_mont_mul(h.b, s_rev))));
    g = sc_mont_mul(h.rho, gi); hh = sc_mont_mul(h.rho, hi);
}

// Per-proof ("dynamic") scalars in MSM order A, S, T_1, T_2, L_0..L_{k-1}, R_0..R_{k-1}, V_0..V_{m-1}
// (mod.rs:422-444 without the static B~, B, G, H block); idx in [0, 4+2k+m); canonical bytes out
BP_HD sc rp_prep_dynamic(const rp_head &h, uint32_t idx, uint32_t k) {
    sc v;
    if (idx == 0) v = sc_mont_one();
    else if (idx == 1) v = h.x;
    else if (idx == 2) v = sc_mont_mul(h.c, h.x);
    else if (idx == 3) v = sc_mont_mul(sc_mont_mul(h.c, h.x), h.x);
    else if (idx < 4 + k) v = h.u_sq[idx - 4];
    else if (idx < 4 + 2 * k) v = h.u_inv_sq[idx - 4 - k];
    else {
        uint32_t j = idx - 4 - 2 * k; sc zj = sc_mont_one(), base = h.z;
        for (uint32_t e = j; e; e >>= 1) { if (e & 1u) zj = sc_mont_mul(zj, base); base = sc_mont_mul(base, base); }
        v = sc_mont_mul(sc_mont_mul(h.c, h.zz), zj);
    }
    return sc_mont_mul(h.rho, v);
}
