"""ctypes view of oracle/liboracle.so — the CPU restatement of the reference used as the checker.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_sz, _vp, _p = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p

L_ORDER = 2**252 + 27742317777372353535851937790883648493
P_FIELD = 2**255 - 19


class Oracle:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            from bulletproofs_b200 import build as b
            b.build_oracle()
        L = self.L = ctypes.CDLL(path)
        L.orc_gens_new.restype = _vp
        L.orc_gens_new.argtypes = [_sz, _sz]
        L.orc_gens_get.argtypes = [_vp, ctypes.c_int, _sz, _sz, _p]
        L.orc_transcript_size.restype = _sz
        L.orc_transcript_new.argtypes = [_p, _p, _sz]
        L.orc_transcript_append.argtypes = [_p, _p, _p, _sz]
        L.orc_transcript_challenge.argtypes = [_p, _p, _p, _sz]
        L.orc_rangeproof_size.restype = _sz
        L.orc_rangeproof_size.argtypes = [_sz, _sz]
        L.orc_rangeproof_verify.argtypes = [_vp, _p, _p, _sz, _p, _sz, _sz, _p]
        L.orc_rangeproof_prove.argtypes = [_vp, _p, ctypes.POINTER(ctypes.c_uint64), _p, _sz, _sz, _p, _p, _p]
        L.orc_rangeproof_verify_many.argtypes = [_vp, _p, _p, _sz, _p, _sz, _sz, _sz, _p, ctypes.c_int, _p]
        L.orc_rangeproof_prove_many.argtypes = [_vp, _p, ctypes.POINTER(ctypes.c_uint64), _p, _sz, _sz, _sz, _p, ctypes.c_int, _p, _p, _p]
        L.orc_msm.argtypes = [_p, _p, _sz, _p]
        L.orc_msm_naive.argtypes = [_p, _p, _sz, _p]
        L.orc_chacha_fill.argtypes = [_p, _p, _sz]
        L.orc_sha3_512.argtypes = [_p, _sz, _p]
        L.orc_shake256.argtypes = [_p, _sz, _p, _sz]
        L.orc_ipp_create.argtypes = [_p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]
        L.orc_ipp_verify.argtypes = [_p, _sz, _p, _p, _p, _p, _p, _p, _p, _sz]
        L.orc_linear_create.argtypes = [_p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]
        L.orc_linear_verify.argtypes = [_p, _p, _sz, _p, _p, _p, _p, _p, _sz]
        L.orc_r1cs_prove.argtypes = [_vp, _p, ctypes.c_int, _p, _p, _sz, ctypes.c_uint64, ctypes.c_uint64, _p, _p, ctypes.POINTER(_sz), _p]
        L.orc_r1cs_verify.argtypes = [_vp, _p, ctypes.c_int, _p, _sz, ctypes.c_uint64, _p, _sz, _p]
        u64 = ctypes.c_uint64
        L.orc_mpc_party_bit_commitment.argtypes = [_vp, u64, _p, _sz, _sz, _p, _p]
        L.orc_mpc_party_poly_commitment.argtypes = [_vp, u64, _sz, _sz, _p, _p, _p, _p]
        L.orc_mpc_party_proof_share.argtypes = [_vp, u64, _p, _sz, _sz, _p, _p, _p, _p, _p]
        L.orc_mpc_dealer_bit_challenge.argtypes = [_vp, _p, _sz, _sz, _p, _p, _p]
        L.orc_mpc_dealer_poly_challenge.argtypes = [_p, _sz, _p, _p]
        L.orc_mpc_dealer_run.argtypes = [_vp, _p, _sz, _sz, _p, _p, _p, ctypes.c_int, _p, _p, _p]
        L.orc_mpc_audit_share.argtypes = [_vp, _sz, _sz, _p, _p, _p, _p, _p, _p]
        L.orc_init()
        assert L.orc_selfcheck() == 0
        self.tsize = L.orc_transcript_size()

    # generators
    def gens(self, cap, parties):
        return _vp(self.L.orc_gens_new(cap, parties))

    def gens_get(self, g, which, party, idx):
        o = ctypes.create_string_buffer(32)
        self.L.orc_gens_get(g, which, party, idx, o)
        return o.raw

    def pedersen(self):
        b, bb = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        self.L.orc_pedersen_gens(b, bb)
        return b.raw, bb.raw

    # transcripts (opaque oracle-side state; its first 203 bytes equal the wire state of the C ABI)
    def transcript(self, label: bytes):
        st = ctypes.create_string_buffer(256)
        self.L.orc_transcript_new(st, label, len(label))
        return st.raw

    def transcript_append(self, st, label, msg):
        b = ctypes.create_string_buffer(st, 256)
        self.L.orc_transcript_append(b, label, msg, len(msg))
        return b.raw

    def transcript_challenge(self, st, label, n):
        b = ctypes.create_string_buffer(st, 256)
        o = ctypes.create_string_buffer(n)
        self.L.orc_transcript_challenge(b, label, o, n)
        return b.raw, o.raw

    # range proofs
    def rangeproof_size(self, n, m):
        return self.L.orc_rangeproof_size(n, m)

    def rangeproof_verify(self, g, tstate, proof, V, m, n, seed=bytes(32)):
        return self.L.orc_rangeproof_verify(g, tstate, proof, len(proof), V, m, n, seed)

    def rangeproof_prove(self, g, tstate, values, blindings, n, seed=bytes(32)):
        m = len(values)
        vals = (ctypes.c_uint64 * m)(*values)
        proof = ctypes.create_string_buffer(self.rangeproof_size(n, m))
        V = ctypes.create_string_buffer(32 * m)
        rc = self.L.orc_rangeproof_prove(g, tstate, vals, blindings, m, n, seed, proof, V)
        return rc, proof.raw, V.raw

    def prove_many(self, g, tstate, values, blindings, n, m, seeds, nthreads=8):
        count = len(values) // m
        vals = (ctypes.c_uint64 * len(values))(*values)
        plen = self.rangeproof_size(n, m)
        proofs = ctypes.create_string_buffer(plen * count)
        Vs = ctypes.create_string_buffer(32 * m * count)
        st = ctypes.create_string_buffer(count)
        self.L.orc_rangeproof_prove_many(g, tstate, vals, blindings, m, n, count, seeds, nthreads, proofs, Vs, st)
        assert not any(st.raw), "oracle prover failed"
        return proofs.raw, Vs.raw

    def verify_many(self, g, tstate, proofs, plen, Vs, n, m, count, seeds=None, nthreads=8):
        st = ctypes.create_string_buffer(count)
        self.L.orc_rangeproof_verify_many(g, tstate, proofs, plen, Vs, m, n, count, seeds or bytes(32 * count), nthreads, st)
        return list(st.raw)

    def verify_rlc(self, g, tstate, proofs, plen, Vs, n, m, count, seed=bytes(32), nthreads=8):
        """CPU random-linear-combination batch (one combined MSM per thread's chunk, per-proof recheck of a failing chunk)"""
        self.L.orc_rangeproof_verify_rlc.argtypes = [_vp, _p, _p, _sz, _p, _sz, _sz, _sz, _p, ctypes.c_int, _p]
        st = ctypes.create_string_buffer(count)
        self.L.orc_rangeproof_verify_rlc(g, tstate, proofs, plen, Vs, m, n, count, seed, nthreads, st)
        return list(st.raw)

    def scalars_from_chacha(self, seed, count, skip=0):
        self.L.orc_scalars_from_chacha.argtypes = [_p, _sz, _sz, _p]
        out = ctypes.create_string_buffer(32 * count)
        self.L.orc_scalars_from_chacha(seed, skip, count, out)
        return out.raw

    # MSM field backends (oracle/vec4_*.h): "u64", "avx2", "ifma", "auto"; process-wide, not thread-safe
    def set_backend(self, name):
        self.L.orc_set_backend.argtypes = [_p]
        return self.L.orc_set_backend(name.encode())

    def backend_name(self):
        self.L.orc_backend_name.restype = ctypes.c_char_p
        return self.L.orc_backend_name().decode()

    def vec_selftest(self, name, points4, rnd8):
        self.L.orc_vec_selftest.argtypes = [_p, _p, _p]
        return self.L.orc_vec_selftest(name.encode(), points4, rnd8)

    # group / scalar helpers
    def msm(self, scalars, points, naive=False):
        o = ctypes.create_string_buffer(32)
        rc = (self.L.orc_msm_naive if naive else self.L.orc_msm)(scalars, points, len(scalars) // 32, o)
        return rc, o.raw

    def from_uniform(self, b64):
        o = ctypes.create_string_buffer(32)
        self.L.orc_from_uniform_bytes(b64, o)
        return o.raw

    def point_is_valid(self, p):
        return self.L.orc_point_is_valid(p)

    def point_add(self, a, b):
        o = ctypes.create_string_buffer(32)
        assert self.L.orc_point_add(a, b, o) == 0
        return o.raw

    def point_double_encode(self, a):
        o = ctypes.create_string_buffer(32)
        assert self.L.orc_point_double_encode(a, o) == 0
        return o.raw

    def chacha(self, seed, n):
        o = ctypes.create_string_buffer(n)
        self.L.orc_chacha_fill(seed, o, n)
        return o.raw

    def scalar_from_wide(self, b64):
        o = ctypes.create_string_buffer(32)
        self.L.orc_scalar_from_wide(b64, o)
        return o.raw

    def random_scalars(self, seed, n):
        ks = self.chacha(seed, 64 * n)
        return b"".join(self.scalar_from_wide(ks[64 * i:64 * i + 64]) for i in range(n))

    def shake256(self, data, n):
        o = ctypes.create_string_buffer(n)
        self.L.orc_shake256(data, len(data), o, n)
        return o.raw

    def sha3_512(self, data):
        o = ctypes.create_string_buffer(64)
        self.L.orc_sha3_512(data, len(data), o)
        return o.raw

    # LinearProof
    def linear_create(self, tstate, seed, C, r, a, b, G, F, B):
        n = len(a) // 32
        st = ctypes.create_string_buffer(tstate, 256); out = ctypes.create_string_buffer(32 * (2 * (n.bit_length() - 1) + 3))
        rc = self.L.orc_linear_create(st, seed, C, r, a, b, G, F, B, n, out)
        return rc, st.raw, out.raw

    def linear_verify(self, tstate, proof, C, G, F, B, b):
        st = ctypes.create_string_buffer(tstate, 256)
        return self.L.orc_linear_verify(st, proof, len(proof), C, G, F, B, b, len(b) // 32)

    # R1CS (gadget 0 shuffle, 1 example, 2 range)
    def r1cs_prove(self, g, tstate, gadget, values, blindings, param=0, aux=0, ext_seed=bytes(32)):
        m = len(values)
        out = ctypes.create_string_buffer(1 + 32 * 14 + 32 * (2 * 32 + 2)); n = _sz(); V = ctypes.create_string_buffer(32 * max(m, 1))
        rc = self.L.orc_r1cs_prove(g, tstate, gadget, b"".join(int(v).to_bytes(32, "little") for v in values), blindings, m, param, aux, ext_seed, out, ctypes.byref(n), V)
        return rc, out.raw[:n.value], V.raw[:32 * m]

    def r1cs_verify(self, g, tstate, gadget, commitments, proof, param=0, ext_seed=bytes(32)):
        return self.L.orc_r1cs_verify(g, tstate, gadget, commitments, len(commitments) // 32, param, proof, len(proof), ext_seed)

    def ipp_create(self, tstate, Q, Gf, Hf, G, H, a, b, n):
        st = ctypes.create_string_buffer(tstate, 256)
        k = n.bit_length() - 1
        out = ctypes.create_string_buffer(32 * (2 * k + 2))
        rc = self.L.orc_ipp_create(st, Q, Gf, Hf, G, H, a, b, n, out)
        return rc, st.raw, out.raw

    def ipp_verify(self, tstate, n, Gf, Hf, P, Q, G, H, proof):
        st = ctypes.create_string_buffer(tstate, 256)
        return self.L.orc_ipp_verify(st, n, Gf, Hf, P, Q, G, H, proof, len(proof))

    # aggregated range-proof MPC messages (oracle/mpc.h); stateless: a party is (v, v_blinding, n, j, seed)
    def mpc_bit_commitment(self, g, v, v_blinding, n, j, seed):
        o = ctypes.create_string_buffer(96)
        return self.L.orc_mpc_party_bit_commitment(g, v, v_blinding, n, j, seed, o), o.raw

    def mpc_poly_commitment(self, g, v, n, j, seed, y, z):
        o = ctypes.create_string_buffer(64)
        return self.L.orc_mpc_party_poly_commitment(g, v, n, j, seed, y, z, o), o.raw

    def mpc_proof_share(self, g, v, v_blinding, n, j, seed, y, z, x):
        o = ctypes.create_string_buffer(32 * (3 + 2 * n))
        return self.L.orc_mpc_party_proof_share(g, v, v_blinding, n, j, seed, y, z, x, o), o.raw

    def mpc_bit_challenge(self, g, tstate, n, m, bitc):
        st = ctypes.create_string_buffer(tstate, len(tstate)); y = ctypes.create_string_buffer(32); z = ctypes.create_string_buffer(32)
        rc = self.L.orc_mpc_dealer_bit_challenge(g, st, n, m, bitc, y, z)
        return rc, st.raw, y.raw, z.raw

    def mpc_poly_challenge(self, tstate, m, polyc):
        st = ctypes.create_string_buffer(tstate, len(tstate)); x = ctypes.create_string_buffer(32)
        rc = self.L.orc_mpc_dealer_poly_challenge(st, m, polyc, x)
        return rc, st.raw, x.raw

    def mpc_dealer_run(self, g, initial_tstate, n, m, bitc, polyc, shares, trusted=False, verify_seed=bytes(32)):
        proof = ctypes.create_string_buffer(self.rangeproof_size(n, m)); bad = ctypes.create_string_buffer(m)
        rc = self.L.orc_mpc_dealer_run(g, initial_tstate, n, m, bitc, polyc, shares, int(trusted), verify_seed, proof, bad)
        return rc, proof.raw, list(bad.raw)

    def mpc_audit_share(self, g, n, j, bitc, y, z, polyc, x, share):
        return self.L.orc_mpc_audit_share(g, n, j, bitc, y, z, polyc, x, share)
