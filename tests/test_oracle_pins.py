"""CPU tests that PIN the oracle: the one literal constant of the reference, the published merlin / Keccak
vectors for the third-party transcript, and the reference's own property tests restated."""
import hashlib
import random

from oracle.pyref import field as F
from oracle.pyref import gates as G
from oracle.pyref import sumcheck as S
from oracle.pyref.transcript import MerlinTranscript, ProofTranscript2, keccak_f1600

P = F.P


def test_coeff_d_literal_of_reference():
    # src/utils.rs:34-37 is the only literal field constant in the reference tree
    assert F.limbs_to_int(F.REF_COEFF_D_MONT_LIMBS) == F.fr_to_mont(F.TE_D)
    # Bandersnatch d = 138827208126141220649022263972958607803 / 171449701953573178309673572579671231137
    assert F.TE_D == 138827208126141220649022263972958607803 * pow(171449701953573178309673572579671231137, -1, P) % P


def test_keccak_against_hashlib():
    for msg in (b"", b"abc", b"x" * 135):
        st = bytearray(200)
        assert len(msg) < 136
        st[: len(msg)] = msg
        st[len(msg)] ^= 0x06
        st[135] ^= 0x80
        keccak_f1600(st)
        assert bytes(st[:32]) == hashlib.sha3_256(msg).digest()


def test_merlin_published_vector():
    # merlin's documented transcript test vector ("Transcript Protocol" page of the merlin docs)
    t = MerlinTranscript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_transcript2_io():
    # proof_transcript.rs:159-193: what the prover wrote is what the verifier reads, challenges replay
    rng = random.Random(3)
    p = ProofTranscript2.start_prover(b"fgstglsp")
    xs = [rng.randrange(P) for _ in range(5)]
    p.write_scalars(xs[:2])
    c1 = p.challenge(128)
    p.write_scalars(xs[2:])
    c2 = p.challenge(512)
    proof = p.end()
    assert len(proof) == 5 * 32
    v = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    assert v.read_scalars(2) == xs[:2]
    assert v.challenge(128) == c1
    assert v.read_scalars(3) == xs[2:]
    assert v.challenge(512) == c2
    assert c1 < (1 << 128)


def test_eq_tables():
    rng = random.Random(5)
    pt = [rng.randrange(P) for _ in range(5)]
    e = S.eq_poly_sequence_last(pt)
    assert sum(e) % P == 1
    for idx in (0, 7, 19, 31):
        bits = [(idx >> (4 - k)) & 1 for k in range(5)]  # pt[0] <-> most significant bit
        assert e[idx] == S.eq_eval(pt, bits)
    # eq_sum_works (src/utils.rs tests): prefix sums
    for k in (0, 1, 13, 32):
        assert S.eq_sum(pt, k) == sum(e[:k]) % P
    # padded sequence == ordinary sequence truncated (utils.rs:189-220)
    seq = S.padded_eq_poly_sequence(2, pt)
    full = S.eq_poly_sequence(pt)
    for i, lvl in enumerate(seq):
        assert lvl == full[i][: len(lvl)]


def test_unipoly_roundtrip():
    rng = random.Random(9)
    for n in (2, 3, 4, 5):
        coeffs = [rng.randrange(P) for _ in range(n)]
        evals = [S.evaluate_univar(coeffs, i) for i in range(n)]
        assert S.unipoly_from_evals(evals) == coeffs
        s = (evals[0] + evals[1]) % P
        assert S.decompress_coefficients(S.compress_coefficients(coeffs), s) == coeffs


def _rand_te_like(rng, nv, n_polys):
    return [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(n_polys)]


def test_dense_deg2_equals_naive_check_univars():
    # dense_eq.rs:259-343 `check_univars`
    rng = random.Random(11)
    for gate in (G.PrjL1(), G.AffL1(), G.PrjL2(), G.PrjL3(), G.AffL3(), G.TriL1()):
        nv = 4
        data = _rand_te_like(rng, nv, gate.n_ins)
        point = [rng.randrange(P) for _ in range(nv)]
        gamma = rng.randrange(P)
        f = G.EqWrapper(G.GammaWrapper(gate, gamma))
        eqp = S.eq_poly_sequence_last(point)
        dense = [list(p) for p in data] + [eqp]
        claims = [0] * gate.n_outs
        for i in range(1 << nv):
            o = gate.exec([d[i] for d in data])
            for k in range(gate.n_outs):
                claims[k] = (claims[k] + o[k] * eqp[i]) % P
        opt = S.DenseDeg2SumcheckObjectSO.rlc(data, gate, claims, point, gamma)
        ex = S.ExampleSumcheckObjectSO(dense, f, nv)
        for _ in range(nv):
            assert opt.unipoly() == ex.unipoly()
            t = rng.randrange(P)
            opt.bind(t)
            ex.bind(t)
            assert ex.claim() == opt.claim
        assert opt.final_evals() == ex.final_evals()[:-1]


def test_vecvec_deg2_equals_naive_check_univars():
    # vecvec_eq.rs:511-600: {Full, Rows, Nothing} density x vertical vars {0,1,3}
    rng = random.Random(13)
    gate = G.PrjL1()
    for dens in range(3):
        for colv in (0, 1, 3):
            nv = 6
            rowv = nv - colv
            nrows = (1 << colv) if dens < 2 else rng.randrange(0, 1 << colv) + 1
            lens = [(1 << rowv) if dens == 0 else rng.randrange(0, 1 << rowv) + 1 for _ in range(nrows)]
            pads = [(0, 0), (1, 1), (1, 1)] * 2
            polys = [S.VecVecPolynomial([[rng.randrange(P) for _ in range(l)] for l in lens], rp, cp, rowv, colv) for rp, cp in pads]
            point = [rng.randrange(P) for _ in range(nv)]
            gamma = rng.randrange(P)
            f = G.EqWrapper(G.GammaWrapper(gate, gamma))
            eqp = S.eq_poly_sequence_last(point)
            dense = [p.vec() for p in polys] + [eqp]
            claims = [0] * 4
            for i in range(1 << nv):
                o = gate.exec([dense[j][i] for j in range(6)])
                for k in range(4):
                    claims[k] = (claims[k] + o[k] * eqp[i]) % P
            opt = S.VecVecDeg2SumcheckObjectSO.rlc(polys, gate, claims, point, colv, gamma)
            ex = S.ExampleSumcheckObjectSO(dense, f, nv)
            for _ in range(nv):
                assert opt.unipoly() == ex.unipoly()
                t = rng.randrange(P)
                opt.bind(t)
                ex.bind(t)
                assert ex.claim() == opt.claim
            assert opt.final_evals() == ex.final_evals()


def test_dense_so_equals_naive_and_verifier_accepts():
    # sumcheck.rs:941-1078 restated: prover <-> verifier round trip + final evals == MLE evaluation
    rng = random.Random(17)
    nv = 5
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    so = S.DenseSumcheckObjectSO(polys, f, nv, claim)
    prot = S.BareSumcheckSO(f, nv)
    point, evs = prot.prove(tp, claim, so)
    proof = tp.end()
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    assert prot.verify(tv, claim) == (point, evs)
    assert [S.evaluate_poly(p, point) for p in polys] == evs


def test_dense_eq_sumcheck_roundtrip():
    # sumcheck.rs:1041-1078 `dense_sumcheck_with_eq_verifier_accepts_prover` with the logup gate
    rng = random.Random(19)
    nv = 4
    gate = G.LogupLayer()
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(4)]
    point = [rng.randrange(P) for _ in range(nv)]
    outs = [[], []]
    for i in range(1 << nv):
        o = gate.exec([p[i] for p in polys])
        outs[0].append(o[0])
        outs[1].append(o[1])
    evs = [S.evaluate_poly(o, point) for o in outs]
    prot = S.DenseEqSumcheck(gate, nv)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    out_claims = prot.prove(tp, (point, evs), polys)
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tp.end())
    assert prot.verify(tv, (point, evs)) == out_claims
    new_point, new_evs = out_claims
    assert [S.evaluate_poly(p, new_point) for p in polys] == new_evs


def test_gamma_wrapper_works():
    # sumcheck.rs:1080-1092
    rng = random.Random(23)
    gate = G.PrjL2()
    a = [rng.randrange(P) for _ in range(4)]
    gamma = rng.randrange(P)
    out = gate.exec(a)
    assert G.GammaWrapper(gate, gamma).exec(a) == sum(pow(gamma, i, P) * x for i, x in enumerate(out)) % P


def test_binary_msm_oracle_matches_reference_property():
    """binary_msm.rs:62-96 (bin_msm, bin_msm_gamma_3) restated: binary_msm(prepare_coefs, prepare_bases) == sum of the selected bases"""
    import random

    from oracle.pyref import commitments as OC
    from oracle.pyref import curves as CV

    rng = random.Random(8)
    for num, gamma in [(20, 8), (20, 3), (5, 4)]:
        bits = [rng.random() < 0.5 for _ in range(num)]
        bases = [CV.g1_mul(rng.randrange(1, CV.G1_ORDER), CV.G1_GEN) for _ in range(num)]
        res = OC.binary_msm(OC.prepare_coefs(bits, gamma), OC.prepare_bases(bases, gamma))
        expected = None
        for c, b in zip(bits, bases):
            if c:
                expected = CV.g1_add(expected, b)
        assert res == expected
    assert OC.into_u8([True, False, True]) == 5


def test_published_curve_parameters_and_wire_format():
    """Third-party facts the restatement depends on, pinned against the PUBLISHED values (not derived from the oracle):
    BLS12-381 base / scalar field moduli and G1 generator (IETF pairing-friendly-curves draft, zkcrypto/bls12_381), the
    standard 48-byte compressed encoding of that generator and of the point at infinity (the format ark-bls12-381 0.4 writes
    into the proof, proof_transcript.rs:52-69), and the Bandersnatch parameters of ark-ed-on-bls12-381-bandersnatch 0.4
    (a = -5, d, prime-subgroup generator and order; Masson-Sanso-Zhang, 'Bandersnatch', table 1)."""
    from oracle.pyref import curves as CV
    from oracle.pyref import pippenger as PP
    from oracle.pyref.field import FQ_MODULUS, P

    assert FQ_MODULUS == 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    assert P == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    gx = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
    gy = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
    assert CV.G1_GEN == (gx, gy) and CV.G1_ORDER == P
    assert (gy * gy - gx * gx * gx - 4) % FQ_MODULUS == 0
    assert CV.g1_mul(P, CV.G1_GEN) is None and CV.g1_mul(P - 1, CV.G1_GEN) == CV.g1_neg(CV.G1_GEN)
    assert PP.g1_serialize(CV.G1_GEN).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                                 "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert PP.g1_serialize(None) == bytes([0xC0]) + bytes(47)
    assert PP.g1_serialize(CV.g1_neg(CV.G1_GEN))[0] == 0xB7  # the other root of y: sign flag set
    for pt in (CV.G1_GEN, CV.g1_neg(CV.G1_GEN), CV.g1_mul(12345, CV.G1_GEN), None):
        assert PP.g1_deserialize(PP.g1_serialize(pt)) == pt
    # Bandersnatch over Fr
    assert CV.TE_A % P == P - 5
    assert CV.TE_D == 45022363124591815672509500913686876175488063829319466900776701791074614335719
    assert CV.TE_D == 0x6389C12633C267CBC66E3BF86BE3B6D8CB66677177E54F92B369F2F5188D58E7
    assert CV.TE_GEN == (18886178867200960497001835917649091219057080094937609519140440539760939937304,
                         19188667384257783945677642223292697773471335439753913231509108946878080696678)
    assert CV.TE_SUBGROUP_ORDER == 13108968793781547619861935127046491459309155893440570251786403306729687672801
    assert CV.te_on_curve(CV.TE_GEN)
