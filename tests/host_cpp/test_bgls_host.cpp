// C++ rendering of the reference's own tests for the hot path, run through the C++ host mirror
// (bgls_b200/host/bgls.hpp) on a GPU:
//   bgls/bgls_test.go:19-77      TestSingleSigner, TestAggregation
//   bgls/blsKosk_test.go:35-64   TestKoskMultiSig / batch multi-signature
//   curves/curve_test.go:143-186 PairingProduct = prod Pair, AggregatePoints = iterated Add
// plus the edge semantics of SURVEY.md 8a (type mismatch -> (nil,false), length mismatch, infinity, n = 0).
//
//   test_bgls_host            run the assertions, exit 0 on success
//   test_bgls_host --dump     print a deterministic transcript (fixed secret keys and messages) as `name hex` lines;
//                             tests/test_gpu_host_cpp.py recomputes every line with the oracle and compares bytes
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../bgls_b200/host/bgls.hpp"

using namespace curves;
using namespace bgls;

static int failures = 0;
#define CHECK(cond, what)                                                              \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            std::fprintf(stderr, "FAIL %s:%d %s -- %s\n", __FILE__, __LINE__, #cond, what); \
            failures++;                                                                \
        }                                                                              \
    } while (0)

static Bytes rand_bytes(size_t n, uint64_t& s) {  // splitmix64: reproducible test data
    Bytes out(n);
    for (size_t i = 0; i < n; i++) {
        if (i % 8 == 0) s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        out[i] = (uint8_t)(z >> (8 * (i % 8)));
    }
    return out;
}
static std::string hex(const Bytes& b) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t v : b) {
        s.push_back(d[v >> 4]);
        s.push_back(d[v & 15]);
    }
    return s;
}

static void TestSingleSigner(const CurveSystem& curve, uint64_t& seed) {  // bgls_test.go:19-38
    auto kp = KeyGen(curve);
    Bytes d = rand_bytes(64, seed);
    Point sig = Sign(curve, kp.first, d);
    CHECK(VerifySingleSignature(curve, sig, kp.second, d), "Standard BLS signature verification failed");
    Point sig2 = sig.Copy().Add(curve.GetG1()).first;
    CHECK(!VerifySingleSignature(curve, sig2, kp.second, d), "Standard BLS signature verification succeeding when it shouldn't");
    Bytes d2 = d;
    d2[0] ^= 1;
    CHECK(!VerifySingleSignature(curve, sig, kp.second, d2), "altered message accepted");
    CHECK(!VerifySingleSignature(curve, sig, KeyGen(curve).second, d), "altered key accepted");
}

static void TestAggregation(const CurveSystem& curve, uint64_t& seed) {  // bgls_test.go:40-77
    const int N = 6, Size = 32;
    std::vector<Bytes> msgs(N + 1);
    std::vector<Point> sigs(N + 1), pubkeys(N + 1);
    for (int i = 0; i < N; i++) {
        msgs[i] = rand_bytes(Size, seed);
        auto kp = KeyGen(curve);
        sigs[i] = Sign(curve, kp.first, msgs[i]);
        pubkeys[i] = kp.second;
    }
    auto head = [](auto& v, int n) { return std::decay_t<decltype(v)>(v.begin(), v.begin() + n); };
    Point aggSig = AggregateSignatures(head(sigs, N));
    CHECK(VerifyAggregateSignature(curve, aggSig, head(pubkeys, N), head(msgs, N)), "Aggregate Point1 verification failed");
    CHECK(!VerifyAggregateSignature(curve, aggSig, head(pubkeys, N - 1), head(msgs, N)), "succeeding without enough pubkeys");
    auto kf = KeyGen(curve);
    pubkeys[N] = kf.second;
    sigs[N] = Sign(curve, kf.first, msgs[0]);
    msgs[N] = msgs[0];
    aggSig = AggregateSignatures(sigs);
    CHECK(!VerifyAggregateSignature(curve, aggSig, pubkeys, msgs), "succeeding with duplicate messages");
    CHECK(!VerifyAggregateSignature(curve, aggSig, head(pubkeys, N), head(msgs, N)), "succeeding with invalid signature");
    msgs[0] = msgs[1];
    msgs[1] = msgs[N];
    aggSig = AggregateSignatures(head(sigs, N));
    CHECK(!VerifyAggregateSignature(curve, aggSig, head(pubkeys, N), head(msgs, N)), "succeeded with messages 0 and 1 switched");
}

static void TestKosk(const CurveSystem& curve, uint64_t& seed) {  // blsKosk_test.go:35-64,96-133
    const int N = 5;
    Bytes msg = rand_bytes(64, seed);
    std::vector<Point> sigs, keys;
    std::vector<Int> sks;
    for (int i = 0; i < N; i++) {
        auto kp = KeyGen(curve);
        sks.push_back(kp.first);
        keys.push_back(kp.second);
        sigs.push_back(KoskSign(curve, kp.first, msg));
        CHECK(KoskVerifySingleSignature(curve, sigs[i], keys[i], msg), "kosk single signature");
        CHECK(CheckAuthentication(curve, keys[i], Authenticate(curve, kp.first)), "authentication");
    }
    Point agg = AggregateSignatures(sigs);
    CHECK(KoskVerifyMultiSignature(curve, agg, keys, msg), "multi-signature rejected");
    std::vector<Point> fewer(keys.begin(), keys.end() - 1);
    CHECK(!KoskVerifyMultiSignature(curve, agg, fewer, msg), "multi-signature accepted without a key");
    CHECK(!KoskVerifyMultiSignature(curve, agg.Add(curve.GetG1()).first, keys, msg), "altered multi-signature accepted");
    // duplicate messages are fine under Kosk
    std::vector<Bytes> same(N, msg);
    CHECK(KoskVerifyAggregateSignature(curve, agg, keys, same), "kosk aggregate with equal messages rejected");
    // multiplicities: signer i signs (i+1) times
    std::vector<int64_t> mult;
    std::vector<Point> msigs;
    for (int i = 0; i < N; i++) {
        mult.push_back(i + 1);
        msigs.push_back(sigs[i].Mul(Int(i + 1)));
    }
    CHECK(KoskVerifyMultiSignatureWithMultiplicity(curve, AggregateSignatures(msigs), keys, mult, msg), "multiplicity multi-signature rejected");
    mult[0] = 3;
    CHECK(!KoskVerifyMultiSignatureWithMultiplicity(curve, AggregateSignatures(msigs), keys, mult, msg), "wrong multiplicity accepted");
    // batch: two messages, two key sets
    Bytes msg2 = rand_bytes(64, seed);
    std::vector<Point> sigs2;
    for (int i = 0; i < N; i++) sigs2.push_back(KoskSign(curve, sks[i], msg2));
    CHECK(KoskVerifyBatchMultiSignature(curve, {agg, AggregateSignatures(sigs2)}, {keys, keys}, {msg, msg2}), "batch multi-signature rejected");
    CHECK(!KoskVerifyBatchMultiSignature(curve, {agg, AggregateSignatures(sigs2)}, {keys, fewer}, {msg, msg2}), "batch multi-signature accepted with a key missing");
}

static void TestHAE(const CurveSystem& curve, uint64_t& seed) {  // blsHAE_test.go:14-83
    const int N = 5;
    std::vector<Bytes> msgs;
    std::vector<Point> sigs, pubkeys;
    for (int i = 0; i < N; i++) {
        msgs.push_back(rand_bytes(32, seed));
        auto kp = KeyGen(curve);
        sigs.push_back(Sign(curve, kp.first, msgs[i]));
        pubkeys.push_back(kp.second);
    }
    Point agg = AggregateSignaturesWithHAE(sigs, pubkeys);
    CHECK(VerifyAggregateSignatureWithHAE(curve, agg, pubkeys, msgs), "HAE aggregate verification failed");
    std::vector<Point> fewer(pubkeys.begin(), pubkeys.end() - 1);
    CHECK(!VerifyAggregateSignatureWithHAE(curve, agg, fewer, msgs), "HAE succeeding without enough pubkeys");
    CHECK(AggregateSignaturesWithHAE(sigs, fewer).nil(), "HAE aggregation with differing counts must give nil");
    CHECK(!VerifyAggregateSignatureWithHAE(curve, AggregateSignatures(sigs), pubkeys, msgs), "plain aggregate accepted as HAE aggregate");
    Bytes msg = rand_bytes(32, seed);
    std::vector<Point> msigs, signers;
    for (int j = 0; j < 8; j++) {
        auto kp = KeyGen(curve);
        msigs.push_back(Sign(curve, kp.first, msg));
        signers.push_back(kp.second);
    }
    Point magg = AggregateSignaturesWithHAE(msigs, signers);
    CHECK(VerifyMultiSignatureWithHAE(curve, magg, signers, msg), "HAE multi-signature verification failed");
    CHECK(!VerifyMultiSignatureWithHAE(curve, magg, signers, rand_bytes(32, seed)), "HAE multi-signature accepted on another message");
    signers[0] = KeyGen(curve).second;
    CHECK(!VerifyMultiSignatureWithHAE(curve, magg, signers, msg), "HAE multi-signature accepted with a wrong signer");
}

static void TestCurveLayer(const CurveSystem& curve, const CurveSystem& other, uint64_t& seed) {
    // curve_test.go:143-165: PairingProduct == product of Pair
    const int N = 5;
    std::vector<Point> p1, p2;
    for (int i = 0; i < N; i++) {
        p1.push_back(curve.HashToG1(rand_bytes(16, seed)));
        p2.push_back(KeyGen(curve).second);
    }
    auto prod = curve.PairingProduct(p1, p2);
    CHECK(prod.second, "PairingProduct not ok");
    PointT acc = curve.GetGTIdentity();
    for (int i = 0; i < N; i++) acc = acc.Add(curve.Pair(p1[i], p2[i]).first).first;
    CHECK(acc.Equals(prod.first), "PairingProduct differs from the product of Pair");
    // curve_test.go:167-186: AggregatePoints == iterated Add
    Point sum = p2[0];
    for (int i = 1; i < N; i++) sum = sum.Add(p2[i]).first;
    CHECK(sum.Equals(AggregatePoints(p2)), "AggregatePoints differs from iterated Add");
    CHECK(AggregatePoints({p2[0]}).Equals(p2[0]), "AggregatePoints of one point");
    // bilinearity through Mul: e(aP, Q) == e(P, aQ)
    Int a(12345);
    CHECK(curve.Pair(p1[0].Mul(a), p2[0]).first.Equals(curve.Pair(p1[0], p2[0].Mul(a)).first), "bilinearity");
    // edge semantics
    CHECK(!curve.PairingProduct(p1, std::vector<Point>(p2.begin(), p2.end() - 1)).second, "length mismatch must give (nil,false)");
    CHECK(!curve.Pair(p2[0], p2[0]).second, "Pair(G2, G2) must give (nil,false)");
    CHECK(!curve.Pair(other.GetG1(), p2[0]).second, "Pair with a point of another curve must give (nil,false)");
    CHECK(!p1[0].Add(p2[0]).second, "Add(G1, G2) must give (nil,false)");
    CHECK(curve.Pair(curve.GetG1Infinity(), p2[0]).first.Equals(curve.GetGTIdentity()), "Pair(inf, Q) != 1");
    CHECK(curve.Pair(p1[0], curve.GetG2Infinity()).first.Equals(curve.GetGTIdentity()), "Pair(P, inf) != 1");
    CHECK(curve.PairingProduct({}, {}).first.Equals(curve.GetGTIdentity()), "empty product != 1");
    CHECK(p1[0].Mul(Int(0)).Equals(curve.GetG1Infinity()), "0 * P != inf");
    CHECK(p1[0].Mul(Int(-1)).Add(p1[0]).first.Equals(curve.GetG1Infinity()), "-P + P != inf");
    // n = 0 signers: true iff sigma is the point at infinity (bgls.go:103-114)
    CHECK(VerifyAggregateSignature(curve, curve.GetG1Infinity(), {}, {}), "n = 0, sigma = inf must verify");
    CHECK(!VerifyAggregateSignature(curve, curve.GetG1(), {}, {}), "n = 0, sigma != inf must not verify");
    // wire formats: compressed round trip through the engine's codec
    auto back = curve.UnmarshalG2(p2[0].Marshal());
    CHECK(back.second && back.first.Equals(p2[0]), "G2 compressed round trip");
    auto back1 = curve.UnmarshalG1(p1[0].Marshal());
    CHECK(back1.second && back1.first.Equals(p1[0]), "G1 compressed round trip");
    CHECK(!curve.UnmarshalG1(Bytes(7, 1)).second, "short record must give (nil,false)");
    std::vector<Bytes> big{Bytes(curve.fp_bytes(), 0xff), Bytes(1, 2)};
    CHECK(!curve.MakeG1Point(big).second, "coordinate >= q must give (nil,false)");
}

static void Dump(const CurveSystem& curve) {
    // fixed keys and messages: every line is recomputed by the oracle in tests/test_gpu_host_cpp.py
    const int N = 4;
    std::vector<Point> keys, sigs;
    std::vector<Bytes> msgs;
    std::string n = curve.Name();
    for (int i = 0; i < N; i++) {
        Int sk(1000003 * (i + 1) + 7);
        Bytes m{(uint8_t)'m', (uint8_t)'s', (uint8_t)'g', (uint8_t)('0' + i)};
        keys.push_back(LoadPublicKey(curve, sk));
        sigs.push_back(Sign(curve, sk, m));
        msgs.push_back(m);
        std::printf("%s.pk%d %s\n", n.c_str(), i, hex(keys[i].MarshalUncompressed()).c_str());
        std::printf("%s.sig%d %s\n", n.c_str(), i, hex(sigs[i].MarshalUncompressed()).c_str());
    }
    Point agg = AggregateSignatures(sigs);
    std::printf("%s.aggsig %s\n", n.c_str(), hex(agg.MarshalUncompressed()).c_str());
    std::printf("%s.aggkey %s\n", n.c_str(), hex(AggregateKeys(keys).MarshalUncompressed()).c_str());
    std::vector<Point> hs = curve.HashToG1Many(msgs);
    std::printf("%s.product %s\n", n.c_str(), hex(curve.PairingProduct(hs, keys).first.Marshal()).c_str());
    std::printf("%s.pair %s\n", n.c_str(), hex(curve.Pair(hs[0], keys[0]).first.Marshal()).c_str());
    {
        Bytes t;
        for (auto& e : hashPubKeysToExponents(keys)) t.insert(t.end(), e.mag.begin() + 16, e.mag.end());
        std::printf("%s.hae_exponents %s\n", n.c_str(), hex(t).c_str());
        std::printf("%s.hae_aggsig %s\n", n.c_str(), hex(AggregateSignaturesWithHAE(sigs, keys).MarshalUncompressed()).c_str());
    }
    std::printf("%s.verify %d\n", n.c_str(), (int)VerifyAggregateSignature(curve, agg, keys, msgs));
    msgs[1][0] ^= 1;
    std::printf("%s.verify_bad %d\n", n.c_str(), (int)VerifyAggregateSignature(curve, agg, keys, msgs));
}

int main(int argc, char** argv) {
    try {
        if (argc > 1 && std::string(argv[1]) == "--blake2x") {   // host-only: no GPU needed
            for (int n : {0, 1, 64, 128, 129, 300}) {
                Bytes d(n);
                for (int k = 0; k < n; k++) d[k] = (uint8_t)(7 * k + n);
                for (uint32_t len : {16u, 64u, 80u, 200u}) std::printf("%d %u %s\n", n, len, hex(bgls::blake2::xof(d, len)).c_str());
            }
            return 0;
        }
        if (argc > 1 && std::string(argv[1]) == "--dump") {
            Dump(Altbn128());
            Dump(Bls12());
            return 0;
        }
        uint64_t seed = 0xB615;
        const CurveSystem* cs[2] = {&Altbn128(), &Bls12()};
        for (int i = 0; i < 2; i++) {
            TestSingleSigner(*cs[i], seed);
            TestAggregation(*cs[i], seed);
            TestKosk(*cs[i], seed);
            TestHAE(*cs[i], seed);
            TestCurveLayer(*cs[i], *cs[1 - i], seed);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::fprintf(stderr, "%d failure(s)\n", failures);
        return 1;
    }
    std::printf("ok\n");
    return 0;
}
