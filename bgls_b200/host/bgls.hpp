// bgls_b200/host/bgls.hpp -- C++ host side above the C ABI (include/bgls_b200.h).
//
// The reference is Go (compiled code) and no Go toolchain exists in the build image, so the native host mirror
// of its operator interface is C++: namespace `curves` restates package curves (curves/curve.go:12-70 CurveSystem /
// Point / PointT, :73-110 AggregatePoints, :190-214 ScalePoints, the Altbn128 / Bls12 singletons of
// curves/altbn128.go:32 and curves/bls12_381.go:31), namespace `bgls` restates the scheme layer of bgls/bgls.go and
// bgls/blsKosk.go.  Same names, argument meaning and error behaviour: Go's `(value, ok)` returns are
// std::pair<value, bool>, a nil Point is a Point with `nil() == true`.
//
// Only glue lives here (byte packing, the y -> q - y negation the reference does through
// ToAffineCoords / MakePoint, curves/altbn128.go:123-128).  Every group, hash and pairing operation is one call
// into libbgls_b200.so; without the library or a GPU the context constructor throws -- there is no CPU path.
// The cgo binding a Go maintainer adds instead of this file is in INTEGRATION.md.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bgls_b200.h"

namespace curves {

using Bytes = std::vector<uint8_t>;

// ---- *big.Int stand-in: sign + 256-bit magnitude (big-endian), all the path ever needs (secret keys, -1,
// multiplicities, 128-bit exponents)
struct Int {
    bool neg = false;
    std::array<uint8_t, 32> mag{};
    Int() = default;
    Int(int64_t v) {  // NOLINT: implicit on purpose, `Mul(-1)` reads like the reference
        neg = v < 0;
        uint64_t m = neg ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
        for (int i = 0; i < 8; i++) mag[31 - i] = (uint8_t)(m >> (8 * i));
    }
    static Int FromBytes(const uint8_t* be, size_t n) {
        if (n > 32) throw std::invalid_argument("curves::Int holds at most 256 bits");
        Int r;
        std::memcpy(r.mag.data() + 32 - n, be, n);
        return r;
    }
    bool IsZero() const {
        for (uint8_t b : mag)
            if (b) return false;
        return true;
    }
    bool IsOne() const {
        for (int i = 0; i < 31; i++)
            if (mag[i]) return false;
        return mag[31] == 1;
    }
};

namespace detail {
inline int cmp_be(const uint8_t* a, const uint8_t* b, size_t n) { return std::memcmp(a, b, n); }
// out = a - b (big-endian, a >= b)
inline void sub_be(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    int borrow = 0;
    for (size_t i = n; i-- > 0;) {
        int d = (int)a[i] - (int)b[i] - borrow;
        borrow = d < 0;
        out[i] = (uint8_t)(d + (borrow << 8));
    }
}
inline Bytes from_hex(const char* hex, size_t nbytes) {
    Bytes out(nbytes, 0);
    size_t len = std::strlen(hex);
    auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
    for (size_t i = 0; i < len; i++) {
        size_t bit = (len - 1 - i) * 4;
        out[nbytes - 1 - bit / 8] |= (uint8_t)(nib(hex[i]) << (bit % 8));
    }
    return out;
}
}  // namespace detail

class EngineError : public std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Process-wide engine context per device (the reference's curve singletons own no state; the GPU context does).
class Engine {
  public:
    static bgls_ctx* Get(int device = -1) {
        static std::mutex mu;
        static std::map<int, std::shared_ptr<Engine>> ctxs;
        if (device < 0) {
            const char* e = std::getenv("BGLS_DEVICE");
            if (!e) e = std::getenv("LOCAL_RANK");
            device = e ? std::atoi(e) : 0;
        }
        std::lock_guard<std::mutex> lk(mu);
        auto it = ctxs.find(device);
        if (it == ctxs.end()) it = ctxs.emplace(device, std::shared_ptr<Engine>(new Engine(device))).first;
        return it->second->ctx_;
    }
    ~Engine() { bgls_ctx_destroy(ctx_); }
    static void Check(bgls_ctx* ctx, int rc, const char* what) {
        if (rc != BGLS_OK) throw EngineError(std::string(what) + ": " + bgls_last_error(ctx));
    }

  private:
    explicit Engine(int device) {
        int rc = bgls_ctx_create(device, &ctx_);
        if (rc != BGLS_OK)
            throw EngineError(rc == BGLS_ERR_NODEV ? "bgls_b200: no CUDA device (the engine has no CPU fallback)"
                                                   : "bgls_b200: context creation failed");
    }
    bgls_ctx* ctx_ = nullptr;
};

class CurveSystem;

// curves/curve.go:52-60.  Holds the uncompressed affine record (MarshalUncompressed bytes); group 0 = nil.
class Point {
  public:
    Point() = default;
    Point(const CurveSystem* c, int group, Bytes raw) : curve_(c), group_(group), raw_(std::move(raw)) {}
    bool nil() const { return curve_ == nullptr; }
    const CurveSystem* curve() const { return curve_; }
    int group() const { return group_; }

    std::pair<Point, bool> Add(const Point& other) const;  // ok false on a type mismatch (altbn128.go:59-66,181-188)
    Point Copy() const { return *this; }
    bool Equals(const Point& o) const {
        return !nil() && !o.nil() && o.curve_ == curve_ && o.group_ == group_ && Canon() == o.Canon();
    }
    Bytes Marshal() const;                                  // compressed (altbn128.go:81-89,203-221)
    Bytes MarshalUncompressed() const { return Canon(); }   // altbn128.go:91-93,223-225
    Point Mul(const Int& k) const;                          // altbn128.go:107-121,235-249
    Point Negate() const;                                   // altbn128.go:123-128,227-233
    std::vector<Bytes> ToAffineCoords() const;              // [x, y] / [x_im, x_re, y_im, y_re], big-endian
    bool IsInfinity() const;
    Bytes Canon() const { return IsInfinity() ? Bytes(raw_.size(), 0) : raw_; }

  private:
    const CurveSystem* curve_ = nullptr;
    int group_ = 0;
    Bytes raw_;
};

// curves/curve.go:63-70: GT element as its 12F-byte marshal form.
class PointT {
  public:
    PointT() = default;
    PointT(const CurveSystem* c, Bytes raw) : curve_(c), raw_(std::move(raw)) {}
    bool nil() const { return curve_ == nullptr; }
    std::pair<PointT, bool> Add(const PointT& other) const;  // GT multiplication (altbn128.go:264-271)
    PointT Copy() const { return *this; }
    bool Equals(const PointT& o) const { return !nil() && !o.nil() && o.curve_ == curve_ && raw_ == o.raw_; }
    Bytes Marshal() const { return raw_; }
    PointT Mul(const Int& scalar) const;   // GT exponentiation (altbn128.go:290-294; bls12_381.go:186-195): bgls_gt_pow

  private:
    const CurveSystem* curve_ = nullptr;
    Bytes raw_;
};

// curves/curve.go:12-49
class CurveSystem {
  public:
    CurveSystem(std::string name, int cid, size_t F, const char* q, const char* order, std::vector<const char*> g1,
                std::vector<const char*> g2)
        : name_(std::move(name)), cid_(cid), F_(F), q_(detail::from_hex(q, F)), order_(detail::from_hex(order, 32)) {
        for (auto h : g1) { Bytes v = detail::from_hex(h, F); g1_.insert(g1_.end(), v.begin(), v.end()); }
        for (auto h : g2) { Bytes v = detail::from_hex(h, F); g2_.insert(g2_.end(), v.begin(), v.end()); }
    }
    std::string Name() const { return name_; }
    int id() const { return cid_; }
    size_t fp_bytes() const { return F_; }
    bgls_ctx* ctx() const { return Engine::Get(device); }
    int device = -1;

    // coords: big-endian field elements of any length <= F; (nil, false) when a coordinate is >= q or the count is wrong.
    // As in the reference the `check` flag does not trigger a subgroup check (altbn128.go:39-42; bls12_381.go:203,222).
    std::pair<Point, bool> MakeG1Point(const std::vector<Bytes>& coords, bool check = true) const { return Make(1, coords, check); }
    std::pair<Point, bool> MakeG2Point(const std::vector<Bytes>& coords, bool check = true) const { return Make(2, coords, check); }
    std::pair<Point, bool> UnmarshalG1(const Bytes& data) const { return Unmarshal(1, data); }
    std::pair<Point, bool> UnmarshalG2(const Bytes& data) const { return Unmarshal(2, data); }
    std::pair<PointT, bool> UnmarshalGT(const Bytes& data) const {
        if (data.size() != 12 * F_) return {PointT(), false};
        return {PointT(this, data), true};
    }
    Point GetG1() const { return Point(this, 1, g1_); }
    Point GetG2() const { return Point(this, 2, g2_); }
    PointT GetGT() const { return Pair(GetG1(), GetG2()).first; }
    Point GetG1Infinity() const { return Point(this, 1, Bytes(2 * F_, 0)); }
    Point GetG2Infinity() const { return Point(this, 2, Bytes(4 * F_, 0)); }
    PointT GetGTIdentity() const {  // Pair(G1, inf) in the reference (altbn128.go:478): the element 1
        Bytes one(12 * F_, 0);
        one.back() = 1;
        return PointT(this, one);
    }
    const Bytes& GetG1Q() const { return q_; }
    const Bytes& GetG1Order() const { return order_; }

    // altbn128.go:509-513 (Keccak-256 try-and-increment) / bls12_381.go:349-351 (blake2b + Fouque-Tibouchi)
    Point HashToG1(const Bytes& message) const { return HashToG1Many({message})[0]; }
    // the n concurrentHash goroutines of verifyAggSig (bgls/bgls.go:106-111,134-137) as ONE kernel launch
    std::vector<Point> HashToG1Many(const std::vector<Bytes>& msgs) const {
        std::vector<Point> out;
        if (msgs.empty()) return out;
        Bytes flat;
        std::vector<uint64_t> offs{0};
        for (auto& m : msgs) {
            flat.insert(flat.end(), m.begin(), m.end());
            offs.push_back(flat.size());
        }
        Bytes pts(msgs.size() * 2 * F_);
        uint8_t dummy = 0;
        Engine::Check(ctx(), bgls_hash_to_g1(ctx(), cid_, flat.empty() ? &dummy : flat.data(), offs.data(), msgs.size(), pts.data()), "HashToG1");
        for (size_t i = 0; i < msgs.size(); i++) out.emplace_back(this, 1, Bytes(pts.begin() + i * 2 * F_, pts.begin() + (i + 1) * 2 * F_));
        return out;
    }

    // altbn128.go:130-141; bls12_381.go:228-236: (nil, false) on a type mismatch
    std::pair<PointT, bool> Pair(const Point& p1, const Point& p2) const {
        if (!Is(p1, 1) || !Is(p2, 2)) return {PointT(), false};
        Bytes gt(12 * F_), a = p1.Canon(), b = p2.Canon();
        Engine::Check(ctx(), bgls_pair(ctx(), cid_, a.data(), b.data(), gt.data()), "Pair");
        return {PointT(this, gt), true};
    }
    // curves/curve.go:125-170 via altbn128.go:143-145 / bls12_381.go:238-240
    std::pair<PointT, bool> PairingProduct(const std::vector<Point>& pts1, const std::vector<Point>& pts2) const {
        if (pts1.size() != pts2.size()) return {PointT(), false};
        Bytes g1, g2;
        for (size_t i = 0; i < pts1.size(); i++) {
            if (!Is(pts1[i], 1) || !Is(pts2[i], 2)) return {PointT(), false};
            Bytes a = pts1[i].Canon(), b = pts2[i].Canon();
            g1.insert(g1.end(), a.begin(), a.end());
            g2.insert(g2.end(), b.begin(), b.end());
        }
        Bytes gt(12 * F_);
        int ident = 0;
        Engine::Check(ctx(), bgls_pairing_product(ctx(), cid_, g1.data(), g2.data(), pts1.size(), gt.data(), &ident), "PairingProduct");
        return {PointT(this, gt), true};
    }
    bool Is(const Point& p, int group) const { return !p.nil() && p.curve() == this && p.group() == group; }

  private:
    // the checks the reference makes before a Point exists (bgls_validate_points, mode REFERENCE): altbn128 builds every
    // point through bn256.Unmarshal (on the curve, G2 also in the order-r subgroup; `check` is ignored, altbn128.go:39-57,
    // 160-179); bls12-381 calls Check() when asked to (bls12_381.go:197-226) and on every Unmarshal (:242-264)
    std::pair<Point, bool> Validated(int group, const Bytes& raw, bool always) const {
        if (always) {
            uint8_t ok = 0;
            Engine::Check(ctx(), bgls_validate_points(ctx(), cid_, group, raw.data(), 1, BGLS_VALIDATE_REFERENCE, &ok), "validate");
            if (!ok) return {Point(), false};
        }
        return {Point(this, group, raw), true};
    }
    std::pair<Point, bool> Make(int group, const std::vector<Bytes>& coords, bool check) const {
        if (coords.size() != (size_t)2 * group) return {Point(), false};
        Bytes raw;
        for (auto& c : coords) {
            if (c.size() > F_) return {Point(), false};
            Bytes v(F_, 0);
            std::memcpy(v.data() + F_ - c.size(), c.data(), c.size());
            if (detail::cmp_be(v.data(), q_.data(), F_) >= 0) return {Point(), false};
            raw.insert(raw.end(), v.begin(), v.end());
        }
        return Validated(group, raw, cid_ == BGLS_ALTBN128 || check);
    }
    // altbn128.go:296-376, bls12_381.go:242-264: uncompressed or compressed by length; the compressed form is decoded
    // on the GPU (bls12-381 with the subgroup check the reference's Check() performs)
    std::pair<Point, bool> Unmarshal(int group, const Bytes& data) const {
        if (data.size() == 2 * group * F_) return Validated(group, data, true);
        if (data.size() != group * F_) return {Point(), false};
        Bytes raw(2 * group * F_);
        uint8_t ok = 0;
        // subgroup membership: bls12-381 Check(), and the upstream twist check of altbn128 G2
        Engine::Check(ctx(), bgls_decompress_points(ctx(), cid_, group, data.data(), 1, cid_ == BGLS_BLS12_381 || group == 2, raw.data(), &ok), "Unmarshal");
        if (!ok) return {Point(), false};
        return {Point(this, group, raw), true};
    }
    std::string name_;
    int cid_;
    size_t F_;
    Bytes q_, order_, g1_, g2_;
};

inline bool Point::IsInfinity() const {
    bool zero = true;
    for (uint8_t b : raw_) zero = zero && b == 0;
    return zero || (curve_ && curve_->id() == BGLS_BLS12_381 && !raw_.empty() && (raw_[0] & 0x40));
}
inline std::pair<Point, bool> Point::Add(const Point& o) const {
    if (nil() || o.nil() || o.curve_ != curve_ || o.group_ != group_) return {Point(), false};
    Bytes in = Canon(), b = o.Canon(), out(in.size());
    in.insert(in.end(), b.begin(), b.end());
    Engine::Check(curve_->ctx(), bgls_aggregate_points(curve_->ctx(), curve_->id(), group_, in.data(), 2, out.data()), "Point.Add");
    return {Point(curve_, group_, out), true};
}
inline Bytes Point::Marshal() const {
    Bytes in = Canon(), out(in.size() / 2);
    Engine::Check(curve_->ctx(), bgls_compress_points(curve_->ctx(), curve_->id(), group_, in.data(), 1, out.data()), "Point.Marshal");
    return out;
}
inline Point Point::Negate() const {
    if (IsInfinity()) return *this;
    const size_t F = curve_->fp_bytes(), half = raw_.size() / 2;
    Bytes out = raw_;
    const Bytes zero(F, 0);
    for (size_t o = half; o < raw_.size(); o += F)
        if (detail::cmp_be(raw_.data() + o, zero.data(), F) != 0) detail::sub_be(curve_->GetG1Q().data(), raw_.data() + o, out.data() + o, F);
    return Point(curve_, group_, out);
}
inline Point Point::Mul(const Int& k) const {
    if (k.IsZero()) return group_ == 1 ? curve_->GetG1Infinity() : curve_->GetG2Infinity();
    Point base = k.neg ? Negate() : *this;
    if (k.IsOne()) return base;
    std::array<uint8_t, 32> s = k.mag;
    const Bytes& r = curve_->GetG1Order();
    while (detail::cmp_be(s.data(), r.data(), 32) >= 0) detail::sub_be(s.data(), r.data(), s.data(), 32);
    Bytes in = base.Canon(), out(in.size());
    Engine::Check(curve_->ctx(), bgls_scale_points(curve_->ctx(), curve_->id(), group_, in.data(), s.data(), 1, out.data()), "Point.Mul");
    return Point(curve_, group_, out);
}
inline std::vector<Bytes> Point::ToAffineCoords() const {
    const size_t F = curve_->fp_bytes();
    Bytes r = Canon();
    std::vector<Bytes> out;
    for (size_t o = 0; o < r.size(); o += F) out.emplace_back(r.begin() + o, r.begin() + o + F);
    return out;
}
inline std::pair<PointT, bool> PointT::Add(const PointT& o) const {
    if (nil() || o.nil() || o.curve_ != curve_) return {PointT(), false};
    Bytes out(raw_.size());
    Engine::Check(curve_->ctx(), bgls_gt_mul(curve_->ctx(), curve_->id(), raw_.data(), o.raw_.data(), out.data()), "PointT.Add");
    return {PointT(curve_, out), true};
}

inline PointT PointT::Mul(const Int& scalar) const {
    if (nil()) throw std::logic_error("Mul on a nil PointT");
    Bytes out(raw_.size());
    Engine::Check(curve_->ctx(), bgls_gt_pow(curve_->ctx(), curve_->id(), raw_.data(), scalar.mag.data(), scalar.neg ? 1 : 0, out.data()), "PointT.Mul");
    return PointT(curve_, out);
}

// curves/curve.go:73-110.  One engine call for the whole list (the reference's goroutine tree computes the same group
// sum).  len 1 returns the point itself; len 0 never returns in the reference (curve.go:94-108) and throws here; a
// mixed list gives nil.
inline Point AggregatePoints(const std::vector<Point>& points) {
    if (points.empty()) throw std::invalid_argument("AggregatePoints of an empty list does not terminate in the reference (curves/curve.go:94-108)");
    if (points.size() == 1) return points[0];
    const Point& f = points[0];
    Bytes in;
    for (auto& p : points) {
        if (p.nil() || p.curve() != f.curve() || p.group() != f.group()) return Point();
        Bytes c = p.Canon();
        in.insert(in.end(), c.begin(), c.end());
    }
    Bytes out(in.size() / points.size());
    const CurveSystem* c = f.curve();
    Engine::Check(c->ctx(), bgls_aggregate_points(c->ctx(), c->id(), f.group(), in.data(), points.size(), out.data()), "AggregatePoints");
    return Point(c, f.group(), out);
}

// curves/curve.go:190-214: a length mismatch returns nil (empty vector + false)
inline std::pair<std::vector<Point>, bool> ScalePoints(const std::vector<Point>& pts, const std::vector<Int>& factors) {
    if (pts.size() != factors.size()) return {{}, false};
    std::vector<Point> out;
    if (pts.empty()) return {out, true};
    const CurveSystem* c = pts[0].curve();
    const int group = pts[0].group();
    Bytes in, ks;
    for (size_t i = 0; i < pts.size(); i++) {
        Bytes b = (factors[i].neg ? pts[i].Negate() : pts[i]).Canon();
        in.insert(in.end(), b.begin(), b.end());
        std::array<uint8_t, 32> s = factors[i].mag;
        while (detail::cmp_be(s.data(), c->GetG1Order().data(), 32) >= 0) detail::sub_be(s.data(), c->GetG1Order().data(), s.data(), 32);
        ks.insert(ks.end(), s.begin(), s.end());
    }
    Bytes res(in.size());
    Engine::Check(c->ctx(), bgls_scale_points(c->ctx(), c->id(), group, in.data(), ks.data(), pts.size(), res.data()), "ScalePoints");
    const size_t rec = res.size() / pts.size();
    for (size_t i = 0; i < pts.size(); i++) out.emplace_back(c, group, Bytes(res.begin() + i * rec, res.begin() + (i + 1) * rec));
    return {out, true};
}

// curve constants: curves/altbn128.go:458-480, curves/altbn128_test.go:26-35, curves/bls12_381.go:328-346
inline const CurveSystem& Altbn128() {
    static const CurveSystem c("altbn128", BGLS_ALTBN128, 32,
        "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47",
        "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001",
        {"1", "2"},
        {"198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2",
         "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
         "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b",
         "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"});
    return c;
}
inline const CurveSystem& Bls12() {
    static const CurveSystem c("bls12", BGLS_BLS12_381, 48,
        "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab",
        "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001",
        {"17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
         "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"},
        {"13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e",
         "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
         "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be",
         "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"});
    return c;
}

}  // namespace curves

// ---- scheme layer: bgls/bgls.go, bgls/blsKosk.go
namespace bgls {
using curves::Bytes;
using curves::CurveSystem;
using curves::Int;
using curves::Point;

// bgls.go:30-37: uniform secret key below the group order (rejection sampling, as crypto/rand.Int does)
inline std::pair<Int, Point> KeyGen(const CurveSystem& curve) {
    std::random_device rd;
    const Bytes& r = curve.GetG1Order();
    Int x;
    do {
        for (auto& b : x.mag) b = (uint8_t)rd();
        x.mag[0] &= 0x7f >> (r[0] < 0x40 ? 1 : 0);
    } while (curves::detail::cmp_be(x.mag.data(), r.data(), 32) >= 0);
    return {x, curve.GetG2().Mul(x)};
}
inline Point LoadPublicKey(const CurveSystem& curve, const Int& sk) { return curve.GetG2().Mul(sk); }          // bgls.go:40-43
inline Point Sign(const CurveSystem& curve, const Int& sk, const Bytes& msg) { return curve.HashToG1(msg).Mul(sk); }  // bgls.go:46-56

// bgls.go:65-70 with the curve's own hash: e(-H(m), pk) e(sig, g2) == 1
inline bool VerifySingleSignature(const CurveSystem& curve, const Point& sig, const Point& pubKey, const Bytes& msg) {
    Point h = curve.HashToG1(msg).Mul(Int(-1));
    auto paired = curve.PairingProduct({h, sig}, {pubKey, curve.GetG2()});
    return paired.second && curve.GetGTIdentity().Equals(paired.first);
}

// bgls.go:139-150
inline bool containsDuplicateMessage(const std::vector<Bytes>& msgs) {
    std::set<Bytes> seen;
    for (auto& m : msgs)
        if (!seen.insert(m).second) return true;
    return false;
}

// bgls.go:94-119: the whole body is ONE engine call (duplicate check, HashToG1 of every message, -sigma, the
// (n+1)-pair product, the comparison with 1); the length and type checks stay here as in the reference.
inline bool verifyAggSig(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys, const std::vector<Bytes>& msgs,
                         bool allowDuplicates) {
    if (keys.size() != msgs.size()) return false;
    if (!curve.Is(aggsig, 1)) return false;
    Bytes flat, kb;
    std::vector<uint64_t> offs{0};
    for (size_t i = 0; i < msgs.size(); i++) {
        if (!curve.Is(keys[i], 2)) return false;   // PairingProduct would return (nil, false)
        flat.insert(flat.end(), msgs[i].begin(), msgs[i].end());
        offs.push_back(flat.size());
        Bytes k = keys[i].Canon();
        kb.insert(kb.end(), k.begin(), k.end());
    }
    uint8_t dummy = 0;
    Bytes sig = aggsig.Canon();
    int ok = 0;
    curves::Engine::Check(curve.ctx(), bgls_verify_aggregate_signature(curve.ctx(), curve.id(), flat.empty() ? &dummy : flat.data(), offs.data(), msgs.size(),
                                                                        kb.empty() ? &dummy : kb.data(), sig.data(), allowDuplicates ? 1 : 0, &ok),
                          "verifyAggSig");
    return ok != 0;
}
inline bool VerifyAggregateSignature(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys, const std::vector<Bytes>& msgs) {
    return verifyAggSig(curve, aggsig, keys, msgs, false);  // bgls.go:82-84
}

inline Point AggregateSignatures(const std::vector<Point>& sigs) { return curves::AggregatePoints(sigs); }  // bgls.go:123-125
inline Point AggregateKeys(const std::vector<Point>& keys) { return curves::AggregatePoints(keys); }        // bgls.go:129-131

// bgls.go:89-92: vs = AggregatePoints(keys), then the single-signature check -- one engine call
inline bool verifyMultiSignature(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys, const Bytes& msg) {
    if (keys.empty()) throw std::invalid_argument("AggregatePoints of an empty list does not terminate in the reference");
    if (!curve.Is(aggsig, 1)) return false;
    Bytes kb;
    for (auto& k : keys) {
        if (!curve.Is(k, 2)) return false;
        Bytes b = k.Canon();
        kb.insert(kb.end(), b.begin(), b.end());
    }
    uint8_t dummy = 0;
    Bytes sig = aggsig.Canon();
    int ok = 0;
    curves::Engine::Check(curve.ctx(), bgls_verify_multi_signature(curve.ctx(), curve.id(), msg.empty() ? &dummy : msg.data(), msg.size(), kb.data(), keys.size(), sig.data(), &ok),
                          "verifyMultiSignature");
    return ok != 0;
}

// ---- blsKosk.go
inline Bytes kosk(const Bytes& msg) {  // append([]byte{1}, msg...)
    Bytes m{1};
    m.insert(m.end(), msg.begin(), msg.end());
    return m;
}
inline Point Authenticate(const CurveSystem& curve, const Int& sk) { return Sign(curve, sk, LoadPublicKey(curve, sk).Marshal()); }          // blsKosk.go:44-55
inline bool CheckAuthentication(const CurveSystem& curve, const Point& pubkey, const Point& auth) {                                          // blsKosk.go:59-69
    return VerifySingleSignature(curve, auth, pubkey, pubkey.Marshal());
}
inline Point KoskSign(const CurveSystem& curve, const Int& sk, const Bytes& msg) { return Sign(curve, sk, kosk(msg)); }                     // blsKosk.go:73-83
inline bool KoskVerifySingleSignature(const CurveSystem& curve, const Point& sig, const Point& pubKey, const Bytes& msg) {                  // blsKosk.go:86-97
    return VerifySingleSignature(curve, sig, pubKey, kosk(msg));
}
inline bool KoskVerifyAggregateSignature(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys, const std::vector<Bytes>& msgs) {
    std::vector<Bytes> m2;
    for (auto& m : msgs) m2.push_back(kosk(m));
    return verifyAggSig(curve, aggsig, keys, m2, true);  // blsKosk.go:100-106
}
inline bool KoskVerifyMultiSignature(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys, const Bytes& msg) {
    return verifyMultiSignature(curve, aggsig, keys, kosk(msg));  // blsKosk.go:117-120
}
inline bool KoskVerifyBatchMultiSignature(const CurveSystem& curve, const std::vector<Point>& aggsigs, const std::vector<std::vector<Point>>& pubkeys,
                                          const std::vector<Bytes>& msgs) {
    Point aggsig = AggregateSignatures(aggsigs);  // blsKosk.go:126-133
    std::vector<Point> keys;
    for (auto& ks : pubkeys) keys.push_back(AggregateKeys(ks));
    return KoskVerifyAggregateSignature(curve, aggsig, keys, msgs);
}
// blsKosk.go:137-150: keys scaled by their multiplicities, then the multi-signature check
inline bool KoskVerifyMultiSignatureWithMultiplicity(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& keys,
                                                     const std::vector<int64_t>& multiplicity, const Bytes& msg) {
    if (multiplicity.size() != keys.size()) return false;
    std::vector<Int> f(multiplicity.begin(), multiplicity.end());
    auto scaled = curves::ScalePoints(keys, f);
    return scaled.second && KoskVerifyMultiSignature(curve, aggsig, scaled.first, msg);
}

// ---- blsHAE.go: hashed aggregation exponents.  The hash G^n -> R^n is BLAKE2Xb (golang.org/x/crypto/blake2b.NewXOF),
// restated from the BLAKE2 / BLAKE2X specifications; host glue only (16 bytes per key), the scalings run on the GPU.
namespace blake2 {
inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
inline void compress(uint64_t h[8], const uint8_t block[128], uint64_t t, bool last) {
    static const uint64_t IV[8] = {0x6A09E667F3BCC908ull, 0xBB67AE8584CAA73Bull, 0x3C6EF372FE94F82Bull, 0xA54FF53A5F1D36F1ull,
                                   0x510E527FADE682D1ull, 0x9B05688C2B3E6C1Full, 0x1F83D9ABFB41BD6Bull, 0x5BE0CD19137E2179ull};
    static const uint8_t S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        m[i] = 0;
        for (int b = 7; b >= 0; b--) m[i] = (m[i] << 8) | block[8 * i + b];
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
        v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
        v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
        v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = S[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
// unkeyed BLAKE2b of `data` under an explicit parameter block (first 18 bytes given, the rest zero): 64-byte state
inline Bytes hash(const Bytes& data, uint8_t digest_len, uint8_t fanout, uint8_t depth, uint32_t leaf_len, uint32_t node_offset,
                  uint32_t xof_len, uint8_t node_depth, uint8_t inner_len) {
    static const uint64_t IV[8] = {0x6A09E667F3BCC908ull, 0xBB67AE8584CAA73Bull, 0x3C6EF372FE94F82Bull, 0xA54FF53A5F1D36F1ull,
                                   0x510E527FADE682D1ull, 0x9B05688C2B3E6C1Full, 0x1F83D9ABFB41BD6Bull, 0x5BE0CD19137E2179ull};
    uint8_t p[64] = {0};
    p[0] = digest_len; p[2] = fanout; p[3] = depth;
    for (int i = 0; i < 4; i++) { p[4 + i] = (uint8_t)(leaf_len >> (8 * i)); p[8 + i] = (uint8_t)(node_offset >> (8 * i)); p[12 + i] = (uint8_t)(xof_len >> (8 * i)); }
    p[16] = node_depth; p[17] = inner_len;
    uint64_t h[8];
    for (int i = 0; i < 8; i++) {
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | p[8 * i + b];
        h[i] = IV[i] ^ w;
    }
    const size_t nblk = data.empty() ? 1 : (data.size() + 127) / 128;
    for (size_t i = 0; i < nblk; i++) {
        uint8_t blk[128] = {0};
        const size_t off = 128 * i, len = data.size() > off ? std::min<size_t>(128, data.size() - off) : 0;
        if (len) std::memcpy(blk, data.data() + off, len);
        const bool last = i == nblk - 1;
        compress(h, blk, last ? data.size() : 128 * (i + 1), last);
    }
    Bytes out(64);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(h[i] >> (8 * b));
    return out;
}
// out_len bytes of BLAKE2Xb(data): H0 = BLAKE2b-64 with the XOF length in the parameter block, block i = BLAKE2b(H0) with
// fanout 0, depth 0, leaf length 64, node offset i, inner length 64
inline Bytes xof(const Bytes& data, uint32_t out_len) {
    Bytes h0 = hash(data, 64, 1, 1, 0, 0, out_len, 0, 0), out;
    for (uint32_t i = 0; 64 * i < out_len; i++) {
        const uint32_t j = std::min<uint32_t>(64, out_len - 64 * i);
        Bytes b = hash(h0, (uint8_t)j, 0, 0, 64, i, out_len, 0, 64);
        out.insert(out.end(), b.begin(), b.begin() + j);
    }
    return out;
}
}  // namespace blake2

inline std::vector<Int> hashPubKeysToExponents(const std::vector<Point>& pubkeys) {  // blsHAE.go:81-93
    std::vector<Int> t;
    if (pubkeys.empty()) return t;
    Bytes all;
    for (auto& pk : pubkeys) {
        Bytes b = pk.MarshalUncompressed();
        all.insert(all.end(), b.begin(), b.end());
    }
    Bytes x = blake2::xof(all, (uint32_t)(16 * pubkeys.size()));
    for (size_t i = 0; i < pubkeys.size(); i++) t.push_back(Int::FromBytes(x.data() + 16 * i, 16));
    return t;
}
inline Point AggregateSignaturesWithHAE(const std::vector<Point>& sigs, const std::vector<Point>& pubkeys) {  // blsHAE.go:39-46
    if (pubkeys.size() != sigs.size()) return Point();
    return curves::AggregatePoints(curves::ScalePoints(sigs, hashPubKeysToExponents(pubkeys)).first);
}
inline bool VerifyAggregateSignatureWithHAE(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& pubkeys,
                                            const std::vector<Bytes>& msgs) {  // blsHAE.go:49-53
    return verifyAggSig(curve, aggsig, curves::ScalePoints(pubkeys, hashPubKeysToExponents(pubkeys)).first, msgs, true);
}
inline Point getAggregatePubKey(const CurveSystem&, const std::vector<Point>& pubkeys) {  // blsHAE.go:75-78
    return curves::AggregatePoints(curves::ScalePoints(pubkeys, hashPubKeysToExponents(pubkeys)).first);
}
inline bool VerifyMultiSignatureWithHAE(const CurveSystem& curve, const Point& aggsig, const std::vector<Point>& pubkeys, const Bytes& msg) {
    return VerifySingleSignature(curve, aggsig, getAggregatePubKey(curve, pubkeys), msg);  // blsHAE.go:56-58
}

}  // namespace bgls
