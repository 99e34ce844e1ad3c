// include/circl_b200.hpp -- C++ host-side mirror of CIRCL's kem.Scheme / sign.Scheme over the C ABI.
//
// CIRCL is Go and this build image has no Go toolchain, so besides the cgo shim delivered as source
// (go/), this header is the *compiled* host side above libcirclb200.so: same method names, argument
// meaning and error behaviour as
//   kem/kem.go:14-121            kem.Scheme, kem.PublicKey/PrivateKey, kem.Err*
//   kem/mlkem/mlkem768/kyber.go:267-407   (scheme boilerplate; mlkem512/1024 identical)
//   sign/sign.go:14-119          sign.Scheme, sign.SignatureOpts, sign.Err*
//   sign/mldsa/mldsa65/dilithium.go:256-366
//   kem/schemes/schemes.go:57-72, sign/schemes/schemes.go:56-71   (ByName, case-insensitive)
// Go's (value, error) returns become exceptions; Go panics (programmer errors) become std::logic_error.
// Batch methods are added beside the single-op ones.  Header-only; link with -lcirclb200.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <cerrno>
#include <string>
#include <utility>
#include <vector>

#include <sys/random.h>

#include "circl_b200.h"

namespace circl {

using Bytes = std::vector<uint8_t>;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void init(int device = 0) {
  if (cb200_init(device) != 0) throw Error(cb200_last_error());
}
// one process, GPUs 0..ndev-1 (0: all): host-pointer batches are sharded by index inside the library
inline void init_devices(int ndev = 0) {
  if (cb200_init_devices(ndev) != 0) throw Error(cb200_last_error());
}
// crypto/rand of the mirror: the kernel's CSPRNG (getrandom(2)), as Go's crypto/rand reads it on Linux
inline void fill_random(uint8_t* p, size_t n) {
  size_t got = 0;
  while (got < n) {
    const ssize_t r = getrandom(p + got, n - got, 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      throw Error("getrandom failed");
    }
    got += (size_t)r;
  }
}

namespace kem {

struct ErrTypeMismatch : Error { ErrTypeMismatch() : Error("kem: type mismatch") {} };
struct ErrSeedSize : Error { ErrSeedSize() : Error("kem: invalid seed size") {} };
struct ErrPubKeySize : Error { ErrPubKeySize() : Error("kem: invalid public key size") {} };
struct ErrCiphertextSize : Error { ErrCiphertextSize() : Error("kem: invalid ciphertext size") {} };
struct ErrPrivKeySize : Error { ErrPrivKeySize() : Error("kem: invalid private key size") {} };
struct ErrPubKey : Error { ErrPubKey() : Error("kem: invalid public key") {} };
struct ErrPrivKey : Error { ErrPrivKey() : Error("kem: invalid private key") {} };

class Scheme;

class PublicKey {  // kem.PublicKey (kem/kem.go:14-20)
 public:
  PublicKey(const Scheme* s, Bytes b) : scheme_(s), packed_(std::move(b)) {}
  const Scheme* GetScheme() const { return scheme_; }
  const Bytes& MarshalBinary() const { return packed_; }
  bool Equal(const PublicKey& o) const { return scheme_ == o.scheme_ && packed_ == o.packed_; }

 private:
  const Scheme* scheme_;
  Bytes packed_;
};

class PrivateKey {  // kem.PrivateKey (kem/kem.go:22-30)
 public:
  PrivateKey(const Scheme* s, Bytes b) : scheme_(s), packed_(std::move(b)) {}
  const Scheme* GetScheme() const { return scheme_; }
  const Bytes& MarshalBinary() const { return packed_; }
  bool Equal(const PrivateKey& o) const { return scheme_ == o.scheme_ && packed_ == o.packed_; }
  PublicKey Public() const;

 private:
  const Scheme* scheme_;
  Bytes packed_;
};

// kem.Scheme (kem/kem.go:33-82).  One class serves the four families behind the C ABI:
//   MLKEM   ML-KEM-512/768/1024            kem/mlkem/mlkem768/kyber.go:267-407
//   KYBER3  round-3 Kyber512/768/1024      kem/kyber/kyber768/kyber.go:267-407
//   XWING   X-Wing                         kem/xwing/scheme.go:1-140
//   HYBRID  X25519MLKEM768, Kyber768-X25519, Kyber512-X25519   kem/hybrid/hybrid.go:76-315
class Scheme {
 public:
  enum Kind { MLKEM, KYBER3, XWING, HYBRID };
  Scheme(std::string name, int k, bool round3 = false) : name_(std::move(name)), kind_(round3 ? KYBER3 : MLKEM), k_(k) {}
  Scheme(std::string name, Kind kind, int id) : name_(std::move(name)), kind_(kind), k_(id) {}
  const std::string& Name() const { return name_; }
  size_t CiphertextSize() const {
    return kind_ == XWING ? 1120 : kind_ == HYBRID ? cb200_hybrid_ciphertext_size(k_) : cb200_mlkem_ciphertext_size(k_);
  }
  size_t SharedKeySize() const { return kind_ == HYBRID ? 64 : 32; }
  size_t PrivateKeySize() const {
    return kind_ == XWING ? 32 : kind_ == HYBRID ? cb200_hybrid_private_key_size(k_) : cb200_mlkem_private_key_size(k_);
  }
  size_t PublicKeySize() const {
    return kind_ == XWING ? 1216 : kind_ == HYBRID ? cb200_hybrid_public_key_size(k_) : cb200_mlkem_public_key_size(k_);
  }
  size_t SeedSize() const { return kind_ == XWING ? 32 : 64; }
  size_t EncapsulationSeedSize() const { return kind_ == XWING ? 64 : 32; }

  std::pair<PublicKey, PrivateKey> DeriveKeyPair(const Bytes& seed) const {  // kyber.go:337-346
    if (seed.size() != SeedSize()) throw std::logic_error("kem: invalid seed size");  // Go panics here
    Bytes ek, dk;
    DeriveKeyPairBatch(seed, ek, dk);
    return {PublicKey(this, std::move(ek)), PrivateKey(this, std::move(dk))};
  }
  std::pair<PublicKey, PrivateKey> GenerateKeyPair() const {  // kyber.go:281-283
    Bytes seed(SeedSize());
    fill_random(seed.data(), seed.size());
    return DeriveKeyPair(seed);
  }
  PublicKey UnmarshalBinaryPublicKey(const Bytes& buf) const {  // kyber.go:390-396
    if (buf.size() != PublicKeySize()) throw ErrPubKeySize();
    return PublicKey(this, buf);
  }
  PrivateKey UnmarshalBinaryPrivateKey(const Bytes& buf) const {  // kyber.go:398-407
    if (buf.size() != PrivateKeySize()) throw ErrPrivKeySize();
    return PrivateKey(this, buf);
  }
  // (ct, ss)
  std::pair<Bytes, Bytes> EncapsulateDeterministically(const PublicKey& pk, const Bytes& seed) const {  // kyber.go:359-374
    if (seed.size() != EncapsulationSeedSize()) throw ErrSeedSize();
    if (pk.GetScheme() != this) throw ErrTypeMismatch();
    Bytes ct, ss;
    EncapsulateBatch(pk.MarshalBinary(), seed, ct, ss);
    return {std::move(ct), std::move(ss)};
  }
  std::pair<Bytes, Bytes> Encapsulate(const PublicKey& pk) const {  // kyber.go:348-357
    Bytes seed(EncapsulationSeedSize());
    fill_random(seed.data(), seed.size());
    return EncapsulateDeterministically(pk, seed);
  }
  Bytes Decapsulate(const PrivateKey& sk, const Bytes& ct) const {  // kyber.go:376-388
    if (sk.GetScheme() != this) throw ErrTypeMismatch();
    if (ct.size() != CiphertextSize()) throw ErrCiphertextSize();
    Bytes ss;
    DecapsulateBatch(sk.MarshalBinary(), ct, ss);
    return ss;
  }
  // ---- batch entry points (keys: one packed key = shared by the batch, or n keys back to back)
  void EncapsulateBatch(const Bytes& eks, const Bytes& seeds, Bytes& cts, Bytes& sss) const {
    const size_t es = EncapsulationSeedSize(), pks = PublicKeySize();
    if (seeds.size() % es) throw ErrSeedSize();
    const size_t n = seeds.size() / es;
    const bool shared = eks.size() == pks;
    if (!shared && eks.size() != n * pks) throw ErrPubKeySize();
    const size_t stride = shared ? 0 : pks;
    cts.resize(n * CiphertextSize());
    sss.resize(n * SharedKeySize());
    switch (kind_) {
      case MLKEM: check(cb200_mlkem_encaps(k_, eks.data(), stride, seeds.data(), cts.data(), sss.data(), nullptr, n)); break;
      case KYBER3: check(cb200_kyber_kem_encaps(k_, eks.data(), stride, seeds.data(), cts.data(), sss.data(), n)); break;
      case XWING: check(cb200_xwing_encaps(eks.data(), stride, seeds.data(), cts.data(), sss.data(), nullptr, n)); break;
      case HYBRID: check(cb200_hybrid_encaps(k_, eks.data(), stride, seeds.data(), cts.data(), sss.data(), nullptr, n)); break;
    }
  }
  void DecapsulateBatch(const Bytes& dks, const Bytes& cts, Bytes& sss) const {
    const size_t sks = PrivateKeySize();
    if (cts.size() % CiphertextSize()) throw ErrCiphertextSize();
    const size_t n = cts.size() / CiphertextSize();
    const bool shared = dks.size() == sks;
    if (!shared && dks.size() != n * sks) throw ErrPrivKeySize();
    const size_t stride = shared ? 0 : sks;
    sss.resize(n * SharedKeySize());
    switch (kind_) {
      case MLKEM: check(cb200_mlkem_decaps(k_, dks.data(), stride, cts.data(), sss.data(), nullptr, n)); break;
      case KYBER3: check(cb200_kyber_kem_decaps(k_, dks.data(), stride, cts.data(), sss.data(), n)); break;
      case XWING: check(cb200_xwing_decaps(dks.data(), stride, cts.data(), sss.data(), n)); break;
      case HYBRID: {
        if (!shared || n == 1) {
          check(cb200_hybrid_decaps(k_, dks.data(), sks, cts.data(), sss.data(), nullptr, n));
        } else {  // the hybrid decapsulation flow takes one key per operation
          Bytes rep(n * sks);
          for (size_t i = 0; i < n; i++) std::copy(dks.begin(), dks.end(), rep.begin() + i * sks);
          check(cb200_hybrid_decaps(k_, rep.data(), sks, cts.data(), sss.data(), nullptr, n));
        }
        break;
      }
    }
  }
  void DeriveKeyPairBatch(const Bytes& seeds, Bytes& eks, Bytes& dks) const {
    if (seeds.size() % SeedSize()) throw ErrSeedSize();
    const size_t n = seeds.size() / SeedSize();
    eks.resize(n * PublicKeySize());
    dks.resize(n * PrivateKeySize());
    switch (kind_) {
      case MLKEM: check(cb200_mlkem_keygen(k_, seeds.data(), eks.data(), dks.data(), n)); break;
      case KYBER3: check(cb200_kyber_kem_keygen(k_, seeds.data(), eks.data(), dks.data(), n)); break;
      case XWING:  // the packed private key is the seed itself (xwing.go:68-73)
        check(cb200_xwing_keygen(seeds.data(), eks.data(), n));
        dks = seeds;
        break;
      case HYBRID: check(cb200_hybrid_keygen(k_, seeds.data(), eks.data(), dks.data(), n)); break;
    }
  }
  int k() const { return k_; }
  Kind kind() const { return kind_; }

 private:
  static void check(int rc) {
    if (rc == 0) return;
    if (rc == CB200_ERR_PUBKEY) throw ErrPubKey();
    if (rc == CB200_ERR_PRIVKEY) throw ErrPrivKey();
    if (rc == CB200_ERR_ARG) throw std::logic_error(cb200_last_error());
    throw Error(cb200_last_error());
  }
  std::string name_;
  Kind kind_;
  int k_;  // K of the lattice scheme, or the hybrid identifier
};

inline PublicKey PrivateKey::Public() const {
  const Scheme* s = scheme_;
  if (s->kind() == Scheme::XWING) {  // sk is the seed: derive again (xwing.go:278-281)
    Bytes pk, sk;
    s->DeriveKeyPairBatch(packed_, pk, sk);
    return PublicKey(s, std::move(pk));
  }
  if (s->kind() == Scheme::HYBRID) {  // {first.Public(), second.Public()} (hybrid.go:151-153, xkem.go:68-84)
    const size_t dk = s->PrivateKeySize() - 32, ek = s->PublicKeySize() - 32, kk = (dk - 96) / 768;
    const bool x_first = s->k() != CB200_HYBRID_X25519MLKEM768;
    Bytes pk(s->PublicKeySize());
    const uint8_t* skm = packed_.data() + (x_first ? 32 : 0);
    const uint8_t* skx = packed_.data() + (x_first ? 0 : dk);
    std::copy(skm + 384 * kk, skm + 384 * kk + ek, pk.begin() + (x_first ? 32 : 0));
    if (cb200_x25519(skx, nullptr, pk.data() + (x_first ? 0 : ek), nullptr, 1)) throw Error(cb200_last_error());
    return PublicKey(s, std::move(pk));
  }
  const int k = s->k();
  return PublicKey(s, Bytes(packed_.begin() + 384 * k, packed_.begin() + 384 * k + 384 * k + 32));
}

inline std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return s;
}
inline const std::vector<const Scheme*>& All() {  // kem/schemes/schemes.go:75
  static const Scheme s512("ML-KEM-512", 2), s768("ML-KEM-768", 3), s1024("ML-KEM-1024", 4);
  static const Scheme k512("Kyber512", 2, true), k768("Kyber768", 3, true), k1024("Kyber1024", 4, true);
  static const Scheme xw("X-Wing", Scheme::XWING, 0), xm("X25519MLKEM768", Scheme::HYBRID, CB200_HYBRID_X25519MLKEM768),
      k7x("Kyber768-X25519", Scheme::HYBRID, CB200_HYBRID_KYBER768_X25519),
      k5x("Kyber512-X25519", Scheme::HYBRID, CB200_HYBRID_KYBER512_X25519);
  static const std::vector<const Scheme*> all = {&s512, &s768, &s1024, &k512, &k768, &k1024, &k5x, &k7x, &xm, &xw};
  return all;
}
inline const Scheme* ByName(const std::string& name) {  // kem/schemes/schemes.go:70 (nullptr = no such scheme)
  for (const Scheme* s : All())
    if (lower(s->Name()) == lower(name)) return s;
  return nullptr;
}

}  // namespace kem

namespace sign {

struct ErrContextTooLong : Error { ErrContextTooLong() : Error("sign: context string too long") {} };
struct ErrContextNotSupported : Error { ErrContextNotSupported() : Error("sign: context not supported") {} };
struct ErrPubKeySize : Error { ErrPubKeySize() : Error("sign: invalid public key size") {} };
struct ErrPrivKeySize : Error { ErrPrivKeySize() : Error("sign: invalid private key size") {} };
struct ErrSeedSize : Error { ErrSeedSize() : Error("sign: invalid seed size") {} };

struct SignatureOpts {  // sign/sign.go:14-18
  std::string Context;
};

class Scheme;
class PublicKey {
 public:
  PublicKey(const Scheme* s, Bytes b) : scheme_(s), packed_(std::move(b)) {}
  const Scheme* GetScheme() const { return scheme_; }
  const Bytes& MarshalBinary() const { return packed_; }
  bool Equal(const PublicKey& o) const { return packed_ == o.packed_; }

 private:
  const Scheme* scheme_;
  Bytes packed_;
};
class PrivateKey {
 public:
  PrivateKey(const Scheme* s, Bytes b) : scheme_(s), packed_(std::move(b)) {}
  const Scheme* GetScheme() const { return scheme_; }
  const Bytes& MarshalBinary() const { return packed_; }
  bool Equal(const PrivateKey& o) const { return packed_ == o.packed_; }

 private:
  const Scheme* scheme_;
  Bytes packed_;
};

class Scheme {  // sign.Scheme (sign/sign.go:48-94) for ML-DSA-44/65/87 (modes 44, 65, 87) and round-3 Dilithium2/3/5 (2, 3, 5)
 public:
  Scheme(std::string name, int mode) : name_(std::move(name)), mode_(mode) {}
  const std::string& Name() const { return name_; }
  size_t PublicKeySize() const { return cb200_mldsa_public_key_size(mode_); }
  size_t PrivateKeySize() const { return cb200_mldsa_private_key_size(mode_); }
  size_t SignatureSize() const { return cb200_mldsa_signature_size(mode_); }
  size_t SeedSize() const { return 32; }
  bool SupportsContext() const { return mode_ > 10; }  // round 3: sign.ErrContextNotSupported

  std::pair<PublicKey, PrivateKey> DeriveKey(const Bytes& seed) const {  // dilithium.go:266-276
    if (seed.size() != SeedSize()) throw std::logic_error("sign: invalid seed size");  // Go panics here
    Bytes pk(PublicKeySize()), sk(PrivateKeySize());
    check(cb200_mldsa_keygen(mode_, seed.data(), pk.data(), sk.data(), 1));
    return {PublicKey(this, std::move(pk)), PrivateKey(this, std::move(sk))};
  }
  std::pair<PublicKey, PrivateKey> GenerateKey() const {
    Bytes seed(SeedSize());
    fill_random(seed.data(), seed.size());
    return DeriveKey(seed);
  }
  PublicKey UnmarshalBinaryPublicKey(const Bytes& b) const {
    if (b.size() != PublicKeySize()) throw ErrPubKeySize();
    return PublicKey(this, b);
  }
  PrivateKey UnmarshalBinaryPrivateKey(const Bytes& b) const {
    if (b.size() != PrivateKeySize()) throw ErrPrivKeySize();
    return PrivateKey(this, b);
  }
  // deterministic signing, as sign.Scheme.Sign does (dilithium.go:282-303)
  Bytes Sign(const PrivateKey& sk, const Bytes& msg, const SignatureOpts* opts = nullptr) const {
    const std::string ctx = opts ? opts->Context : std::string();
    if (!ctx.empty() && !SupportsContext()) throw ErrContextNotSupported();
    if (ctx.size() > 255) throw ErrContextTooLong();
    Bytes sig(SignatureSize());
    const uint64_t off[2] = {0, msg.size()};
    Bytes padded(msg);
    padded.resize(msg.size() + 8);
    check(cb200_mldsa_sign(mode_, sk.MarshalBinary().data(), 0, padded.data(), off, (const uint8_t*)ctx.data(), ctx.size(), nullptr,
                             sig.data(), nullptr, 1, 0, nullptr));
    return sig;
  }
  bool Verify(const PublicKey& pk, const Bytes& msg, const Bytes& sig, const SignatureOpts* opts = nullptr) const {
    const std::string ctx = opts ? opts->Context : std::string();
    if (ctx.size() > 255 || sig.size() != SignatureSize()) return false;  // dilithium.go:116-118
    uint8_t ok = 0;
    const uint64_t off[2] = {0, msg.size()};
    Bytes padded(msg);
    padded.resize(msg.size() + 8);
    check(cb200_mldsa_verify(mode_, pk.MarshalBinary().data(), 0, padded.data(), off, (const uint8_t*)ctx.data(), ctx.size(),
                               sig.data(), &ok, 1, 0));
    return ok != 0;
  }
  // batch: messages back to back with n+1 offsets; sks: one key (shared) or n keys
  void SignBatch(const Bytes& sks, const Bytes& msgs, const std::vector<uint64_t>& off, const std::string& ctx, Bytes& sigs) const {
    if (!ctx.empty() && !SupportsContext()) throw ErrContextNotSupported();
    if (ctx.size() > 255) throw ErrContextTooLong();
    const size_t n = off.size() - 1;
    const bool shared = sks.size() == PrivateKeySize();
    if (!shared && sks.size() != n * PrivateKeySize()) throw ErrPrivKeySize();
    sigs.resize(n * SignatureSize());
    Bytes padded(msgs);
    padded.resize(msgs.size() + 8);
    check(cb200_mldsa_sign(mode_, sks.data(), shared ? 0 : PrivateKeySize(), padded.data(), off.data(), (const uint8_t*)ctx.data(),
                             ctx.size(), nullptr, sigs.data(), nullptr, n, 0, nullptr));
  }
  // ok[i] = 1 where signature i verifies; pks: one key (shared) or n keys
  void VerifyBatch(const Bytes& pks, const Bytes& msgs, const std::vector<uint64_t>& off, const std::string& ctx,
                   const Bytes& sigs, Bytes& ok) const {
    if (!ctx.empty() && !SupportsContext()) throw ErrContextNotSupported();
    const size_t n = off.size() - 1;
    const bool shared = pks.size() == PublicKeySize();
    if (!shared && pks.size() != n * PublicKeySize()) throw ErrPubKeySize();
    ok.assign(n, 0);
    if (ctx.size() > 255 || sigs.size() != n * SignatureSize()) return;  // dilithium.go:116-118: such signatures are invalid
    Bytes padded(msgs);
    padded.resize(msgs.size() + 8);
    check(cb200_mldsa_verify(mode_, pks.data(), shared ? 0 : PublicKeySize(), padded.data(), off.data(), (const uint8_t*)ctx.data(),
                               ctx.size(), sigs.data(), ok.data(), n, 0));
  }
  void DeriveKeyBatch(const Bytes& seeds, Bytes& pks, Bytes& sks) const {
    if (seeds.size() % SeedSize()) throw ErrSeedSize();
    const size_t n = seeds.size() / SeedSize();
    pks.resize(n * PublicKeySize());
    sks.resize(n * PrivateKeySize());
    check(cb200_mldsa_keygen(mode_, seeds.data(), pks.data(), sks.data(), n));
  }

 private:
  static void check(int rc) {
    if (rc == 0) return;
    if (rc == CB200_ERR_ARG) throw std::logic_error(cb200_last_error());
    throw Error(cb200_last_error());  // includes CB200_ERR_SIGN_ATTEMPTS (the reference panics after 576 attempts)
  }
  std::string name_;
  int mode_;
};

inline const std::vector<const Scheme*>& All() {
  static const Scheme s44("ML-DSA-44", 44), s65("ML-DSA-65", 65), s87("ML-DSA-87", 87);
  static const Scheme d2("Dilithium2", 2), d3("Dilithium3", 3), d5("Dilithium5", 5);
  static const std::vector<const Scheme*> all = {&s44, &s65, &s87, &d2, &d3, &d5};
  return all;
}
inline const Scheme* ByName(const std::string& name) {  // sign/schemes/schemes.go:69 (nullptr = no such scheme)
  for (const Scheme* s : All())
    if (kem::lower(s->Name()) == kem::lower(name)) return s;
  return nullptr;
}

}  // namespace sign
}  // namespace circl
