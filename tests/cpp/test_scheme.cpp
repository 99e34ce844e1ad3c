// tests/cpp/test_scheme.cpp -- exercises include/circl_b200.hpp (the compiled host-side mirror of kem.Scheme /
// sign.Scheme) the way kem/schemes/schemes_test.go:53 and sign/schemes/schemes_test.go:17 exercise the Go API.
// Prints hex transcripts for fixed seeds; tests/test_gpu_cpp_host.py compares them with the oracle.
#include <cstdio>
#include <cstdlib>

#include "circl_b200.hpp"

using namespace circl;

static void hex(const char* tag, const Bytes& b) {
  printf("%s=", tag);
  for (uint8_t x : b) printf("%02x", x);
  printf("\n");
}
#define REQUIRE(c)                                             \
  do {                                                         \
    if (!(c)) {                                                \
      fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); \
      return 1;                                                \
    }                                                          \
  } while (0)

int main() {
  init(0);
  for (const char* name : {"ML-KEM-512", "ml-kem-768", "ML-KEM-1024"}) {
    const kem::Scheme* s = kem::ByName(name);
    REQUIRE(s != nullptr);
    Bytes seed(s->SeedSize()), eseed(s->EncapsulationSeedSize());
    for (size_t i = 0; i < seed.size(); i++) seed[i] = (uint8_t)(i * 7 + s->k());
    for (size_t i = 0; i < eseed.size(); i++) eseed[i] = (uint8_t)(255 - i);
    auto kp = s->DeriveKeyPair(seed);
    REQUIRE(kp.first.MarshalBinary().size() == s->PublicKeySize());
    REQUIRE(kp.second.Public().Equal(kp.first));
    auto enc = s->EncapsulateDeterministically(kp.first, eseed);
    REQUIRE(enc.first.size() == s->CiphertextSize() && enc.second.size() == s->SharedKeySize());
    REQUIRE(s->Decapsulate(kp.second, enc.first) == enc.second);
    Bytes bad = enc.first;
    bad[3] ^= 1;
    REQUIRE(s->Decapsulate(kp.second, bad) != enc.second);  // implicit rejection
    printf("scheme=%s\n", s->Name().c_str());
    hex("ek", kp.first.MarshalBinary());
    hex("dk", kp.second.MarshalBinary());
    hex("ct", enc.first);
    hex("ss", enc.second);
    hex("ss_rejected", s->Decapsulate(kp.second, bad));
    // error behaviour
    bool threw = false;
    try { s->UnmarshalBinaryPublicKey(Bytes(10)); } catch (const kem::ErrPubKeySize&) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { s->EncapsulateDeterministically(kp.first, Bytes(31)); } catch (const kem::ErrSeedSize&) { threw = true; }
    REQUIRE(threw);
    threw = false;
    Bytes noncanon = kp.first.MarshalBinary();
    noncanon[0] = 0xff;
    noncanon[1] |= 0x0f;
    try { s->EncapsulateDeterministically(s->UnmarshalBinaryPublicKey(noncanon), eseed); } catch (const kem::ErrPubKey&) { threw = true; }
    REQUIRE(threw);
    // batch of 300 with a shared key equals 300 single calls on the first and last
    Bytes seeds(32 * 300), cts, sss;
    for (size_t i = 0; i < seeds.size(); i++) seeds[i] = (uint8_t)(i * 13);
    s->EncapsulateBatch(kp.first.MarshalBinary(), seeds, cts, sss);
    auto one = s->EncapsulateDeterministically(kp.first, Bytes(seeds.end() - 32, seeds.end()));
    REQUIRE(Bytes(cts.end() - s->CiphertextSize(), cts.end()) == one.first);
    Bytes back;
    s->DecapsulateBatch(kp.second.MarshalBinary(), cts, back);
    REQUIRE(back == sss);
  }
  REQUIRE(kem::ByName("no-such-kem") == nullptr);

  const sign::Scheme* d = sign::ByName("ML-DSA-65");
  REQUIRE(d != nullptr && d->SupportsContext());
  Bytes dseed(32);
  for (size_t i = 0; i < 32; i++) dseed[i] = (uint8_t)(3 * i + 1);
  auto dk = d->DeriveKey(dseed);
  Bytes msg = {'h', 'e', 'l', 'l', 'o'};
  sign::SignatureOpts opts{"ctx"};
  Bytes sig = d->Sign(dk.second, msg, &opts);
  REQUIRE(sig.size() == d->SignatureSize());
  REQUIRE(d->Verify(dk.first, msg, sig, &opts));
  REQUIRE(!d->Verify(dk.first, msg, sig));  // wrong context
  Bytes tam = sig;
  tam[100] ^= 4;
  REQUIRE(!d->Verify(dk.first, msg, tam, &opts));
  REQUIRE(!d->Verify(dk.first, msg, Bytes(sig.begin(), sig.end() - 1), &opts));
  bool threw = false;
  sign::SignatureOpts longctx{std::string(256, 'x')};
  try { d->Sign(dk.second, msg, &longctx); } catch (const sign::ErrContextTooLong&) { threw = true; }
  REQUIRE(threw);
  for (const char* other : {"ML-DSA-44", "ml-dsa-87"}) {  // the other parameter sets: sign/verify round trip
    const sign::Scheme* o = sign::ByName(other);
    REQUIRE(o != nullptr);
    auto ok2 = o->DeriveKey(dseed);
    Bytes s2 = o->Sign(ok2.second, msg);
    REQUIRE(s2.size() == o->SignatureSize() && o->Verify(ok2.first, msg, s2));
  }
  {  // batch entry points of sign.Scheme: 40 keys, messages of growing length, one tampered signature
    const size_t nb = 40;
    Bytes seeds(32 * nb), pks, sks, msgs, sigs, okv;
    for (size_t i = 0; i < seeds.size(); i++) seeds[i] = (uint8_t)(i * 17 + 5);
    d->DeriveKeyBatch(seeds, pks, sks);
    REQUIRE(Bytes(pks.begin(), pks.begin() + d->PublicKeySize()) ==
            d->DeriveKey(Bytes(seeds.begin(), seeds.begin() + 32)).first.MarshalBinary());
    std::vector<uint64_t> off(nb + 1, 0);
    for (size_t i = 0; i < nb; i++) {
      for (size_t j = 0; j <= i; j++) msgs.push_back((uint8_t)(i + 3 * j));
      off[i + 1] = msgs.size();
    }
    d->SignBatch(sks, msgs, off, "batch", sigs);
    sigs[5 * d->SignatureSize() + 77] ^= 1;
    d->VerifyBatch(pks, msgs, off, "batch", sigs, okv);
    for (size_t i = 0; i < nb; i++) REQUIRE((okv[i] != 0) == (i != 5));
  }
  REQUIRE(sign::ByName("ML-DSA-99") == nullptr);
  // round-3 wrappers (kem/kyber/kyber768, sign/dilithium/mode3): same classes, other entry points / parameter sets
  for (const char* name : {"Kyber512", "kyber768", "Kyber1024"}) {
    const kem::Scheme* s = kem::ByName(name);
    REQUIRE(s != nullptr);
    Bytes seed(s->SeedSize()), eseed(s->EncapsulationSeedSize());
    for (size_t i = 0; i < seed.size(); i++) seed[i] = (uint8_t)(i * 11 + s->k());
    for (size_t i = 0; i < eseed.size(); i++) eseed[i] = (uint8_t)(i ^ 0x5a);
    auto kp = s->DeriveKeyPair(seed);
    auto enc = s->EncapsulateDeterministically(kp.first, eseed);
    REQUIRE(s->Decapsulate(kp.second, enc.first) == enc.second);
    printf("scheme=%s\n", s->Name().c_str());
    hex("ek", kp.first.MarshalBinary());
    hex("dk", kp.second.MarshalBinary());
    hex("ct", enc.first);
    hex("ss", enc.second);
  }
  for (const char* name : {"X-Wing", "X25519MLKEM768", "kyber768-x25519", "Kyber512-X25519"}) {
    const kem::Scheme* s = kem::ByName(name);
    REQUIRE(s != nullptr);
    Bytes seed(s->SeedSize()), eseed(s->EncapsulationSeedSize());
    for (size_t i = 0; i < seed.size(); i++) seed[i] = (uint8_t)(i * 5 + 3);
    for (size_t i = 0; i < eseed.size(); i++) eseed[i] = (uint8_t)(i * 9 + 1);
    auto kp = s->DeriveKeyPair(seed);
    REQUIRE(kp.first.MarshalBinary().size() == s->PublicKeySize() && kp.second.MarshalBinary().size() == s->PrivateKeySize());
    REQUIRE(kp.second.Public().Equal(kp.first));
    auto enc = s->EncapsulateDeterministically(kp.first, eseed);
    REQUIRE(enc.first.size() == s->CiphertextSize() && enc.second.size() == s->SharedKeySize());
    REQUIRE(s->Decapsulate(kp.second, enc.first) == enc.second);
    // batch with one shared key: last element equals the single call
    Bytes seeds(s->EncapsulationSeedSize() * 40), cts, sss, back;
    for (size_t i = 0; i < seeds.size(); i++) seeds[i] = (uint8_t)(i * 31 + 7);
    s->EncapsulateBatch(kp.first.MarshalBinary(), seeds, cts, sss);
    s->DecapsulateBatch(kp.second.MarshalBinary(), cts, back);
    REQUIRE(back == sss);
    if (s->kind() == kem::Scheme::HYBRID) {  // kem/hybrid/xkem_test.go: small-order X25519 share -> kem.ErrPubKey
      const uint8_t low[32] = {0xe0, 0xeb, 0x7a, 0x7c, 0x3b, 0x41, 0xb8, 0xae, 0x16, 0x56, 0xe3, 0xfa, 0xf1, 0x9f, 0xc4, 0x6a,
                               0xda, 0x09, 0x8d, 0xeb, 0x9c, 0x32, 0xb1, 0xfd, 0x86, 0x62, 0x05, 0x16, 0x5f, 0x49, 0xb8, 0x00};
      const bool x_first = s->k() != CB200_HYBRID_X25519MLKEM768;
      Bytes badpk = kp.first.MarshalBinary();
      std::copy(low, low + 32, x_first ? badpk.begin() : badpk.end() - 32);
      bool threw2 = false;
      try { s->EncapsulateDeterministically(s->UnmarshalBinaryPublicKey(badpk), eseed); } catch (const kem::ErrPubKey&) { threw2 = true; }
      REQUIRE(threw2);
      Bytes badct = enc.first;
      std::copy(low, low + 32, x_first ? badct.begin() : badct.end() - 32);
      threw2 = false;
      try { s->Decapsulate(kp.second, badct); } catch (const kem::ErrPubKey&) { threw2 = true; }
      REQUIRE(threw2);
    }
    printf("scheme=%s\n", s->Name().c_str());
    hex("pk", kp.first.MarshalBinary());
    hex("sk", kp.second.MarshalBinary());
    hex("ct", enc.first);
    hex("ss", enc.second);
  }
  {
    const sign::Scheme* r3 = sign::ByName("dilithium3");
    REQUIRE(r3 != nullptr && !r3->SupportsContext() && r3->SignatureSize() == 3293);
    auto k3 = r3->DeriveKey(dseed);
    Bytes s3 = r3->Sign(k3.second, msg);
    REQUIRE(r3->Verify(k3.first, msg, s3));
    threw = false;
    try { r3->Sign(k3.second, msg, &opts); } catch (const sign::ErrContextNotSupported&) { threw = true; }
    REQUIRE(threw);
    printf("scheme=%s\n", r3->Name().c_str());
    hex("pk", k3.first.MarshalBinary());
    hex("sk", k3.second.MarshalBinary());
    hex("sig", s3);
  }
  printf("scheme=%s\n", d->Name().c_str());
  hex("pk", dk.first.MarshalBinary());
  hex("sk", dk.second.MarshalBinary());
  hex("sig", sig);
  printf("ALL OK\n");
  cb200_shutdown();
  return 0;
}
