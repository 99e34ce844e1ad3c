"""GPU replay of the Wycheproof ML-DSA vectors (sign/schemes/wycheproof_test.go): every signing case
of a file with the same context string goes through ONE batched call with per-operation keys -- so the cases flagged
ManySteps (dozens of rejection-loop iterations) share a batch with ordinary ones -- and every verification group through
one batched call per context.  Both through the C ABI."""
import numpy as np
import pytest

from wycheproof_replay import replay_file

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


class GpuBackend:
    def __init__(self, name):
        from circl_b200 import mldsa
        self.s = mldsa.ByName(name)
        self.sk_size, self.pk_size, self.sig_size = self.s.PrivateKeySize(), self.s.PublicKeySize(), self.s.SignatureSize()

    def derive(self, seed):
        return self.s.DeriveKey(seed)[1].MarshalBinary()

    def sign_many(self, sks, msgs, ctx):
        arr = np.stack([np.frombuffer(sk, dtype=np.uint8) for sk in sks])
        return [row.tobytes() for row in self.s.SignBatch(arr, msgs, ctx=ctx)]

    def verify_many(self, pk, msgs, sigs, ctx):
        arr = np.stack([np.frombuffer(s, dtype=np.uint8) for s in sigs])
        return self.s.VerifyBatch(self.s.UnmarshalBinaryPublicKey(pk), msgs, arr, ctx=ctx).tolist()


@pytest.mark.parametrize("name", ["mldsa_44_sign_noseed_test", "mldsa_44_sign_seed_test", "mldsa_44_verify_test",
                                  "mldsa_65_noseed_sign_test", "mldsa_65_seed_sign_test", "mldsa_65_verify_test",
                                  "mldsa_87_sign_noseed_test", "mldsa_87_sign_seed_test", "mldsa_87_verify_test"])
def test_wycheproof_file(cb, mldsa_wycheproof, name):
    entry = mldsa_wycheproof[name]
    ns, nv = replay_file(entry, GpuBackend(entry["algorithm"]))
    assert ns + nv >= 50


def test_scheme_wrapper_rejects_what_the_reference_wrapper_rejects(cb):
    # sign/mldsa/mldsa65/dilithium.go:56-70,337-349: context > 255 bytes, wrong key sizes
    from circl_b200 import mldsa
    s = mldsa.ByName("ML-DSA-65")
    pk, sk = s.DeriveKey(bytes(32))
    with pytest.raises(mldsa.ErrContextTooLong):
        s.Sign(sk, b"m", mldsa.SignatureOpts(Context=bytes(256)))
    assert s.Verify(pk, b"m", s.Sign(sk, b"m"), mldsa.SignatureOpts(Context=bytes(256))) is False
    with pytest.raises(mldsa.ErrPrivKeySize):
        s.UnmarshalBinaryPrivateKey(bytes(4031))
    with pytest.raises(mldsa.ErrPubKeySize):
        s.UnmarshalBinaryPublicKey(bytes(1953))
