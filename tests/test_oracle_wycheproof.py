"""Pins the oracle against the Wycheproof ML-DSA vectors the reference replays (sign/schemes/wycheproof_test.go,
sign/schemes/testdata/wycheproof/mldsa_{44,65,87}_*): malformed keys and signatures, contexts,
hint encodings, signatures that need many rejection-loop iterations.  CPU only."""
import pytest

import oracle
from wycheproof_replay import replay_file

MODES = {"ML-DSA-44": 44, "ML-DSA-65": 65, "ML-DSA-87": 87}


class OracleBackend:
    def __init__(self, mode):
        self.mode = mode
        self.sk_size, self.pk_size, self.sig_size = oracle.mldsa_sizes(mode)[1], oracle.mldsa_sizes(mode)[0], oracle.mldsa_sizes(mode)[2]

    def derive(self, seed):
        return oracle.mldsa_keygen(self.mode, seed)[1]

    def sign_many(self, sks, msgs, ctx):
        return [oracle.mldsa_sign(self.mode, sk, m, ctx=ctx)[0] for sk, m in zip(sks, msgs)]

    def verify_many(self, pk, msgs, sigs, ctx):
        return [oracle.mldsa_verify(self.mode, pk, m, s, ctx=ctx) for m, s in zip(msgs, sigs)]


def test_sizes_helper_order():
    pk, sk, sig = oracle.mldsa_sizes(65)
    assert (pk, sk, sig) == (1952, 4032, 3309)


@pytest.mark.parametrize("name", ["mldsa_44_sign_noseed_test", "mldsa_44_sign_seed_test", "mldsa_44_verify_test",
                                  "mldsa_65_noseed_sign_test", "mldsa_65_seed_sign_test", "mldsa_65_verify_test",
                                  "mldsa_87_sign_noseed_test", "mldsa_87_sign_seed_test", "mldsa_87_verify_test"])
def test_wycheproof_file(mldsa_wycheproof, name):
    entry = mldsa_wycheproof[name]
    ns, nv = replay_file(entry, OracleBackend(MODES[entry["algorithm"]]))
    assert ns + nv >= 50
