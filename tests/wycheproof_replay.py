"""Replay of sign/schemes/wycheproof_test.go:40-150 over any sign.Scheme-shaped backend (the oracle on the CPU, the CUDA
path on the GPU).  A backend offers

    derive(seed32) -> sk bytes                     scheme.DeriveKey
    sk_size / pk_size / sig_size
    sign_many(sks, msgs, ctx) -> list of sig bytes  (same context for the whole batch)
    verify_many(pk, msgs, sigs, ctx) -> list of bool (one key for the whole batch)

Key and signature length checks and the 255-byte context limit live in the reference's scheme wrapper
(sign/mldsa/mldsa65/dilithium.go:56-70,115-118,337-349); they are mirrored here the same way for both backends.
"""
from collections import defaultdict

SKIP = ("private key with s1 vector out of range", "private key with s2 vector out of range")  # wycheproof_test.go:81-86


def replay_file(entry, be):
    """Returns (checked sign cases, checked verify cases)."""
    n_sign = n_verify = 0
    sign_jobs = defaultdict(list)  # ctx -> [(sk, msg, want sig, tcId)]
    for g in entry["groups"]:
        if g["type"] == "MlDsaSign":
            assert ("privateKey" in g) != ("privateSeed" in g) and "publicKey" not in g
            sk = None
            if "privateSeed" in g:
                sk = be.derive(bytes.fromhex(g["privateSeed"]))
            else:
                raw = bytes.fromhex(g["privateKey"])
                sk = raw if len(raw) == be.sk_size else None  # UnmarshalBinaryPrivateKey
            for t in g["tests"]:
                if t["comment"] in SKIP:
                    continue
                if sk is None:
                    assert t["result"] == "invalid", t["tcId"]  # a key that does not parse only carries invalid cases
                    n_sign += 1
                    continue
                ctx = bytes.fromhex(t["ctx"] or "")
                if t["result"] == "invalid":
                    assert len(ctx) > 255, (t["tcId"], t["comment"])  # the only way Sign fails with a parsed key
                    n_sign += 1
                    continue
                sign_jobs[ctx].append((sk, bytes.fromhex(t["msg"]), bytes.fromhex(t["sig"]), t["tcId"]))
        elif g["type"] == "MlDsaVerify":
            assert "privateKey" not in g and "privateSeed" not in g
            raw = bytes.fromhex(g["publicKey"])
            pk = raw if len(raw) == be.pk_size else None  # UnmarshalBinaryPublicKey
            by_ctx = defaultdict(list)
            for t in g["tests"]:
                if pk is None:
                    assert t["result"] == "invalid", t["tcId"]
                    n_verify += 1
                    continue
                ctx, sig = bytes.fromhex(t["ctx"] or ""), bytes.fromhex(t["sig"])
                if len(ctx) > 255 or len(sig) != be.sig_size:  # scheme.Verify returns false before the lattice code
                    assert t["result"] == "invalid", t["tcId"]
                    n_verify += 1
                    continue
                by_ctx[ctx].append((bytes.fromhex(t["msg"]), sig, t["result"] == "valid", t["tcId"]))
            for ctx, items in by_ctx.items():
                got = be.verify_many(pk, [m for m, _, _, _ in items], [s for _, s, _, _ in items], ctx)
                for ok, (_, _, want, tc) in zip(got, items):
                    assert bool(ok) == want, ("verify", tc)
                    n_verify += 1
        else:
            raise AssertionError(g["type"])
    for ctx, items in sign_jobs.items():
        sigs = be.sign_many([sk for sk, _, _, _ in items], [m for _, m, _, _ in items], ctx)
        for sig, (_, _, want, tc) in zip(sigs, items):
            assert bytes(sig) == want, ("sign", tc)
            n_sign += 1
    return n_sign, n_verify
