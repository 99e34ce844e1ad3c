"""AES-256-CTR DRBG of the NIST PQC KAT harness (test helper).

Behavioural restatement of internal/nist/drbg.go:43-80 using the
`cryptography` package's AES (test infrastructure only).
"""
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes


class DRBG:
    def __init__(self, seed48: bytes):
        self.key = bytes(32)
        self.v = bytes(16)
        self._update(seed48)

    def _inc(self):
        self.v = ((int.from_bytes(self.v, "big") + 1) % (1 << 128)).to_bytes(16, "big")

    def _block(self) -> bytes:
        enc = Cipher(algorithms.AES(self.key), modes.ECB()).encryptor()
        return enc.update(self.v)

    def _update(self, pd):
        buf = b""
        for _ in range(3):
            self._inc()
            buf += self._block()
        if pd is not None:
            buf = bytes(a ^ b for a, b in zip(buf, pd))
        self.key, self.v = buf[:32], buf[32:]

    def fill(self, n: int) -> bytes:
        out = b""
        while len(out) < n:
            self._inc()
            out += self._block()
        self._update(None)
        return out[:n]
