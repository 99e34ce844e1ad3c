"""GPU tests of the runtime behind the C ABI (SURVEY.md 8(b) "Threading", 8(e)):

  * concurrent callers: kem.Scheme / sign.Scheme are goroutine-safe in the reference (stateless singletons,
    kem/mlkem/mlkem768/kyber.go:269); here several threads call the same entry points at once, on host pointers and on
    device pointers with one CUDA stream per thread, and every result must equal the single-threaded one -- including
    the per-call error state (kem.ErrPubKey must be reported to the caller that passed the bad key, to nobody else);
  * one process, several GPUs: cb200_init_devices shards a host-pointer batch by index inside the library and the
    caller's buffers hold the results in index order (skipped on a one-GPU box).
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def _keys(n, seed=0):
    from circl_b200 import mlkem
    rng = np.random.default_rng(seed)
    return mlkem.ByName("ML-KEM-768").DeriveKeyPairBatch(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))


def test_concurrent_host_callers_keep_their_own_results_and_errors(cb):
    from circl_b200 import _ffi, mlkem
    L = _ffi.lib()
    n = 1 << 13
    eks, _ = _keys(n)
    rng = np.random.default_rng(5)
    seeds = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    scheme = mlkem.ByName("ML-KEM-768")
    want_ct, want_ss = scheme.EncapsulateBatch(eks, seeds)
    bad = eks.copy()
    bad[7, :2] = 0xFF  # coefficient 0xfff >= q: kem.ErrPubKey for op 7 only (cpapke.go:48-54)
    results = {}

    def good(tag, size):
        ct = np.empty((size, 1088), dtype=np.uint8)
        ss = np.empty((size, 32), dtype=np.uint8)
        for _ in range(4):
            rc = L.cb200_mlkem_encaps(3, eks.ctypes.data, 1184, seeds.ctypes.data, ct.ctypes.data, ss.ctypes.data, None, size)
            if rc != 0:
                results[tag] = ("rc", rc, L.cb200_last_error())
                return
        results[tag] = (np.array_equal(ct, want_ct[:size]), np.array_equal(ss, want_ss[:size]))

    def faulty(tag):
        ct = np.empty((n, 1088), dtype=np.uint8)
        ss = np.empty((n, 32), dtype=np.uint8)
        st = np.zeros(n, dtype=np.uint8)
        ok = True
        for _ in range(4):
            rc = L.cb200_mlkem_encaps(3, bad.ctypes.data, 1184, seeds.ctypes.data, ct.ctypes.data, ss.ctypes.data,
                                      st.ctypes.data, n)
            ok &= rc == -3 and b"kem.ErrPubKey" in L.cb200_last_error()
            ok &= int(st.sum()) == 1 and st[7] == 1 and not ct[7].any() and np.array_equal(ct[8], want_ct[8])
        results[tag] = (ok, ok)

    threads = [threading.Thread(target=good, args=("a", n)), threading.Thread(target=good, args=("b", 100)),
               threading.Thread(target=faulty, args=("c",)), threading.Thread(target=good, args=("d", n // 2 + 3))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert results == {k: (True, True) for k in "abcd"}, results


def test_concurrent_device_callers_on_their_own_streams(cb):
    import torch
    from circl_b200 import _ffi
    L = _ffi.lib()
    n = 1 << 12
    eks, _ = _keys(n, 1)
    seeds = np.random.default_rng(6).integers(0, 256, size=(n, 32), dtype=np.uint8)
    from circl_b200 import mlkem
    want_ct, want_ss = mlkem.ByName("ML-KEM-768").EncapsulateBatch(eks, seeds)
    eks_d, seeds_d = torch.from_numpy(eks).cuda(), torch.from_numpy(seeds).cuda()
    torch.cuda.synchronize()
    out = {}

    def worker(tag, size):
        stream = torch.cuda.Stream()
        ct = torch.empty((size, 1088), dtype=torch.uint8, device="cuda")
        ss = torch.empty((size, 32), dtype=torch.uint8, device="cuda")
        st = torch.zeros((size,), dtype=torch.uint8, device="cuda")
        L.cb200_set_stream(stream.cuda_stream)  # per thread
        for _ in range(6):
            rc = L.cb200_mlkem_encaps(3, eks_d.data_ptr(), 1184, seeds_d.data_ptr(), ct.data_ptr(), ss.data_ptr(),
                                      st.data_ptr(), size)
            assert rc == 0, L.cb200_last_error()
        stream.synchronize()
        out[tag] = (np.array_equal(ct.cpu().numpy(), want_ct[:size]) and np.array_equal(ss.cpu().numpy(), want_ss[:size])
                    and int(st.sum().item()) == 0)
        L.cb200_release_stream(stream.cuda_stream)

    threads = [threading.Thread(target=worker, args=(i, n - 17 * i)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert out == {i: True for i in range(4)}, out


def test_misaligned_device_pointers_are_an_argument_error_not_a_sticky_fault(cb):
    # a CUDA tensor view with an odd storage offset must come back as CB200_ERR_ARG
    import torch
    from circl_b200 import _ffi
    L = _ffi.lib()
    buf = torch.zeros(64 * 1216 + 64, dtype=torch.uint8, device="cuda")
    odd = buf[1:]
    out = torch.zeros(64 * 32, dtype=torch.uint8, device="cuda")
    assert L.cb200_x25519(odd.data_ptr(), None, out.data_ptr(), None, 64) == -1
    assert L.cb200_xwing_keygen(odd.data_ptr(), buf.data_ptr(), 16) == -1
    assert L.cb200_kyber_ntt(odd.data_ptr(), 4, 0) == -1
    # the context is still healthy
    p = torch.zeros((4, 256), dtype=torch.int16, device="cuda")
    assert L.cb200_kyber_ntt(p.data_ptr(), 4, 0) == 0
    torch.cuda.synchronize()


def test_one_process_many_gpus_shards_host_batches_by_index():
    import circl_b200
    from circl_b200 import kyber, mldsa, mlkem
    import oracle
    if circl_b200.device_count() < 2:
        pytest.skip("needs at least two GPUs in this process")
    try:
        ndev = circl_b200.init_devices(0)
        assert ndev == circl_b200.device_count() >= 2
        rng = np.random.default_rng(8)
        n = (1 << 16) + 11
        scheme = mlkem.ByName("ML-KEM-1024")
        pool, _ = scheme.DeriveKeyPairBatch(rng.integers(0, 256, size=(64, 64), dtype=np.uint8))
        eks = np.ascontiguousarray(pool[np.arange(n) % 64])
        seeds = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        ct, ss = scheme.EncapsulateBatch(eks, seeds)
        for i in list(range(0, n, 4099)) + [n - 1]:
            wct, wss = oracle.mlkem_encaps(4, eks[i].tobytes(), seeds[i].tobytes())
            assert ct[i].tobytes() == wct and ss[i].tobytes() == wss, i
        # a bad key in the last shard is reported once, at its own index
        eks[n - 2, :2] = 0xFF
        with pytest.raises(mlkem.ErrPubKey) as ei:
            scheme.EncapsulateBatch(eks, seeds)
        st = ei.value.status
        assert int(st.sum()) == 1 and st[n - 2] == 1
        # raw ring op and a signature batch through the same sharding
        p = rng.integers(-3329, 3329, size=(1 << 17, 256), dtype=np.int64).astype(np.int16)
        assert np.array_equal(kyber.ntt_(p.copy())[::1031], oracle.kyber_ntt(p[::1031]))
        sch = mldsa.ByName("ML-DSA-65")
        pk, sk = sch.DeriveKeyBatch(rng.integers(0, 256, size=(4, 32), dtype=np.uint8))
        m = 1 << 14
        msgs = [bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8)) for _ in range(m)]
        sks = np.ascontiguousarray(sk[np.arange(m) % 4])
        sigs = sch.SignBatch(sks, msgs)
        for i in (0, 4095, 4096, 8191, 8192, m - 1):
            want, _ = oracle.mldsa65_sign(sks[i].tobytes(), msgs[i])
            assert sigs[i].tobytes() == want, i
        ok = sch.VerifyBatch(np.ascontiguousarray(pk[np.arange(m) % 4]), msgs, sigs)
        assert bool(ok.all())
    finally:
        circl_b200.shutdown()
