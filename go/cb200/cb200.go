//go:build cuda && cgo

// Package cb200 is the cgo binding of libcirclb200.so (include/circl_b200.h):
// the B200 batch engine for CIRCL's module-lattice hot path.
//
// NOTE: this file is delivered as source.  The build image has no Go toolchain
// (`go version` fails), so it is reviewed, not compiled, here; every executable
// test and timing drives the same C ABI from Python/C (see INTEGRATION.md).
package cb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../circl_b200 -lcirclb200 -Wl,-rpath,${SRCDIR}/../../circl_b200
#include "circl_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"sync"
	"unsafe"
)

// Errors mirror the negative codes of the C ABI.
var (
	ErrNotInit = errors.New("cb200: not initialised / no CUDA device")
	ErrPubKey  = errors.New("cb200: encapsulation key is not canonical") // mapped to kem.ErrPubKey by callers
	ErrPrivKey = errors.New("cb200: H(ek) stored in the decapsulation key does not match") // kem.ErrPrivKey
)

var initOnce sync.Once
var initErr error

// Init drives every visible GPU from this process (cb200_init_devices(0)): host-slice batches are cut into one
// contiguous index range per GPU inside the library and the results land in the caller's slices in index order, so a
// kem.Scheme / sign.Scheme caller never sees devices.  Safe to call many times and from any goroutine.
func Init() error { return InitDevices(0) }

// InitDevices restricts the library to GPUs 0..ndev-1 (ndev <= 0: all).
func InitDevices(ndev int) error {
	initOnce.Do(func() {
		initErr = call(func() C.int { return C.cb200_init_devices(C.int(ndev)) })
	})
	return initErr
}

// ActiveDevices reports how many GPUs the library drives (0 before Init).
func ActiveDevices() int { return int(C.cb200_active_devices()) }

// call runs one C entry point and turns its return code into an error.  cb200_last_error() is per OS thread and a
// goroutine may migrate between two cgo calls, so the call and the error fetch are pinned to one thread.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return lastErr(f())
}

// lastErr must run on the OS thread that made the failing call (see call).
func lastErr(rc C.int) error {
	switch rc {
	case 0:
		return nil
	case C.CB200_ERR_PUBKEY:
		return ErrPubKey
	case C.CB200_ERR_PRIVKEY:
		return ErrPrivKey
	case C.CB200_ERR_NOT_INIT:
		return ErrNotInit
	case C.CB200_ERR_ARG:
		// programmer error: the reference panics on wrong lengths (kem/mlkem/mlkem768/kyber.go:110-121)
		panic(C.GoString(C.cb200_last_error()))
	}
	return errors.New(C.GoString(C.cb200_last_error()))
}

// KyberNTT runs (*Poly).NTT (inverse=false) or InvNTT on every polynomial, in place.
// Replaces nttAVX2 / invNttAVX2 (pke/kyber/internal/common/stubs_amd64.go:8-14) for batches.
// The slice is caller-owned; the library does not retain the pointer past the call (cgo rule).
func KyberNTT(polys [][256]int16, inverse bool) error {
	if len(polys) == 0 {
		return nil
	}
	inv := C.int(0)
	if inverse {
		inv = 1
	}
	return call(func() C.int {
		return C.cb200_kyber_ntt((*C.int16_t)(unsafe.Pointer(&polys[0])), C.size_t(len(polys)), inv)
	})
}

// KyberMulHat: p[i] = a[i] (*) b[i]  (mulHatAVX2, stubs_amd64.go:17).
func KyberMulHat(p, a, b [][256]int16) error {
	if len(p) != len(a) || len(p) != len(b) {
		panic("cb200: length mismatch")
	}
	if len(p) == 0 {
		return nil
	}
	return call(func() C.int {
		return C.cb200_kyber_mulhat((*C.int16_t)(unsafe.Pointer(&p[0])), (*C.int16_t)(unsafe.Pointer(&a[0])),
		(*C.int16_t)(unsafe.Pointer(&b[0])), C.size_t(len(p)))
	})
}

// DilithiumNTT: (*Poly).NTT / InvNTT of sign/internal/dilithium over a batch, in place.
func DilithiumNTT(polys [][256]uint32, inverse bool) error {
	if len(polys) == 0 {
		return nil
	}
	inv := C.int(0)
	if inverse {
		inv = 1
	}
	return call(func() C.int {
		return C.cb200_dil_ntt((*C.uint32_t)(unsafe.Pointer(&polys[0])), C.size_t(len(polys)), inv)
	})
}

// MLKEMEncaps: batched UnmarshalBinaryPublicKey + EncapsulateDeterministically.
// k = 3 (ML-KEM-768) or 4 (ML-KEM-1024).  ek holds either one packed key (shared) or
// n keys back to back; seeds is n*32 bytes.  ct and ss are written in place.
func MLKEMEncaps(k int, ek []byte, shared bool, seeds, ct, ss []byte) error {
	n := len(seeds) / 32
	ekSize := int(C.cb200_mlkem_public_key_size(C.int(k)))
	ctSize := int(C.cb200_mlkem_ciphertext_size(C.int(k)))
	stride := C.size_t(ekSize)
	if shared {
		stride = 0
		if len(ek) != ekSize {
			panic("cb200: ek must be one packed key")
		}
	} else if len(ek) != n*ekSize {
		panic("cb200: ek must hold n packed keys")
	}
	if len(ct) != n*ctSize || len(ss) != n*32 {
		panic("cb200: ct/ss have the wrong length")
	}
	if n == 0 {
		return nil
	}
	return call(func() C.int {
		return C.cb200_mlkem_encaps(C.int(k), (*C.uint8_t)(unsafe.Pointer(&ek[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&seeds[0])), (*C.uint8_t)(unsafe.Pointer(&ct[0])),
		(*C.uint8_t)(unsafe.Pointer(&ss[0])), nil, C.size_t(n))
	})
}

// MLKEMDecaps: batched UnmarshalBinaryPrivateKey + Decapsulate (implicit rejection included).
func MLKEMDecaps(k int, dk []byte, shared bool, ct, ss []byte) error {
	dkSize := int(C.cb200_mlkem_private_key_size(C.int(k)))
	ctSize := int(C.cb200_mlkem_ciphertext_size(C.int(k)))
	n := len(ct) / ctSize
	stride := C.size_t(dkSize)
	if shared {
		stride = 0
	}
	if len(ct) != n*ctSize || len(ss) != n*32 || (shared && len(dk) != dkSize) || (!shared && len(dk) != n*dkSize) {
		panic("cb200: dk/ct/ss have the wrong length")
	}
	if n == 0 {
		return nil
	}
	return call(func() C.int {
		return C.cb200_mlkem_decaps(C.int(k), (*C.uint8_t)(unsafe.Pointer(&dk[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&ct[0])), (*C.uint8_t)(unsafe.Pointer(&ss[0])), nil, C.size_t(n))
	})
}

// MLDSA65Sign signs len(msgs) messages.  sk holds one packed 4032-byte key (shared) or one per message.
// rnd == nil selects deterministic signing (sign/mldsa/mldsa65/dilithium.go:57-63).  ctx must be <= 255 bytes
// (the caller returns sign.ErrContextTooLong otherwise).  sig receives len(msgs)*3309 bytes.
func MLDSA65Sign(sk []byte, shared bool, msgs [][]byte, ctx, rnd, sig []byte) error {
	n := len(msgs)
	if n == 0 {
		return nil
	}
	off := make([]uint64, n+1)
	total := 0
	for i, m := range msgs {
		off[i] = uint64(total)
		total += len(m)
	}
	off[n] = uint64(total)
	blob := make([]byte, total+8)
	for i, m := range msgs {
		copy(blob[off[i]:], m)
	}
	stride := C.size_t(4032)
	if shared {
		stride = 0
	}
	var pctx, prnd *C.uint8_t
	if len(ctx) > 0 {
		pctx = (*C.uint8_t)(unsafe.Pointer(&ctx[0]))
	}
	if rnd != nil {
		prnd = (*C.uint8_t)(unsafe.Pointer(&rnd[0]))
	}
	return call(func() C.int {
		return C.cb200_mldsa65_sign((*C.uint8_t)(unsafe.Pointer(&sk[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&blob[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])), pctx, C.size_t(len(ctx)),
		prnd, (*C.uint8_t)(unsafe.Pointer(&sig[0])), nil, C.size_t(n), 0, nil)
	})
}

// Hybrid scheme identifiers of include/circl_b200.h (kem/hybrid/hybrid.go:34-62).
const (
	HybridX25519MLKEM768 = 0
	HybridKyber768X25519 = 1
	HybridKyber512X25519 = 2
)

// X25519Batch: x25519.KeyGen (points == nil) or x25519.Shared (dh/x25519/key.go:44-56) on n 32-byte keys.
// ok[i] is false where Shared would have returned false (point of small order); out[i] is then all zero.
func X25519Batch(scalars, points, out []byte, ok []bool) error {
	n := len(scalars) / 32
	if len(out) != n*32 || (points != nil && len(points) != n*32) || (ok != nil && len(ok) != n) {
		panic("cb200: X25519Batch buffers have the wrong length")
	}
	if n == 0 {
		return nil
	}
	status := make([]byte, n)
	var pp *C.uint8_t
	if points != nil {
		pp = (*C.uint8_t)(unsafe.Pointer(&points[0]))
	}
	var rc C.int
	err := call(func() C.int {
		rc = C.cb200_x25519((*C.uint8_t)(unsafe.Pointer(&scalars[0])), pp, (*C.uint8_t)(unsafe.Pointer(&out[0])),
			(*C.uint8_t)(unsafe.Pointer(&status[0])), C.size_t(n))
		if rc == C.CB200_ERR_PUBKEY { // reported per operation through ok, like the bool of x25519.Shared
			return 0
		}
		return rc
	})
	for i := range ok {
		ok[i] = status[i] == 0
	}
	return err
}

// XWingEncaps: batched xwing.Encapsulate (kem/xwing/xwing.go:173-182).  pk: one packed 1216-byte key (shared) or n keys;
// seeds: n*64 bytes; ct: n*1120; ss: n*32.  kem.ErrPubKey is returned if any ML-KEM half is not canonical.
func XWingEncaps(pk []byte, shared bool, seeds, ct, ss []byte) error {
	n := len(seeds) / 64
	stride := C.size_t(1216)
	if shared {
		stride = 0
	}
	if len(ct) != n*1120 || len(ss) != n*32 || (shared && len(pk) != 1216) || (!shared && len(pk) != n*1216) {
		panic("cb200: XWingEncaps buffers have the wrong length")
	}
	if n == 0 {
		return nil
	}
	return call(func() C.int {
		return C.cb200_xwing_encaps((*C.uint8_t)(unsafe.Pointer(&pk[0])), stride, (*C.uint8_t)(unsafe.Pointer(&seeds[0])),
		(*C.uint8_t)(unsafe.Pointer(&ct[0])), (*C.uint8_t)(unsafe.Pointer(&ss[0])), nil, C.size_t(n))
	})
}

// HybridEncaps: batched hybrid.scheme.EncapsulateDeterministically (kem/hybrid/hybrid.go:233-261); id is one of the
// Hybrid* constants; seeds: n*32 bytes; ss: n*64 bytes (the two shared secrets side by side, hybrid.go:260).
func HybridEncaps(id int, pk []byte, shared bool, seeds, ct, ss []byte) error {
	n := len(seeds) / 32
	pkSize := int(C.cb200_hybrid_public_key_size(C.int(id)))
	ctSize := int(C.cb200_hybrid_ciphertext_size(C.int(id)))
	stride := C.size_t(pkSize)
	if shared {
		stride = 0
	}
	if len(ct) != n*ctSize || len(ss) != n*64 || (shared && len(pk) != pkSize) || (!shared && len(pk) != n*pkSize) {
		panic("cb200: HybridEncaps buffers have the wrong length")
	}
	if n == 0 {
		return nil
	}
	return call(func() C.int {
		return C.cb200_hybrid_encaps(C.int(id), (*C.uint8_t)(unsafe.Pointer(&pk[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&seeds[0])), (*C.uint8_t)(unsafe.Pointer(&ct[0])), (*C.uint8_t)(unsafe.Pointer(&ss[0])),
		nil, C.size_t(n))
	})
}

// PinnedBytes returns a Go slice over cudaHostAlloc'd memory: large batches should live
// here so that host<->device copies run at full PCIe speed and overlap with the kernels.
func PinnedBytes(n int) ([]byte, func()) {
	runtime.LockOSThread()
	p := C.cb200_host_alloc(C.size_t(n))
	if p == nil {
		msg := C.GoString(C.cb200_last_error())
		runtime.UnlockOSThread()
		panic(msg)
	}
	runtime.UnlockOSThread()
	return unsafe.Slice((*byte)(p), n), func() { C.cb200_host_free(p) }
}

// KeccakF1600 permutes every 25-lane state in place: the batched counterpart of (*keccakf1600.StateX4).Permute
// (simd/keccakf1600/f1600x.go:115-121); turbo selects the 12-round variant.
func KeccakF1600(states [][25]uint64, turbo bool) error {
	if len(states) == 0 {
		return nil
	}
	t := C.int(0)
	if turbo {
		t = 1
	}
	return call(func() C.int {
		return C.cb200_keccak_f1600((*C.uint64_t)(unsafe.Pointer(&states[0])), C.size_t(len(states)), t)
	})
}

// KyberDeriveUniform: (*Poly).DeriveUniform (pke/kyber/internal/common/sample.go:192-236) for len(polys) (seed, x, y)
// triples; seeds holds one 32-byte seed (shared) or one per polynomial, xy two bytes per polynomial.
func KyberDeriveUniform(polys [][256]int16, seeds []byte, xy []byte) error {
	n := len(polys)
	if len(xy) != 2*n || (len(seeds) != 32 && len(seeds) != 32*n) {
		panic("cb200: KyberDeriveUniform buffers have the wrong length")
	}
	if n == 0 {
		return nil
	}
	stride := C.size_t(32)
	if len(seeds) == 32 {
		stride = 0
	}
	return call(func() C.int {
		return C.cb200_kyber_derive_uniform((*C.int16_t)(unsafe.Pointer(&polys[0])), (*C.uint8_t)(unsafe.Pointer(&seeds[0])),
			stride, (*C.uint8_t)(unsafe.Pointer(&xy[0])), C.size_t(n))
	})
}
