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

// Init binds the process to one GPU.  One process per GPU; safe to call many times.
func Init(device int) error {
	initOnce.Do(func() {
		if rc := C.cb200_init(C.int(device)); rc != 0 {
			initErr = errors.New(C.GoString(C.cb200_last_error()))
		}
	})
	return initErr
}

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
	return lastErr(C.cb200_kyber_ntt((*C.int16_t)(unsafe.Pointer(&polys[0])), C.size_t(len(polys)), inv))
}

// KyberMulHat: p[i] = a[i] (*) b[i]  (mulHatAVX2, stubs_amd64.go:17).
func KyberMulHat(p, a, b [][256]int16) error {
	if len(p) != len(a) || len(p) != len(b) {
		panic("cb200: length mismatch")
	}
	if len(p) == 0 {
		return nil
	}
	return lastErr(C.cb200_kyber_mulhat((*C.int16_t)(unsafe.Pointer(&p[0])), (*C.int16_t)(unsafe.Pointer(&a[0])),
		(*C.int16_t)(unsafe.Pointer(&b[0])), C.size_t(len(p))))
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
	return lastErr(C.cb200_dil_ntt((*C.uint32_t)(unsafe.Pointer(&polys[0])), C.size_t(len(polys)), inv))
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
	return lastErr(C.cb200_mlkem_encaps(C.int(k), (*C.uint8_t)(unsafe.Pointer(&ek[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&seeds[0])), (*C.uint8_t)(unsafe.Pointer(&ct[0])),
		(*C.uint8_t)(unsafe.Pointer(&ss[0])), nil, C.size_t(n)))
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
	return lastErr(C.cb200_mlkem_decaps(C.int(k), (*C.uint8_t)(unsafe.Pointer(&dk[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&ct[0])), (*C.uint8_t)(unsafe.Pointer(&ss[0])), nil, C.size_t(n)))
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
	return lastErr(C.cb200_mldsa65_sign((*C.uint8_t)(unsafe.Pointer(&sk[0])), stride,
		(*C.uint8_t)(unsafe.Pointer(&blob[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])), pctx, C.size_t(len(ctx)),
		prnd, (*C.uint8_t)(unsafe.Pointer(&sig[0])), nil, C.size_t(n), 0, nil))
}

// PinnedBytes returns a Go slice over cudaHostAlloc'd memory: large batches should live
// here so that host<->device copies run at full PCIe speed and overlap with the kernels.
func PinnedBytes(n int) ([]byte, func()) {
	p := C.cb200_host_alloc(C.size_t(n))
	if p == nil {
		panic(C.GoString(C.cb200_last_error()))
	}
	return unsafe.Slice((*byte)(p), n), func() { C.cb200_host_free(p) }
}
