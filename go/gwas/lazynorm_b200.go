//go:build b200

// lazynorm_b200.go -- cgo shim for the two lazy-normalisation wrappers of package gwas (gwas/matmult.go:27-77, 83-116 of
// hhcho/sfgwas): QXLazyNormStream and QXtLazyNormStream.  Their ciphertext algebra -- crypto.CMult, InnerProd, InnerSumAll,
// CMultScalar, MaskTrunc, eval.Sub (crypto/basics.go:110-127,236-293,386-427,553-566) -- runs in libsfgwas_b200.so
// (sfg_ct_mul_relin, sfg_inner_sum_all, sfg_ct_mul_plain, sfg_ct_sub); the collective bootstrap stays the reference's network
// protocol.  Go signatures are verbatim.  Not compiled in this repo's image (no Go toolchain); the same C entry points are
// exercised through sfgwas_b200/gwas.py (tests/test_gpu_ctalg.py).
package gwas

/*
#include <stdlib.h>
#include "sfgwas_b200.h"
*/
import "C"

import (
	"fmt"
	"math"
	"unsafe"

	"github.com/hhcho/sfgwas/crypto"
	"github.com/hhcho/sfgwas/mpc"
	"github.com/ldsec/lattigo/v2/ckks"
)

// flat copies a CipherVector into one C-layout buffer [n][2][nl][N], truncated to nl limbs (= DropLevel)
func b200Flat(X crypto.CipherVector, nl int) []uint64 {
	// every element must carry at least nl limbs and the vector one scale: the flat device calls take ONE level / scale per operand
	for i, ct := range X {
		if ct.Level()+1 < nl || math.Abs(ct.Scale()/X[0].Scale()-1) > 1e-9 {
			panic(fmt.Sprintf("b200Flat: element %d has level %d / scale %g, the vector is used at %d limbs / scale %g", i, ct.Level(), ct.Scale(), nl, X[0].Scale()))
		}
	}
	N := len(X[0].Value()[0].Coeffs[0])
	buf := make([]uint64, 0, len(X)*2*nl*N)
	for _, ct := range X {
		for k := 0; k < 2; k++ {
			for l := 0; l < nl; l++ {
				buf = append(buf, ct.Value()[k].Coeffs[l]...)
			}
		}
	}
	return buf
}

func b200Unflat(cps *crypto.CryptoParams, buf []uint64, n, level int, scale float64) crypto.CipherVector {
	N := int(cps.Params.N())
	out := make(crypto.CipherVector, n)
	off := 0
	for i := range out {
		out[i] = ckks.NewCiphertext(cps.Params, 1, level, scale)
		for k := 0; k < 2; k++ {
			for l := 0; l <= level; l++ {
				copy(out[i].Value()[k].Coeffs[l], buf[off:off+N])
				off += N
			}
		}
	}
	return out
}

// evaluator.Rescale(ct, params.Scale(), ct): number of DivRoundByLastModulusNTT steps and the resulting scale (Lattigo v2.1)
func b200RescaleSteps(cps *crypto.CryptoParams, scale float64, level int) (int, float64) {
	qi := cps.Params.Qi()
	n := 0
	for level-n > 0 && scale >= cps.Params.Scale()*float64(qi[level-n])/2 {
		scale /= float64(qi[level-n])
		n++
	}
	return n, scale
}

func minInt(a, b int) int {
	if a < b {
		return a
	}
	return b
}

func b200UploadRlk(c *b200Ctx, cps *crypto.CryptoParams) {
	// cryptoParams.Rlk.Keys[0].Value[i][0|1].Coeffs[*] (crypto/crypto.go:45-60), flat copy
	var buf []uint64
	for _, v := range cps.Rlk.Keys[0].Value {
		for k := 0; k < 2; k++ {
			for _, limb := range v[k].Coeffs {
				buf = append(buf, limb...)
			}
		}
	}
	b200Check(c, C.sfg_ctx_set_relin_key(c.h, (*C.uint64_t)(unsafe.Pointer(&buf[0]))), "sfg_ctx_set_relin_key")
	// InnerSumAll needs the power-of-two rotation keys as well (crypto/crypto.go:232-249): uploaded by b200Context's `upload`
	// for k = 1, 2, 4, .. slots/2 when the wrappers are used.
}

// crypto.CMult / CMultScalar (crypto/basics.go:386-427,553-566)
func b200CMult(cps *crypto.CryptoParams, X, Y crypto.CipherVector) crypto.CipherVector {
	c := b200Context(cps)
	level := minInt(X[0].Level(), Y[0].Level())
	nres, scale := b200RescaleSteps(cps, X[0].Scale()*Y[0].Scale(), level)
	n := len(X)
	if len(Y) > n {
		n = len(Y)
	}
	fx, fy := b200Flat(X, level+1), b200Flat(Y, level+1)
	N := int(cps.Params.N())
	out := make([]uint64, n*2*(level+1-nres)*N)
	c.mu.Lock()
	rc := C.sfg_ct_mul_relin(c.h, C.int(level), (*C.uint64_t)(unsafe.Pointer(&fx[0])), C.int(len(X)), C.int(level+1),
		(*C.uint64_t)(unsafe.Pointer(&fy[0])), C.int(len(Y)), C.int(level+1), C.int(nres), (*C.uint64_t)(unsafe.Pointer(&out[0])))
	c.mu.Unlock()
	b200Check(c, rc, "CMult")
	return b200Unflat(cps, out, n, level-nres, scale)
}

// crypto.InnerSumAll (crypto/basics.go:278-293)
func b200InnerSumAll(cps *crypto.CryptoParams, X crypto.CipherVector) *ckks.Ciphertext {
	c := b200Context(cps)
	level := X[0].Level()
	fx := b200Flat(X, level+1)
	out := make([]uint64, 2*(level+1)*int(cps.Params.N()))
	c.mu.Lock()
	rc := C.sfg_inner_sum_all(c.h, C.int(level), (*C.uint64_t)(unsafe.Pointer(&fx[0])), 1, C.int(len(X)), (*C.uint64_t)(unsafe.Pointer(&out[0])))
	c.mu.Unlock()
	b200Check(c, rc, "InnerSumAll")
	return b200Unflat(cps, out, 1, level, X[0].Scale())[0]
}

// eval.Sub(a, b, a) for operands of matching scale (gwas/matmult.go:54,100)
func b200Sub(cps *crypto.CryptoParams, a, b *ckks.Ciphertext) *ckks.Ciphertext {
	// Lattigo's evaluator.Sub first multiplies the lower-scale operand by floor(ratio) when the scales differ by 2x or more; the device
	// path subtracts as-is, which is only the reference's result for matching scales -- refuse anything else instead of silently diverging
	if r := math.Max(a.Scale(), b.Scale()) / math.Min(a.Scale(), b.Scale()); math.Floor(r) > 1 {
		panic(fmt.Sprintf("b200Sub: scales %g and %g need Lattigo's scale matching (ratio %.3f)", a.Scale(), b.Scale(), r))
	}
	c := b200Context(cps)
	level := minInt(a.Level(), b.Level())
	fa, fb := b200Flat(crypto.CipherVector{a}, level+1), b200Flat(crypto.CipherVector{b}, level+1)
	out := make([]uint64, 2*(level+1)*int(cps.Params.N()))
	c.mu.Lock()
	rc := C.sfg_ct_sub(c.h, C.int(level), (*C.uint64_t)(unsafe.Pointer(&fa[0])), 1, C.int(level+1), (*C.uint64_t)(unsafe.Pointer(&fb[0])), 1,
		C.int(level+1), (*C.uint64_t)(unsafe.Pointer(&out[0])))
	c.mu.Unlock()
	b200Check(c, rc, "Sub")
	scale := a.Scale()
	if b.Scale() > scale {
		scale = b.Scale()
	}
	return b200Unflat(cps, out, 1, level, scale)[0]
}

// crypto.MaskTrunc (crypto/basics.go:110-127): the mask plaintext is encoded by the reference's own encoder (bit parity with
// its float64 FFT) and multiplied on the device
func b200MaskTrunc(cps *crypto.CryptoParams, ct *ckks.Ciphertext, N int) *ckks.Ciphertext {
	if N == cps.GetSlots() {
		return ct
	}
	m := make([]float64, cps.GetSlots())
	for i := 0; i < N; i++ {
		m[i] = 1.0
	}
	mask, _ := crypto.EncodeFloatVector(cps, m)
	c := b200Context(cps)
	level := minInt(ct.Level(), mask[0].Level())
	nres, scale := b200RescaleSteps(cps, ct.Scale()*mask[0].Scale(), level)
	var pt []uint64
	for l := 0; l <= level; l++ {
		pt = append(pt, mask[0].Value()[0].Coeffs[l]...)
	}
	fc := b200Flat(crypto.CipherVector{ct}, level+1)
	out := make([]uint64, 2*(level+1-nres)*int(cps.Params.N()))
	c.mu.Lock()
	rc := C.sfg_ct_mul_plain(c.h, C.int(level), (*C.uint64_t)(unsafe.Pointer(&pt[0])), 1, C.int(level+1), (*C.uint64_t)(unsafe.Pointer(&fc[0])), 1,
		C.int(level+1), C.int(nres), (*C.uint64_t)(unsafe.Pointer(&out[0])))
	c.mu.Unlock()
	b200Check(c, rc, "MaskTrunc")
	return b200Unflat(cps, out, 1, level-nres, scale)[0]
}

// QXLazyNormStream: gwas/matmult.go:27-77, (Q*S)*X - ((Q*S)*m)*1^T
func QXLazyNormStream(cps *crypto.CryptoParams, mpcObj *mpc.MPC, Q crypto.CipherMatrix, Xcachefile string, XMean, XStdInv crypto.CipherVector, numInd int) (out crypto.CipherMatrix) {
	if mpcObj.GetPid() == 0 {
		return
	}
	slots := cps.GetSlots()
	QS := make(crypto.CipherMatrix, len(Q))
	for i := range Q {
		QS[i] = b200CMult(cps, Q[i], XStdInv)
	}
	out = MatMult4StreamCompute(cps, QS, 5, Xcachefile)
	out = mpcObj.Network.BootstrapMatAll(cps, out)
	for i := range QS {
		QSm := b200InnerSumAll(cps, b200CMult(cps, QS[i], XMean))
		for j := range out[i] {
			out[i][j] = b200Sub(cps, out[i][j], QSm)
			n := slots
			if j == len(out[i])-1 {
				n = ((numInd - 1) % slots) + 1
			}
			out[i][j] = b200MaskTrunc(cps, out[i][j], n)
		}
	}
	return
}

// QXtLazyNormStream: gwas/matmult.go:83-116, ((Q*X^T) - ((Q*1)*m^T))*S
func QXtLazyNormStream(cps *crypto.CryptoParams, mpcObj *mpc.MPC, Q crypto.CipherMatrix, XTcachefile string, XMean, XStdInv crypto.CipherVector) (out crypto.CipherMatrix) {
	if mpcObj.GetPid() == 0 {
		return
	}
	out = MatMult4StreamCompute(cps, Q, 5, XTcachefile)
	out = mpcObj.Network.BootstrapMatAll(cps, out)
	for i := range out {
		rowSum := b200InnerSumAll(cps, Q[i])
		Q1m := b200CMult(cps, XMean, crypto.CipherVector{rowSum})
		for j := range out[i] {
			out[i][j] = b200Sub(cps, out[i][j], Q1m[j])
		}
	}
	for i := range out {
		out[i] = b200CMult(cps, out[i], XStdInv)
	}
	return
}
