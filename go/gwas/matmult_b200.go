//go:build b200

// matmult_b200.go -- cgo shim that replaces the bodies of the three stream MatMult entry points of package gwas
// (gwas/matmult.go:914, 1043, 1238 of hhcho/sfgwas) with calls into libsfgwas_b200.so (include/sfgwas_b200.h).
//
// Build: copy this file into the reference's gwas/ directory, build with `-tags b200`, and exclude the original bodies with
// `//go:build !b200` on a file holding MatMult4StreamPreprocess / MatMult4StreamCompute / MatMult4Stream.  The Go signatures
// are verbatim, so sfgwas.go, mpc/, pca.go and assoc.go are unchanged.  (No Go toolchain exists in the build image of this
// repo: this file is the binding a maintainer adds, it is not compiled here.  The same calls are exercised through the
// Python mirror sfgwas_b200/gwas.py.)
package gwas

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../sfgwas_b200/lib -lsfgwas_b200 -Wl,-rpath,${SRCDIR}/../../sfgwas_b200/lib
#include <stdlib.h>
#include "sfgwas_b200.h"
*/
import "C"

import (
	"fmt"
	"math"
	"os"
	"runtime"
	"strconv"
	"sync"
	"unsafe"

	"github.com/hhcho/sfgwas/crypto"
	"github.com/ldsec/lattigo/v2/ckks"
)

// one GPU context per CryptoParams (ring tables + Galois keys + stream); created lazily, destroyed by a finalizer
type b200Ctx struct {
	h  *C.sfg_ctx
	mu sync.Mutex // calls on ONE context are serialised (include/sfgwas_b200.h conventions); use several contexts for concurrency
}

var (
	b200Mu   sync.Mutex
	b200Ctxs = map[*crypto.CryptoParams]*b200Ctx{}
	// device-resident diagonal caches keyed by the reference's cacheFilePrefix (replaces the prefix_<bi>.bin files)
	b200Caches = map[string]*C.sfg_cache{}
)

func b200Check(c *b200Ctx, rc C.int, what string) {
	if rc != 0 {
		// the reference panics / log.Fatal's on failure (gwas/matmult.go:360-362, gwas/filestream.go:59-61)
		panic(fmt.Sprintf("%s: %s", what, C.GoString(C.sfg_last_error(c.h))))
	}
}

// cPtrs copies a Go slice of per-limb pointers into C memory (cgo: no Go pointer to Go pointer). The limb backing arrays
// are pinned for the duration of the call with runtime.Pinner (Go >= 1.21; the reference's go.mod says 1.18: with an older toolchain
// replace the Pinner by C.malloc'ed staging copies of the limbs -- the library never retains the pointers after a call returns).
func cPtrs(pin *runtime.Pinner, limbs [][]uint64) (**C.uint64_t, func()) {
	n := len(limbs)
	arr := (**C.uint64_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	view := unsafe.Slice(arr, n)
	for i, l := range limbs {
		pin.Pin(&l[0])
		view[i] = (*C.uint64_t)(unsafe.Pointer(&l[0]))
	}
	return arr, func() { C.free(unsafe.Pointer(arr)) }
}

func b200Context(cps *crypto.CryptoParams) *b200Ctx {
	b200Mu.Lock()
	defer b200Mu.Unlock()
	if c, ok := b200Ctxs[cps]; ok {
		return c
	}
	p := cps.Params
	qi, pi := p.Qi(), p.Pi() // never hard-code the chain (SURVEY App. B.1)
	c := &b200Ctx{}
	// device: SFG_B200_DEVICE (one process per GPU, like the torch.distributed deployment of sfgwas_b200/dist.py), default 0
	dev := 0
	if v := os.Getenv("SFG_B200_DEVICE"); v != "" {
		if d, err := strconv.Atoi(v); err == nil {
			dev = d
		}
	}
	rc := C.sfg_ctx_create(C.int(dev), C.int(p.LogN()), (*C.uint64_t)(unsafe.Pointer(&qi[0])), C.int(len(qi)),
		(*C.uint64_t)(unsafe.Pointer(&pi[0])), C.int(len(pi)), C.double(p.Scale()), nil, &c.h)
	if rc != 0 {
		panic("sfg_ctx_create: " + C.GoString(C.sfg_last_error(nil)))
	}
	// the 2(d-1) BSGS Galois keys: cryptoParams.RotKs.Keys[galEl].Value[i][0|1].Coeffs[*] (crypto/crypto.go:251-264)
	slots := cps.GetSlots()
	d := int(math.Ceil(math.Sqrt(float64(slots))))
	upload := func(k int) {
		galEl := p.GaloisElementForColumnRotationBy(k)
		swk, ok := cps.RotKs.Keys[galEl]
		if !ok {
			return // reported by the library when the rotation is actually needed, like the reference's evaluator
		}
		var limbs [][]uint64
		for i := range swk.Value {
			for c01 := 0; c01 < 2; c01++ {
				limbs = append(limbs, swk.Value[i][c01].Coeffs...)
			}
		}
		var pin runtime.Pinner
		defer pin.Unpin()
		arr, free := cPtrs(&pin, limbs)
		defer free()
		b200Check(c, C.sfg_ctx_set_rotation_key_ptrs(c.h, C.int(k), arr), "sfg_ctx_set_rotation_key_ptrs")
	}
	for k := 1; k < d; k++ {
		upload(k)
		if k*d < slots {
			upload(k * d)
		}
	}
	for k := 1; k < slots; k *= 2 { // power-of-two keys of crypto.InnerSumAll (crypto/crypto.go:232-249), used by lazynorm_b200.go
		if C.sfg_ctx_has_rotation_key(c.h, C.int(k)) == 0 {
			upload(k)
		}
	}
	if cps.Rlk != nil {
		b200UploadRlk(c, cps)
	}
	// released explicitly with B200Release (b200Ctxs keeps the context reachable, so a finalizer would never run)
	b200Ctxs[cps] = c
	return c
}

func b200PushGeno(c *b200Ctx, gfs *GenoFileStream) *C.sfg_geno {
	nrows, ncols := gfs.NumRowsToKeep(), gfs.NumColsToKeep() // gwas/matmult.go:1241
	var g *C.sfg_geno
	b200Check(c, C.sfg_geno_create(c.h, C.size_t(nrows), C.size_t(ncols), &g), "sfg_geno_create")
	const chunk = 1024
	buf := make([]int8, 0, chunk*int(ncols))
	flush := func() {
		if len(buf) > 0 {
			b200Check(c, C.sfg_geno_push_rows(g, (*C.int8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)/int(ncols))), "sfg_geno_push_rows")
			buf = buf[:0]
		}
	}
	gfs.Reset()
	for r := uint64(0); r < nrows; r++ {
		buf = append(buf, gfs.NextRow()...) // filters and missing->0 already applied by the stream (gwas/filestream.go:327-360)
		if len(buf) == cap(buf) {
			flush()
		}
	}
	flush()
	return g
}

func ctLimbs(A crypto.CipherMatrix) [][]uint64 {
	var limbs [][]uint64
	for i := range A {
		for j := range A[i] {
			for k := 0; k < 2; k++ {
				limbs = append(limbs, A[i][j].Value()[k].Coeffs...)
			}
		}
	}
	return limbs
}

// MatMult4StreamPreprocess: gwas/matmult.go:914-1041. The diagonal cache is built in HBM and registered under cacheFilePrefix.
func MatMult4StreamPreprocess(cryptoParams *crypto.CryptoParams, gfs *GenoFileStream, maxLevel int, cacheFilePrefix string) {
	c := b200Context(cryptoParams)
	c.mu.Lock()
	defer c.mu.Unlock()
	g := b200PushGeno(c, gfs)
	defer C.sfg_geno_destroy(g) // the cache keeps its own reference
	var cache *C.sfg_cache
	b200Check(c, C.sfg_matmult4_stream_preprocess(c.h, g, C.int(maxLevel), &cache), "MatMult4StreamPreprocess")
	b200Mu.Lock()
	if old, ok := b200Caches[cacheFilePrefix]; ok {
		C.sfg_cache_destroy(old)
	}
	b200Caches[cacheFilePrefix] = cache
	b200Mu.Unlock()
}

// B200ReleaseCache frees the HBM-resident diagonal cache registered under cacheFilePrefix (the reference's cache lives in files and
// costs no memory between calls; here a 10k x 100k cache is 55 GB of HBM, so the protocol driver releases X / XT caches when PCA ends).
func B200ReleaseCache(cacheFilePrefix string) {
	b200Mu.Lock()
	defer b200Mu.Unlock()
	if cache, ok := b200Caches[cacheFilePrefix]; ok {
		C.sfg_cache_destroy(cache)
		delete(b200Caches, cacheFilePrefix)
	}
}

// B200Release destroys the GPU context of cps together with every cache (end of the protocol run; the finalizer alone never fires
// because b200Ctxs keeps the context reachable).
func B200Release(cps *crypto.CryptoParams) {
	b200Mu.Lock()
	defer b200Mu.Unlock()
	for k, cache := range b200Caches {
		C.sfg_cache_destroy(cache)
		delete(b200Caches, k)
	}
	if c, ok := b200Ctxs[cps]; ok {
		C.sfg_ctx_destroy(c.h)
		c.h = nil
		delete(b200Ctxs, cps)
	}
}

// MatMult4StreamCompute: gwas/matmult.go:1043-1236.
func MatMult4StreamCompute(cryptoParams *crypto.CryptoParams, A crypto.CipherMatrix, maxLevel int, cacheFilePrefix string) crypto.CipherMatrix {
	c := b200Context(cryptoParams)
	b200Mu.Lock()
	cache, ok := b200Caches[cacheFilePrefix]
	b200Mu.Unlock()
	if !ok {
		// no HBM cache under this prefix: build it from the reference's <prefix>_<bi>.bin files (written by an earlier run or by the
		// CPU path); a missing file fails like NewDiagCacheStream's panic (gwas/filestream.go:56-61)
		cp := C.CString(cacheFilePrefix)
		c.mu.Lock()
		rc := C.sfg_cache_load_files(c.h, cp, 0, 0, C.int(maxLevel), &cache)
		c.mu.Unlock()
		C.free(unsafe.Pointer(cp))
		b200Check(c, rc, "NewDiagCacheStream")
		b200Mu.Lock()
		b200Caches[cacheFilePrefix] = cache
		b200Mu.Unlock()
	}
	var mct, nbr C.int
	C.sfg_cache_info(cache, nil, nil, nil, &mct, &nbr)
	_ = nbr // the library checks len(A[0]) against the cache's block rows and reports the reference-style error
	s := len(A)
	outScale := A[0][0].Scale() * cryptoParams.Params.Scale() // gwas/matmult.go:1045
	S := make(crypto.CipherMatrix, s)
	for i := range S {
		S[i] = make(crypto.CipherVector, int(mct))
		for j := range S[i] {
			S[i][j] = ckks.NewCiphertext(cryptoParams.Params, 1, maxLevel-1, outScale) // level len(acc0)-1 (gwas/matmult.go:350)
		}
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	aPtrs, freeA := cPtrs(&pin, ctLimbs(A))
	defer freeA()
	oPtrs, freeO := cPtrs(&pin, ctLimbs(S))
	defer freeO()
	c.mu.Lock()
	// num_block_rows = len(A[0]): the pointer list is built from A, so a mismatch with the cache is caught by the library's check_args
	rc := C.sfg_matmult4_stream_compute_ptrs(c.h, aPtrs, C.int(s), C.int(len(A[0])), C.int(A[0][0].Level()), C.int(maxLevel), cache, oPtrs)
	c.mu.Unlock()
	b200Check(c, rc, "MatMult4StreamCompute")
	// reference semantics: out starts as a FRESH randomised encryption of zero and the sum is added with eva.Add
	// (gwas/matmult.go:1174,1223-1227; SURVEY App. A.5) -- stays on the Go side
	out := crypto.CZeroMat(cryptoParams, int(mct), s)
	cryptoParams.WithEvaluator(func(eva ckks.Evaluator) error {
		for i := range out {
			for j := range out[i] {
				eva.Add(out[i][j], S[i][j], out[i][j])
			}
		}
		return nil
	})
	return out
}

// MatMult4Stream: gwas/matmult.go:1238-1505 (fused; no cache kept).
func MatMult4Stream(cryptoParams *crypto.CryptoParams, A crypto.CipherMatrix, gfs *GenoFileStream, maxLevel int,
	computeSquaredSum, square bool, nproc int) (crypto.CipherMatrix, []float64, []float64) {
	_ = nproc // host parallelism knob of the reference (gwas/matmult.go:1242-1244); the GPU schedule does not use it
	c := b200Context(cryptoParams)
	c.mu.Lock()
	defer c.mu.Unlock()
	g := b200PushGeno(c, gfs)
	defer C.sfg_geno_destroy(g)
	s, ncols := len(A), int(gfs.NumColsToKeep())
	slots := cryptoParams.GetSlots()
	mct := (ncols-1)/slots + 1
	outScale := A[0][0].Scale() * cryptoParams.Params.Scale()
	S := make(crypto.CipherMatrix, s)
	for i := range S {
		S[i] = make(crypto.CipherVector, mct)
		for j := range S[i] {
			S[i][j] = ckks.NewCiphertext(cryptoParams.Params, 1, maxLevel-1, outScale)
		}
	}
	var sum, sqSum []float64 // stay nil unless requested, like the reference (gwas/matmult.go:1246-1251)
	var pSum, pSq *C.double
	if computeSquaredSum {
		sum, sqSum = make([]float64, ncols), make([]float64, ncols)
		pSum, pSq = (*C.double)(unsafe.Pointer(&sum[0])), (*C.double)(unsafe.Pointer(&sqSum[0]))
	}
	// flat variant: A / out are copied into contiguous C buffers (the _ptrs variant avoids this copy for Compute)
	nl := A[0][0].Level() + 1
	N := int(cryptoParams.Params.N())
	flatA := make([]uint64, 0, s*len(A[0])*2*nl*N)
	for _, l := range ctLimbs(A) {
		flatA = append(flatA, l...)
	}
	flatO := make([]uint64, s*mct*2*maxLevel*N)
	csq, sq := 0, 0
	if computeSquaredSum {
		csq = 1
	}
	if square {
		sq = 1
	}
	b200Check(c, C.sfg_matmult4_stream(c.h, (*C.uint64_t)(unsafe.Pointer(&flatA[0])), C.int(s), C.int(nl-1), g, C.int(maxLevel),
		C.int(csq), C.int(sq), (*C.uint64_t)(unsafe.Pointer(&flatO[0])), pSum, pSq), "MatMult4Stream")
	off := 0
	for _, l := range ctLimbs(S) {
		copy(l, flatO[off:off+N])
		off += N
	}
	out := crypto.CZeroMat(cryptoParams, mct, s)
	cryptoParams.WithEvaluator(func(eva ckks.Evaluator) error {
		for i := range out {
			for j := range out[i] {
				eva.Add(out[i][j], S[i][j], out[i][j])
			}
		}
		return nil
	})
	return out, sum, sqSum
}
