// katdump -- known-answer-test dump of the Lattigo-fork behaviour the MatMult hot path relies on (SURVEY.md 4(iv), 7 hard part 1,
// App. B "[ALL UNVERIFIED]"; VERDICT r1 item 8).  Run it on any box that has Go >= 1.18 and the modules of go.mod:5-16:
//
//	cp -r go/harness/katdump $SFGWAS/cmd/katdump && cd $SFGWAS && go run ./cmd/katdump -out /path/to/sfgwas-b200/tests/golden/lattigo
//
// For every parameter set it writes tests/golden/lattigo/<set>/ with meta.json + raw little-endian uint64 files (inputs AND outputs, so
// nothing has to be regenerated on the other side).  tests/test_lattigo_kat.py consumes the directory when present and compares the
// CPU oracle and the CUDA library bit for bit: psi per modulus, ring.NTT, EncoderBig.EncodeNTT, RotateNew under the dumped Galois keys,
// MulRelinNew + Rescale, Ciphertext.MarshalBinary, the DiagCacheStream files MatMult4StreamPreprocess leaves on disk, and the
// deterministic part S of MatMult4StreamCompute (SURVEY App. A.5) recomputed with the reference's own exported functions.
//
// NOT COMPILED in the authoring environment (no Go toolchain there): written against the API the reference itself uses
// (crypto/crypto.go:89-275, gwas/matmult.go:208-440,711-731,1043-1236, gwas/filestream.go:42-282).
package main

import (
	"encoding/binary"
	"encoding/json"
	"flag"
	"fmt"
	"math"
	"os"
	"path/filepath"

	"github.com/hhcho/sfgwas/crypto"
	"github.com/hhcho/sfgwas/gwas"
	"github.com/ldsec/lattigo/v2/ckks"
	"github.com/ldsec/lattigo/v2/ring"
)

// splitmix64: the same generator tests/test_lattigo_kat.py uses to cross-check the dumped inputs
type sm64 struct{ s uint64 }

func (r *sm64) next() uint64 {
	r.s += 0x9E3779B97F4A7C15
	z := r.s
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func check(err error) {
	if err != nil {
		panic(err)
	}
}

func writeU64(path string, rows ...[]uint64) {
	f, err := os.Create(path)
	check(err)
	defer f.Close()
	buf := make([]byte, 8)
	for _, r := range rows {
		for _, v := range r {
			binary.LittleEndian.PutUint64(buf, v)
			_, err = f.Write(buf)
			check(err)
		}
	}
}

func polyRows(p *ring.Poly, nLimbs int) [][]uint64 { return p.Coeffs[:nLimbs] }

// ciphertext -> [2][level+1][N]
func ctRows(ct *ckks.Ciphertext) [][]uint64 {
	out := make([][]uint64, 0)
	for k := 0; k < 2; k++ {
		out = append(out, ct.Value()[k].Coeffs[:ct.Level()+1]...)
	}
	return out
}

// switching key -> [beta][2][nQ+nP][N] exactly as cryptoParams.RotKs.Keys[galEl].Value[i][0|1].Coeffs holds it (gwas/matmult.go:1110)
func swkRows(value [][2]*ring.Poly) [][]uint64 {
	out := make([][]uint64, 0)
	for i := range value {
		for c := 0; c < 2; c++ {
			out = append(out, value[i][c].Coeffs...)
		}
	}
	return out
}

type meta struct {
	Set          string    `json:"set"`
	LogN         int       `json:"logN"`
	Qi           []uint64  `json:"Qi"`
	Pi           []uint64  `json:"Pi"`
	Scale        float64   `json:"scale"`
	Psi          []uint64  `json:"psi"`            // primitive 2N-th root per modulus (Q then P), plain (InvMForm of NttPsi[brv(1)])
	Beta         int       `json:"beta"`
	Rotations    []int     `json:"rotations"`      // left rotations whose Galois keys were dumped (rotkey_<k>.bin)
	GaloisEls    []uint64  `json:"galois_elements"`
	CtLevel      int       `json:"ct_level"`
	CtScale      float64   `json:"ct_scale"`
	MulLevel     int       `json:"mulrelin_level"` // level and scale of mulrelin_out.bin (after Rescale(params.Scale))
	MulScale     float64   `json:"mulrelin_scale"`
	GenoRows     int       `json:"geno_rows"`
	GenoCols     int       `json:"geno_cols"`
	MaxLevel     int       `json:"max_level"`
	S            int       `json:"s"`
	OutScale     float64   `json:"out_scale"`
	EncodeNrot   int       `json:"encode_nrot"`
	SkMontgomery bool      `json:"sk_is_ntt_montgomery"`
	Notes        string    `json:"notes"`
}

func dumpSet(name string, params *ckks.Parameters, outDir string) {
	dir := filepath.Join(outDir, name)
	check(os.MkdirAll(dir, 0o755))
	N := int(params.N())
	slots := N / 2
	logSlots := int(params.LogSlots())
	d := int(math.Ceil(math.Sqrt(float64(slots))))
	nQ, nP := len(params.Qi()), len(params.Pi())
	m := meta{Set: name, LogN: int(params.LogN()), Qi: params.Qi(), Pi: params.Pi(), Scale: params.Scale(), Beta: int(params.Beta()),
		MaxLevel: 5, SkMontgomery: true}
	rng := &sm64{s: 2024}

	// ---- ring constants: psi (SURVEY App. B.3: first primitive root found by incrementing from g = 2) ----
	ringQP, err := ring.NewRing(N, append(append([]uint64{}, params.Qi()...), params.Pi()...))
	check(err)
	for i, q := range ringQP.Modulus {
		m.Psi = append(m.Psi, ring.InvMForm(ringQP.NttPsi[i][N>>1], q, ringQP.MredParams[i])) // NttPsi[brv(1)] = MForm(psi)
	}

	// ---- ring.NTT / ring.InvNTT of a seeded polynomial over every modulus ----
	p := ringQP.NewPoly()
	for i, q := range ringQP.Modulus {
		for j := 0; j < N; j++ {
			p.Coeffs[i][j] = rng.next() % q
		}
	}
	writeU64(filepath.Join(dir, "ntt_in.bin"), p.Coeffs...)
	ringQP.NTT(p, p)
	writeU64(filepath.Join(dir, "ntt_out.bin"), p.Coeffs...)

	// ---- keys: one party, exactly the helpers the reference uses (crypto/crypto.go:159-179, 182-275) ----
	cps := crypto.NewCryptoParamsForNetwork(params, 1, 256)[0]
	if nQ > 5 { // the MatMult path needs level 5 (SURVEY App. B.1: impossible at logN = 12)
		cps.SetRotKeys(crypto.GenerateRotKeys(slots, 20, true))
	} else {
		cps.SetRotKeys([]crypto.RotationType{{Value: 1, Side: crypto.SideLeft}, {Value: d, Side: crypto.SideLeft}})
	}
	writeU64(filepath.Join(dir, "sk.bin"), cps.Sk.Value.Coeffs...)
	writeU64(filepath.Join(dir, "rlk.bin"), swkRows(cps.Rlk.Keys[0].Value)...)
	for _, k := range []int{1, d} {
		galEl := params.GaloisElementForColumnRotationBy(k)
		m.Rotations = append(m.Rotations, k)
		m.GaloisEls = append(m.GaloisEls, galEl)
		writeU64(filepath.Join(dir, fmt.Sprintf("rotkey_%d.bin", k)), swkRows(cps.RotKs.Keys[galEl].Value)...)
	}

	// ---- EncoderBig.EncodeNTT of a 0/1/2 slot vector (the diagonals of the path), right-rotated like convertToComplex128WithRot ----
	enc := ckks.NewEncoderBig(params, 256)
	vals := make([]complex128, slots)
	raw := make([]uint64, slots)
	m.EncodeNrot = 3 * d
	for j := 0; j < slots; j++ {
		raw[j] = rng.next() % 3
	}
	for j := 0; j < slots; j++ {
		vals[(j+m.EncodeNrot)%slots] = complex(float64(raw[j]), 0)
	}
	writeU64(filepath.Join(dir, "encode_values.bin"), raw)
	pt := ckks.NewPlaintext(params, params.MaxLevel(), params.Scale())
	enc.EncodeNTT(pt, vals, logSlots)
	writeU64(filepath.Join(dir, "encode_out.bin"), pt.Value()[0].Coeffs...)

	// ---- rotations, MulRelin + Rescale, MarshalBinary on a fresh ciphertext ----
	x := make([]float64, slots)
	for j := range x {
		x[j] = float64(int64(rng.next()%2001)-1000) / 500.0
	}
	cv, _ := crypto.EncryptFloatVector(cps, x)
	ct := cv[0]
	m.CtLevel, m.CtScale = ct.Level(), ct.Scale()
	xb := make([]uint64, slots)
	for j := range x {
		xb[j] = math.Float64bits(x[j])
	}
	writeU64(filepath.Join(dir, "ct_values_f64bits.bin"), xb)
	writeU64(filepath.Join(dir, "ct_in.bin"), ctRows(ct)...)
	eva := ckks.NewEvaluator(params, ckks.EvaluationKey{Rlk: cps.Rlk, Rtks: cps.RotKs})
	for _, k := range []int{1, d} {
		r := crypto.RotateRightWithEvaluator(cps, ct, -k, eva) // = RotateNew(ct, k): left rotation (crypto/basics.go:201-210)
		writeU64(filepath.Join(dir, fmt.Sprintf("rot_out_%d.bin", k)), ctRows(r)...)
	}
	prod := eva.MulRelinNew(ct, ct)
	check(eva.Rescale(prod, params.Scale(), prod))
	m.MulLevel, m.MulScale = prod.Level(), prod.Scale()
	writeU64(filepath.Join(dir, "mulrelin_out.bin"), ctRows(prod)...)
	mb, err := ct.MarshalBinary()
	check(err)
	check(os.WriteFile(filepath.Join(dir, "marshal.bin"), mb, 0o644))

	// ---- MatMult4StreamPreprocess cache files + the deterministic part S of MatMult4StreamCompute ----
	if nQ > 5 {
		nr, nc, s := slots+37, slots+29, 2 // 2 x 2 ragged blocks
		m.GenoRows, m.GenoCols, m.S = nr, nc, s
		geno := make([]byte, nr*nc)
		for i := range geno {
			geno[i] = byte(rng.next() % 3)
		}
		genoPath := filepath.Join(dir, "geno.bin")
		check(os.WriteFile(genoPath, geno, 0o644))
		prefix := filepath.Join(dir, "diagcache")
		gfs := gwas.NewGenoFileStream(genoPath, uint64(nr), uint64(nc), true)
		gwas.MatMult4StreamPreprocess(cps, gfs, 5, prefix)

		nbr := (nr-1)/slots + 1
		mct := (nc-1)/slots + 1
		A := make(crypto.CipherMatrix, s)
		for i := range A {
			row := make([]float64, nbr*slots)
			for j := 0; j < nr; j++ {
				row[j] = float64(int64(rng.next()%2001)-1000) / 500.0
			}
			A[i], _ = crypto.EncryptFloatVector(cps, row)
			for b := 0; b < nbr; b++ {
				writeU64(filepath.Join(dir, fmt.Sprintf("A_%d_%d.bin", i, b)), ctRows(A[i][b])...)
			}
		}
		m.OutScale = A[0][0].Scale() * params.Scale()
		// S[i][bj] = sum_g RotL_{g d}(reduce(acc[i][g])[bj]) with the reference's own building blocks, single-threaded
		// (gwas/matmult.go:1083-1236 without the randomised CZeroMat, SURVEY App. A.5)
		S := make([][]*ckks.Ciphertext, s)
		for i := range S {
			S[i] = make([]*ckks.Ciphertext, mct)
		}
		acc := make([][]gwas.CipherVectorAccV2, s)
		used := make([][]bool, s)
		for i := range acc {
			acc[i] = make([]gwas.CipherVectorAccV2, d)
			used[i] = make([]bool, d)
		}
		for bi := 0; bi < nbr; bi++ {
			dcs, _ := gwas.NewDiagCacheStream(cps, prefix, bi, false)
			babyTable, giantTable := dcs.GetIndexTables()
			rot := make([][]*ckks.Ciphertext, s)
			for i := range rot {
				rot[i] = make([]*ckks.Ciphertext, d)
				for b := 0; b < d; b++ {
					if babyTable[b] {
						rot[i][b] = crypto.RotateRightWithEvaluator(cps, A[i][bi], -b, eva)
					}
				}
				for g := 0; g < d; g++ {
					if giantTable[g] && !used[i][g] {
						acc[i][g] = gwas.NewCipherVectorAccV2(cps, mct, 5)
						used[i][g] = true
					}
				}
			}
			for pv, shift := dcs.ReadDiag(); pv != nil; pv, shift = dcs.ReadDiag() {
				baby, giant := shift%d, shift/d
				for i := range A {
					gwas.CPMultAccWithoutMRedV2(crypto.CipherVector{rot[i][baby]}, pv, acc[i][giant])
				}
			}
			dcs.Close()
		}
		for i := 0; i < s; i++ {
			for g := 0; g < d; g++ {
				if !used[i][g] {
					continue
				}
				cvr := gwas.ModularReduceV2(cps, acc[i][g], m.OutScale)
				for bj := range cvr {
					if g > 0 {
						cvr[bj] = crypto.RotateRightWithEvaluator(cps, cvr[bj], -g*d, eva)
					}
					if S[i][bj] == nil {
						S[i][bj] = cvr[bj]
					} else {
						eva.Add(S[i][bj], cvr[bj], S[i][bj])
					}
				}
			}
			for bj := 0; bj < mct; bj++ {
				writeU64(filepath.Join(dir, fmt.Sprintf("S_%d_%d.bin", i, bj)), ctRows(S[i][bj])...)
			}
		}
	}
	m.Notes = "raw little-endian uint64 files; ciphertexts [2][level+1][N], keys [beta][2][nQ+nP][N] as stored by Lattigo (NTT + Montgomery)"
	js, err := json.MarshalIndent(m, "", " ")
	check(err)
	check(os.WriteFile(filepath.Join(dir, "meta.json"), js, 0o644))
	fmt.Println("wrote", dir)
}

func main() {
	out := flag.String("out", "tests/golden/lattigo", "output directory")
	flag.Parse()
	sets := map[string]int{"PN12QP109": ckks.PN12QP109, "PN13QP218": ckks.PN13QP218, "PN14QP438": ckks.PN14QP438}
	for _, name := range []string{"PN12QP109", "PN13QP218", "PN14QP438"} {
		dumpSet(name, ckks.DefaultParams[sets[name]], *out)
	}
}
