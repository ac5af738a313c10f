// cmfile.cpp -- crypto.SaveCipherMatrixToFile / LoadCipherMatrixFromFile (crypto/utilities.go:82-141): the
// on-disk format of `assoc_cache_mult.%d.bin` (gwas/assoc.go:317-333,434-437) and `Qcomb.bin`, so that a GPU run can hand its MatMult
// outputs to -- or pick them up from -- a CPU run of the reference.  Host code: the format is I/O, not arithmetic.
//
//   file  = u32 LE nrows | u32 LE ncols | u64 LE len(sizes) | sizes: nrows*ncols u64 LE (bytes of each marshalled ciphertext)
//           | u64 LE len(blob) | blob = concatenated Ciphertext.MarshalBinary()                      (crypto/utilities.go:35-56,82-113)
//   Ciphertext.MarshalBinary (Lattigo v2.1 ckks.Element; [UNVERIFIED vs the fork]):
//           u8 degree+1 | f64 LE scale | u8 isNTT | per polynomial: u8 log2(N) | u8 numModuli | numModuli*N u64 BIG-endian coefficients
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace sfg {

static void put_le(std::vector<unsigned char> &b, uint64_t v, int n) {
    for (int i = 0; i < n; i++) b.push_back((unsigned char)(v >> (8 * i)));
}
static uint64_t get_le(const unsigned char *p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

size_t cm_ct_bytes(int logN, int nl) { return 10 + 2 * (2 + ((size_t)nl << logN) * 8); }

// cts: [nrows][ncols][2][nl][N] canonical residues (NTT domain, what ct.Value()[k].Coeffs[l][j] holds); scales: [nrows][ncols]
int cm_save(const char *filename, int logN, const uint64_t *cts, const double *scales, int nrows, int ncols, int nl, std::string &err) {
    if (nrows < 1 || ncols < 1 || nl < 1 || nl > 255 || logN < 1 || logN > 17) {
        err = "SaveCipherMatrixToFile: bad dimensions";
        return -1;
    }
    const size_t N = (size_t)1 << logN, ctb = cm_ct_bytes(logN, nl), nct = (size_t)nrows * ncols;
    FILE *f = fopen(filename, "wb");
    if (!f) {
        err = std::string("create ") + filename + ": " + strerror(errno);  // the reference log.Fatal's (crypto/utilities.go:85-87)
        return -1;
    }
    std::vector<unsigned char> hdr;
    put_le(hdr, (uint64_t)nrows, 4);
    put_le(hdr, (uint64_t)ncols, 4);
    put_le(hdr, nct * 8, 8);
    for (size_t k = 0; k < nct; k++) put_le(hdr, ctb, 8);
    put_le(hdr, nct * ctb, 8);
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    std::vector<unsigned char> buf(ctb);
    for (size_t k = 0; k < nct && ok; k++) {
        unsigned char *p = buf.data();
        *p++ = 2;  // degree + 1
        uint64_t sb;
        memcpy(&sb, &scales[k], 8);
        for (int i = 0; i < 8; i++) *p++ = (unsigned char)(sb >> (8 * i));
        *p++ = 1;  // isNTT
        for (int comp = 0; comp < 2; comp++) {
            *p++ = (unsigned char)logN;
            *p++ = (unsigned char)nl;
            const uint64_t *src = cts + (k * 2 + comp) * (size_t)nl * N;
            for (size_t j = 0; j < (size_t)nl * N; j++, p += 8) {
                const uint64_t be = __builtin_bswap64(src[j]);  // ring.WriteCoeffsTo: binary.BigEndian
                memcpy(p, &be, 8);
            }
        }
        ok = fwrite(buf.data(), 1, ctb, f) == ctb;
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) {
        err = std::string("write ") + filename + " failed: " + strerror(errno);
        return -1;
    }
    return 0;
}

// header only: dimensions and the limb count of the first ciphertext (what a caller needs to size its buffers)
int cm_info(const char *filename, int *nrows, int *ncols, int *nl, int *logN, std::string &err) {
    FILE *f = fopen(filename, "rb");
    if (!f) {
        err = std::string("open ") + filename + ": " + strerror(errno);
        return -1;
    }
    unsigned char h[16];
    bool ok = fread(h, 1, 16, f) == 16;
    uint64_t r = 0, c = 0, slen = 0;
    if (ok) {
        r = get_le(h, 4);
        c = get_le(h + 4, 4);
        slen = get_le(h + 8, 8);
        ok = r >= 1 && c >= 1 && slen == r * c * 8;
    }
    unsigned char ct[22];
    if (ok) ok = fseeko(f, (off_t)(16 + slen + 8), SEEK_SET) == 0 && fread(ct, 1, 12, f) == 12;
    fclose(f);
    if (!ok || ct[0] != 2) {
        err = std::string(filename) + ": not a CipherMatrix file of degree-1 ciphertexts";
        return -1;
    }
    *nrows = (int)r;
    *ncols = (int)c;
    *logN = ct[10];
    *nl = ct[11];
    return 0;
}

int cm_load(const char *filename, int logN, uint64_t *cts, double *scales, int nrows, int ncols, int nl, std::string &err) {
    const size_t N = (size_t)1 << logN, ctb = cm_ct_bytes(logN, nl), nct = (size_t)nrows * ncols;
    FILE *f = fopen(filename, "rb");
    if (!f) {
        err = std::string("open ") + filename + ": " + strerror(errno);
        return -1;
    }
    auto fail = [&](const std::string &m) {
        fclose(f);
        err = std::string(filename) + ": " + m;
        return -1;
    };
    unsigned char h[16];
    if (fread(h, 1, 16, f) != 16) return fail("truncated header");
    if (get_le(h, 4) != (uint64_t)nrows || get_le(h + 4, 4) != (uint64_t)ncols) return fail("dimensions do not match the caller's");
    if (get_le(h + 8, 8) != nct * 8) return fail("size table does not match the dimensions");
    std::vector<unsigned char> sizes(nct * 8);
    if (fread(sizes.data(), 1, sizes.size(), f) != sizes.size()) return fail("truncated size table");
    for (size_t k = 0; k < nct; k++)
        if (get_le(&sizes[k * 8], 8) != ctb) return fail("ciphertext " + std::to_string(k) + " is not a degree-1 ciphertext of " + std::to_string(nl) + " limbs (mixed levels)");
    unsigned char lb[8];
    if (fread(lb, 1, 8, f) != 8 || get_le(lb, 8) != nct * ctb) return fail("blob length does not match the size table");
    std::vector<unsigned char> buf(ctb);
    for (size_t k = 0; k < nct; k++) {
        if (fread(buf.data(), 1, ctb, f) != ctb) return fail("truncated ciphertext " + std::to_string(k));
        const unsigned char *p = buf.data();
        if (p[0] != 2) return fail("ciphertext " + std::to_string(k) + " has degree " + std::to_string((int)p[0] - 1));
        const uint64_t sb = get_le(p + 1, 8);
        memcpy(&scales[k], &sb, 8);
        p += 10;
        for (int comp = 0; comp < 2; comp++) {
            if (p[0] != logN || p[1] != nl) return fail("polynomial header (logN, numModuli) mismatch in ciphertext " + std::to_string(k));
            p += 2;
            uint64_t *dst = cts + (k * 2 + comp) * (size_t)nl * N;
            for (size_t j = 0; j < (size_t)nl * N; j++, p += 8) {
                uint64_t be;
                memcpy(&be, p, 8);
                dst[j] = __builtin_bswap64(be);
            }
        }
    }
    fclose(f);
    return 0;
}

}  // namespace sfg
