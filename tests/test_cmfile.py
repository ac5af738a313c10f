"""crypto.SaveCipherMatrixToFile / LoadCipherMatrixFromFile (crypto/utilities.go:82-141): the library's
host-side writer / reader against the oracle's line-by-line restatement of the Go code.  No device needed."""
import os

import numpy as np
import pytest


def _matrix(rng, nr, nc, nl, N, Q):
    return [[(np.stack([np.stack([rng.integers(0, Q[l], N, dtype=np.uint64) for l in range(nl)]) for _ in range(2)]),
              float(2.0 ** 30 * (1 + 0.25 * rng.random()))) for _ in range(nc)] for _ in range(nr)]


@pytest.mark.parametrize("nr,nc,nl,logN", [(3, 5, 5, 8), (1, 1, 1, 6), (10, 2, 6, 10)])
def test_cipher_matrix_file_bytes_and_round_trip(tmp_path, nr, nc, nl, logN):
    from oracle.oracle import save_cipher_matrix
    from sfgwas_b200 import Ciphertext, LoadCipherMatrixFromFile, SaveCipherMatrixToFile

    Q = [0x1FFFEC001, 0x3FFF4001, 0x3FFE8001, 0x40020001, 0x40038001, (1 << 56) - 5]
    rng = np.random.default_rng(nr * 100 + nc)
    cm = _matrix(rng, nr, nc, nl, 1 << logN, Q)
    f_ref, f_lib = str(tmp_path / "ref.bin"), str(tmp_path / "lib.bin")
    save_cipher_matrix(cm, f_ref)
    SaveCipherMatrixToFile(None, [[Ciphertext(v, sc) for v, sc in row] for row in cm], f_lib)
    with open(f_ref, "rb") as a, open(f_lib, "rb") as b:
        assert a.read() == b.read()
    assert os.path.getsize(f_lib) == 4 + 4 + 8 + 8 * nr * nc + 8 + nr * nc * (10 + 2 * (2 + nl * (8 << logN)))
    back = LoadCipherMatrixFromFile(None, f_ref)
    assert len(back) == nr and len(back[0]) == nc
    for i in range(nr):
        for j in range(nc):
            assert (back[i][j].value == cm[i][j][0]).all() and back[i][j].scale == cm[i][j][1] and back[i][j].Level() == nl - 1


def test_cipher_matrix_file_errors(tmp_path):
    from sfgwas_b200 import Ciphertext, LoadCipherMatrixFromFile, SaveCipherMatrixToFile, SfgError

    with pytest.raises(SfgError):  # the reference log.Fatal's on a missing file (crypto/utilities.go:117-120)
        LoadCipherMatrixFromFile(None, str(tmp_path / "missing.bin"))
    rng = np.random.default_rng(0)
    v = rng.integers(0, 1 << 30, (2, 3, 64), dtype=np.uint64)
    f = str(tmp_path / "t.bin")
    SaveCipherMatrixToFile(None, [[Ciphertext(v, 1.0), Ciphertext(v, 2.0)]], f)
    with open(f, "r+b") as fh:  # truncate: the last ciphertext is incomplete
        fh.truncate(os.path.getsize(f) - 100)
    with pytest.raises(SfgError):
        LoadCipherMatrixFromFile(None, f)
    with pytest.raises(SfgError):  # mixed levels are rejected by the flat writer
        SaveCipherMatrixToFile(None, [[Ciphertext(v, 1.0), Ciphertext(v[:, :2], 1.0)]], f)
