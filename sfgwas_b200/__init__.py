"""sfgwas_b200 -- B200 (sm_100a) implementation of the genotype x CKKS-ciphertext MatMult hot path of hhcho/sfgwas.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of include/sfgwas_b200.h), ``build.py`` and the
host-side mirror of the reference interface (``gwas.py``).  There is no CPU fallback.
"""
from ._lib import LIB_PATH, SYMBOLS, SfgError, load  # noqa: F401
from .gwas import (CAdd, Ciphertext, CMult, CMultScalar, CryptoParams, CSub, DeviceCipherVector, DiagCache, GenoFileStream,  # noqa: F401
                   InnerProd, InnerSumAll, LoadCipherMatrixFromFile, MaskTrunc, MatMult4Stream, MatMult4StreamCompute, MatMult4StreamPreprocess,
                   QXLazyNormStream, QXtLazyNormStream, QXtLazyNormStreamDevice, RefreshFinish, RefreshGenShares, SaveCipherMatrixToFile, SetRelinKey)
