"""barretenberg_b200 — B200-native MSM and NTT behind Barretenberg's call signatures.

The product is `libbbgpu.so` (hand-written sm_100a CUDA kernels behind the C ABI of `include/bbgpu.h`);
this package is the thin Python mirror of the reference interface used by the tests and `bench.py`:

    scalar_multiplication.pippenger / batched_scalar_multiplications / generate_pippenger_point_table
        (reference curves/bn254/scalar_multiplication.hpp:41,60-61,88-96)
    polynomial_arithmetic.fft / ifft / coset_fft / coset_ifft / fft_with_constant / ifft_with_constant /
        coset_fft_with_constant on an EvaluationDomain   (reference polynomials/polynomial_arithmetic.hpp:28-39)

There is no CPU fallback: importing works anywhere, but every compute call raises BbgError unless the CUDA
library is built and a GPU is present.
"""
from ._lib import BbgError, Library, default_library, library_path  # noqa: F401
from . import polynomial_arithmetic, scalar_multiplication  # noqa: F401
from .polynomial_arithmetic import EvaluationDomain  # noqa: F401

__all__ = ["BbgError", "Library", "default_library", "library_path", "polynomial_arithmetic",
           "scalar_multiplication", "EvaluationDomain"]
