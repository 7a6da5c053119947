"""Mirror of barretenberg::polynomial_arithmetic (reference polynomials/polynomial_arithmetic.hpp:28-39) and of
the parts of evaluation_domain its callers touch (polynomials/evaluation_domain.hpp:9-59).

All transforms are in place on a uint64 array of shape (domain.size, 4) holding Montgomery-form Fr limbs,
natural order in and out, canonical outputs — exactly the reference contract.
"""
from ._lib import default_library


class EvaluationDomain:
    """size / log2_size holder.  The reference object also carries root, root_inverse, domain_inverse, generator and
    host twiddle tables (evaluation_domain.cpp:57-75, :172-178); the GPU library derives all of them on the
    device from the curve constants, so only the size crosses the boundary."""

    def __init__(self, domain_size, library=None):
        if domain_size < 2 or domain_size & (domain_size - 1):
            raise ValueError("evaluation_domain: size must be a power of two >= 2")
        self.size = int(domain_size)
        self.log2_size = self.size.bit_length() - 1
        self.library = library

    def compute_lookup_table(self):
        """No host tables are needed (kept for call-site compatibility)."""
        return None

    def _lib(self):
        return self.library or default_library()


def _run(op, coeffs, domain, constant=None):
    if coeffs.shape[-2] != domain.size:
        raise ValueError("coefficient count %d != domain size %d" % (coeffs.shape[-2], domain.size))
    return domain._lib().ntt(op, coeffs, constant)


def fft(coeffs, domain):
    return _run("fft", coeffs, domain)


def ifft(coeffs, domain):
    return _run("ifft", coeffs, domain)


def fft_with_constant(coeffs, domain, value):
    return _run("fft_with_constant", coeffs, domain, value)


def ifft_with_constant(coeffs, domain, value):
    return _run("ifft_with_constant", coeffs, domain, value)


def coset_fft(coeffs, domain):
    return _run("coset_fft", coeffs, domain)


def coset_fft_with_constant(coeffs, domain, constant):
    return _run("coset_fft_with_constant", coeffs, domain, constant)


def coset_ifft(coeffs, domain):
    return _run("coset_ifft", coeffs, domain)


def compute_lagrange_polynomial_fft(src_domain, target_domain):
    """The target-domain coset evaluations of L_1 (polynomial_arithmetic.cpp:381-476) as a new (target.size, 4) array."""
    return target_domain._lib().compute_lagrange_polynomial_fft(src_domain.log2_size, target_domain.log2_size)
