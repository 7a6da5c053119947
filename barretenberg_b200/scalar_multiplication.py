"""Mirror of barretenberg::scalar_multiplication (reference curves/bn254/scalar_multiplication.hpp:41,60-61,88-96).

Points are the reference's interleaved 2n table [P_0, phi(P_0), P_1, phi(P_1), ...] as a uint64 array (2n, 8);
scalars are Montgomery-form Fr limbs (n, 4).  Results are Jacobian triples (12 uint64) in the normalised form
batched_scalar_multiplications leaves in multiplication_state.output (z = fq::one, canonical x, y) or with the
infinity flag (bit 63 of y limb 3) set.
"""
from dataclasses import dataclass, field

import numpy as np

from ._lib import default_library


def generate_pippenger_point_table(points, library=None):
    """(n, 8) affine points -> (2n, 8) table, entry 2i+1 = (beta * x_i, -y_i)  (scalar_multiplication.cpp:131-140)."""
    return (library or default_library()).generate_pippenger_point_table(points)


def pippenger(scalars, points, num_initial_points=None, forced_bucket_width=0, library=None):
    """sum_i scalars[i] * P_i.  `forced_bucket_width` is accepted for signature compatibility; the result does not
    depend on it (SURVEY.md §8 note 3).  Unlike the reference the returned Jacobian point is already normalised."""
    del forced_bucket_width
    n = int(scalars.shape[0]) if num_initial_points is None else int(num_initial_points)
    return (library or default_library()).msm(scalars, points, n)


@dataclass
class MultiplicationState:
    """multiplication_state (scalar_multiplication.hpp:88-94)."""
    points: np.ndarray
    scalars: np.ndarray
    num_elements: int
    output: np.ndarray = field(default_factory=lambda: np.zeros(12, dtype=np.uint64))


def batched_scalar_multiplications(mul_state, num_batches=None, library=None):
    states = list(mul_state)[: num_batches if num_batches is not None else None]
    if not states:
        return
    n = states[0].num_elements
    if any(s.num_elements != n for s in states):
        # reference: printf("... each scalar mul must be same size") and return (scalar_multiplication.cpp:677-685)
        raise ValueError("batched_scalar_multiplications err: each scalar mul must be same size.")
    lib = library or default_library()
    outs = lib.msm_batched([s.scalars[:n] for s in states], [s.points for s in states])
    for s, o in zip(states, outs):
        s.output = o
