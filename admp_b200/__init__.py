"""admp_b200 - B200-native (sm_100a) implementation of ADMP's multipolar PME hot path.

Public surface mirrors the reference package ``admp`` for that path:
    admp_b200.settings                       (PRECISION, DO_JIT, POL_CONV, MAX_N_POL)
    admp_b200.pme.ADMPPmeForce, setup_ewald_parameters
    admp_b200.disp_pme.ADMPDispPmeForce
    admp_b200.pairwise.generate_pairwise_interaction, TT_damping_qq_c6_kernel
    admp_b200.recip.generate_pme_recip, Ck_1, Ck_6, Ck_8, Ck_10
    admp_b200.multipole / admp_b200.spatial  (helpers the reference's tests exercise)
    admp_b200.neighbor.neighbor_list         (replaces the jax_md call of the scripts)
Everything numerical runs in hand-written CUDA kernels behind the C ABI of
include/admp_b200.h (admp_b200/lib/libadmp_b200.so); importing this package never
falls back to a CPU implementation.
"""
from . import settings  # noqa: F401

__all__ = ['settings', 'pme', 'disp_pme', 'pairwise', 'recip', 'multipole', 'spatial', 'neighbor', 'covalent']
