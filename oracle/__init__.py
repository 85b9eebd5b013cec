"""CPU oracle for the ADMP multipolar-PME hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 CPU restatement (PyTorch on
CPU, so that ``torch.autograd`` plays the role ``jax.grad`` plays in the
reference) of the algorithms in the reference's ``admp/{multipole,spatial,
pairwise,pme,recip,disp_pme}.py``.  Every function cites the reference
``file:line`` it follows.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs, and there only as the checker / the baseline being timed -
never by ``admp_b200`` (the product), which must fail loudly when its CUDA
library is missing.

Parity pin status (see DESIGN.md "Oracle"):
  * the reference itself (JAX, jax_md, OpenMM) cannot be imported in this image,
    so nothing under ``oracle/_ref`` exists;
  * pinned against every known-answer vector the reference's own tests hold for
    this path (``tests/test_multipole.py``, ``tests/test_sptial.py``) and against
    the reference's ``examples/water_pol_1024/ref_out`` induced dipoles
    (ADMP column and MPID column) - see ``tests/test_oracle_pins.py``;
  * energies / forces / virial / parameter gradients have NO fixture in the
    reference. Energies (permanent multipoles, and the polarizable energy_fn(U)
    with Thole damping) and forces are pinned instead against an independent
    formulation of the same physics - exact multipolar Ewald summation in
    Cartesian-tensor form with automatically differentiated kernels
    (``tests/test_oracle_independent_ewald.py``: 6e-8 / 1.6e-7 on the energies,
    forces converging as h^3); virial and parameter gradients remain "parity
    unpinned" beyond the restatement and are covered by finite-difference,
    replica- and Ewald-parameter-invariance self checks
    (``tests/test_oracle_selfchecks.py``);
  * the neighbour pair *set* (third-party jax_md, version unpinned in the
    reference) is "parity unpinned": only the set predicate is restated.
"""

DIELECTRIC = 1389.35455846       # admp/pme.py:16
DEFAULT_THOLE_WIDTH = 0.3        # admp/pme.py:17
POL_CONV = 10.0                  # admp/settings.py:29
MAX_N_POL = 30                   # admp/settings.py:30
