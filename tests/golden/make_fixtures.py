#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's shipped inputs.

Runs ONLY in the build container (needs /root/reference; the GPU box has no copy).
Imports the reference's NumPy-only ``admp/parser.py`` (no JAX dependency) to read
``water1024.pdb`` / ``mpidwater.xml`` / ``water2.pdb`` exactly as the example
scripts do (examples/water_pol_1024/run_admp.py:19-78), and stores the arrays the
harness, the oracle pins and bench.py need.  Commit the outputs with this script.

    python tests/golden/make_fixtures.py
"""
import os
import sys
import importlib.util

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

spec = importlib.util.spec_from_file_location('ref_parser', os.path.join(REF, 'admp', 'parser.py'))
P = importlib.util.module_from_spec(spec)
spec.loader.exec_module(P)


def load_system(pdb, xml):
    info = P.read_pdb(pdb)
    atomT, resT = P.read_xml(xml)
    atoms, residues = P.init_residues(info['serials'], info['names'], info['resNames'], info['resSeqs'],
                                      info['positions'], info['charges'], atomT, resT)
    n = len(info['serials'])
    # run_admp.py:49-53 : e.nm -> e.A (x10), MPID quadrupole nm^2 -> A^2 and Theta/3 (x300)
    Q_cart = np.vstack([(a.c0, a.dX * 10, a.dY * 10, a.dZ * 10, a.qXX * 300, a.qYY * 300, a.qZZ * 300,
                         a.qXY * 300, a.qXZ * 300, a.qYZ * 300) for a in atoms.values()])
    axis_type = np.array([a.axisType for a in atoms.values()])
    axis_indices = np.vstack([a.axis_indices for a in atoms.values()]).astype(np.int64)
    # run_admp.py:60-70 (float32 cast, A12)
    pol = np.vstack([(a.polarizabilityXX, a.polarizabilityYY, a.polarizabilityZZ) for a in atoms.values()])
    pol = 1000 * np.mean(pol.astype(np.float32), axis=1)
    th = np.vstack([a.thole for a in atoms.values()]).astype(np.float32)
    th = np.mean(th, axis=1)
    cov = P.assemble_covalent(residues, n)
    ii, jj = np.nonzero(cov)
    return dict(positions=np.asarray(info['positions'], dtype=np.float64),
                box=np.asarray(info['box'][:3], dtype=np.float64),
                Q_cart=Q_cart, axis_type=axis_type, axis_indices=axis_indices,
                pol=pol.astype(np.float64), tholes=th.astype(np.float64),
                cov_i=ii.astype(np.int32), cov_j=jj.astype(np.int32), cov_n=cov[ii, jj].astype(np.int8))


def main():
    d = os.path.join(REF, 'examples', 'water_pol_1024')
    s = load_system(os.path.join(d, 'water1024.pdb'), os.path.join(d, 'mpidwater.xml'))
    # MPID induced dipoles, e.nm -> e.A (run_admp.py:73-78)
    s['dipole_mpid'] = 10.0 * np.loadtxt(os.path.join(d, 'dipole_1024'))
    # ref_out rows: "mpid mpid admp" for O atoms, x then y then z (run_admp.py:143-145)
    ro = np.loadtxt(os.path.join(d, 'ref_out'), comments='#')
    s['refout_mpid'] = ro[:, 0].reshape(1024, 3)
    s['refout_admp'] = ro[:, 2].reshape(1024, 3)
    np.savez_compressed(os.path.join(HERE, 'water1024.npz'), **s)

    s2 = load_system(os.path.join(d, 'water2.pdb'), os.path.join(d, 'mpidwater.xml'))
    s2['dipole_mpid'] = 10.0 * np.loadtxt(os.path.join(d, 'dipole_2'))
    np.savez_compressed(os.path.join(HERE, 'water2.npz'), **s2)
    print('wrote', os.path.join(HERE, 'water1024.npz'), os.path.join(HERE, 'water2.npz'))


if __name__ == '__main__':
    main()


# ---- the reference's own force-field input of the api.py front end (a DATA file, 44 lines; examples/openmm_api/
# forcefield.xml): committed verbatim so that the GPU box can parse exactly what a user of the reference passes to
# Hamiltonian(...) (tests/test_api_parsing.py, tests/test_gpu_api.py). residues.xml is not needed: the <Residues> block
# of forcefield.xml carries the bonds.
def copy_forcefield_xml():
    import shutil
    shutil.copyfile(os.path.join(REF, 'examples', 'openmm_api', 'forcefield.xml'), os.path.join(HERE, 'openmm_api_forcefield.xml'))


if __name__ == '__main__':
    copy_forcefield_xml()
