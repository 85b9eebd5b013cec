"""Force-field front end (admp_b200/api.py - the potential_fn(positions, box, pairs, params) convention of
admp/api.py) against the oracle. The XML / PDB inputs are written by the test from the committed
water fixture (schema of examples/openmm_api/forcefield.xml; nothing is read from the reference tree)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dispersion as odisp                      # noqa: E402
from oracle import fixtures, pairlist                       # noqa: E402
from oracle import realspace as orc                         # noqa: E402
from oracle.shortrange import TT_damping_qq_c6_kernel as o_tt, generate_pairwise_interaction as o_pairwise   # noqa: E402

RTOL = 1e-6
# per-type dispersion / exchange parameters in the XML's units (O, H): A kJ/mol, B nm^-1, Q e, C6 ... (kJ/mol nm^p)
DISP = {'380': dict(A=1203470.743, B=37.81265679, Q=-0.741706, C6=0.001383816, C8=7.27065e-05, C10=1.8076465e-6),
        '381': dict(A=83.2283563, B=37.78544799, Q=0.370853, C6=5.7929e-05, C8=1.416624e-06, C10=2.26525e-08)}


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-300)


def _write_inputs(tmp_path, s):
    """forcefield.xml + water.pdb for the (carved) fixture: multipoles, polarizabilities and Thole widths are the
    fixture's own (converted back to the XML's nm units), the dispersion block is DISP."""
    Qc = s.Q_cart
    o, h = Qc[0], Qc[1]

    def atom(tp, kz, kx, q):
        v = dict(c0=q[0], dX=q[1] / 10, dY=q[2] / 10, dZ=q[3] / 10, qXX=q[4] / 300, qYY=q[5] / 300, qZZ=q[6] / 300,
                 qXY=q[7] / 300, qXZ=q[8] / 300, qYZ=q[9] / 300)
        return '<Atom type="%s" kz="%s" kx="%s" %s/>' % (tp, kz, kx, ' '.join('%s="%.12g"' % kv for kv in v.items()))

    scales = ' '.join('%sScale1%d="%s"' % (c, i, '0.00' if i < 5 else '1.00') for c in 'mpd' for i in range(2, 7))
    polO, thO = float(s.pol[0]) / 1000, float(s.tholes[0])
    xml = '''<ForceField>
 <AtomTypes><Type name="380" class="OW" element="O" mass="15.999"/><Type name="381" class="HW" element="H" mass="1.008"/></AtomTypes>
 <Residues><Residue name="HOH"><Atom name="H1" type="381"/><Atom name="H2" type="381"/><Atom name="O" type="380"/>
   <Bond from="0" to="2"/><Bond from="1" to="2"/></Residue></Residues>
 <ADMPDispForce %s>%s</ADMPDispForce>
 <ADMPPmeForce lmax="2" pmax="10" %s>
   %s
   %s
   <Polarize type="380" polarizabilityXX="%.9g" polarizabilityYY="%.9g" polarizabilityZZ="%.9g" thole="%.9g"/>
   <Polarize type="381" polarizabilityXX="0.0" polarizabilityYY="0.0" polarizabilityZZ="0.0" thole="0.0"/>
 </ADMPPmeForce>
</ForceField>''' % (' '.join('mScale1%d="%s"' % (i, '0.00' if i < 5 else '1.00') for i in range(2, 7)),
                    ''.join('<Atom type="%s" %s/>' % (t, ' '.join('%s="%.10g"' % kv for kv in d.items())) for t, d in DISP.items()),
                    scales, atom('380', '-381', '-381', o), atom('381', '380', '381', h), polO, polO, polO, thO)
    ff = tmp_path / 'forcefield.xml'
    ff.write_text(xml)
    L = torch.diagonal(s.box).numpy()
    lines = ['CRYST1%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f P 1           1' % (L[0], L[1], L[2], 90, 90, 90)]
    names = ['O', 'H1', 'H2']
    pos = s.positions.numpy()
    for a in range(s.n_atoms):
        lines.append('HETATM%5d %-4s HOH A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s' % (
            (a + 1) % 100000, names[a % 3], (a // 3 + 1) % 10000, pos[a, 0], pos[a, 1], pos[a, 2], names[a % 3][0]))
    pdb = tmp_path / 'water.pdb'
    pdb.write_text('\n'.join(lines) + '\nEND\n')
    return str(ff), str(pdb)


def test_hamiltonian_potentials_match_oracle(tmp_path):
    from admp_b200.api import Hamiltonian, PDBFile
    s = fixtures.lattice_water(4, 3.15, seed=5)               # 64 waters, 12.6 A box, liquid-like (the SCF converges)
    ff, pdbfile = _write_inputs(tmp_path, s)
    H = Hamiltonian(ff)
    pdb = PDBFile(pdbfile)
    assert pdb.topology.getNumAtoms() == s.n_atoms and np.allclose(pdb.topology.box, s.box.numpy())
    disp_gen, pme_gen = H.getGenerators()
    rc = 5.0
    pot_disp, pot_pme = H.createPotential(pdb.topology, nonbondedCutoff=rc)
    # topology-derived tables: axis types / anchors of water, covalent map 1 (O-H) and 2 (H-H)
    assert pme_gen.axis_types.tolist() == s.axis_type.tolist()
    assert np.array_equal(pme_gen.axis_indices[:, :2], s.axis_indices[:, :2])
    positions = pdb.positions                                 # 3 decimals: the oracle gets the same rounded coordinates
    pairs, _ = pairlist.build_pairs(positions, s.box.numpy(), rc)
    tp = torch.tensor(positions, dtype=torch.float64)

    # ---- multipolar polarizable PME: energy, induced dipoles and d/d(params)
    for k in ('mScales', 'Q_local'):
        pme_gen.params[k].requires_grad_(True)
    E = pot_pme(positions, s.box, pairs, pme_gen.params)
    E.backward()
    f = pme_gen.force
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, rc, 1e-5, 2, lpol=True)
    assert (f.kappa, f.K1, f.K2, f.K3) == (ref.kappa, ref.K1, ref.K2, ref.K3)
    mS = s.mScales.clone().requires_grad_(True)
    Ql = s.Q_local.clone().requires_grad_(True)
    pol32 = torch.tensor((1000 * (s.pol.numpy() / 1000).astype(np.float32)).astype(np.float64))     # upstream's float32 cast (A12)
    Eo = ref.get_energy(tp, s.box, pairs, Ql, pol32, s.tholes, mS, s.pScales, s.dScales,
                        U_init=torch.zeros(s.n_atoms, 3, dtype=torch.float64))
    go = torch.autograd.grad(Eo, [mS, Ql])
    assert f.n_cycle == ref.n_cycle
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item()), (E.item(), Eo.item())
    assert rel(pme_gen.params['mScales'].grad, go[0]) < RTOL
    assert rel(pme_gen.params['Q_local'].grad, go[1]) < RTOL

    # ---- dispersion: Tang-Toennies short range minus dispersion PME, gradient w.r.t. the raw XML parameters
    for k in ('mScales', 'A', 'C6'):
        disp_gen.params[k].requires_grad_(True)
    Ed = pot_disp(positions, s.box, pairs, disp_gen.params)
    Ed.backward()
    idx = torch.as_tensor(disp_gen.map_atomtype)
    raw = {k: torch.tensor([DISP[t][k] for t in ('380', '381')], dtype=torch.float64, requires_grad=True) for k in DISP['380']}
    mS2 = s.mScales.clone().requires_grad_(True)
    a_l, b_l, q_l = raw['A'][idx] / 2625.5, raw['B'][idx] * 0.0529177249, raw['Q'][idx]
    c = torch.stack([torch.sqrt(raw['C6'][idx] * 1e6), torch.sqrt(raw['C8'][idx] * 1e8), torch.sqrt(raw['C10'][idx] * 1e10)], 1)
    kappa, K1, K2, K3 = orc.setup_ewald_parameters(rc, 1e-5, s.box)
    E_sr = o_pairwise(o_tt, s.covalent_map)(tp, s.box, pairs, mS2, a_l, b_l, q_l, c[:, 0])
    E_lr = odisp.energy_disp_pme(tp, s.box, pairs, c, mS2, s.covalent_map, kappa, K1, K2, K3, 10)
    Edo = E_sr - E_lr
    gdo = torch.autograd.grad(Edo, [mS2, raw['A'], raw['C6']])
    assert abs(Ed.item() - Edo.item()) < RTOL * abs(Edo.item()), (Ed.item(), Edo.item())
    assert rel(disp_gen.params['mScales'].grad, gdo[0]) < RTOL
    assert rel(disp_gen.params['A'].grad, gdo[1]) < RTOL
    assert rel(disp_gen.params['C6'].grad, gdo[2]) < RTOL


def test_reference_own_forcefield_xml_on_its_1024_water_box(tmp_path):
    """The reference's own front-end input (examples/openmm_api/forcefield.xml + the topology / positions of its
    water1024.pdb, run.py:15-45): both potentials against the oracle evaluated with the parameters of the XML."""
    import os
    from admp_b200.api import Hamiltonian, PDBFile
    from admp_b200.multipole import convert_cart2harm, rot_ind_global2local
    ff = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'openmm_api_forcefield.xml')
    s = fixtures.water1024()
    _, pdbfile = _write_inputs(tmp_path, s)
    H = Hamiltonian(ff)
    pdb = PDBFile(pdbfile)
    disp_gen, pme_gen = H.getGenerators()
    rc = 4.0
    pot_disp, pot_pme = H.createPotential(pdb.topology, nonbondedCutoff=rc)
    positions = pdb.positions
    pairs, _ = pairlist.build_pairs(positions, s.box.numpy(), rc)
    tp = torch.tensor(positions, dtype=torch.float64)
    # parameters exactly as the XML states them (admp/api.py:319-338)
    cart = np.array([[-1.0614, 0, 0, -0.023671684 * 10, 0.000150963 * 300, 0.00008707 * 300, -0.000238034 * 300, 0, 0, 0],
                     [0.5307, 0, 0, 0, 0, 0, 0, 0, 0, 0]])
    Ql = torch.tensor(np.asarray(convert_cart2harm(cart, 2)))[torch.tensor([0, 1, 1] * 1024)]
    polO = float((1000 * np.full((1, 3), 0.00088, dtype=np.float32).mean(axis=1)).astype(np.float64)[0])   # upstream's float32 arithmetic (A12)
    pol = torch.tensor(np.tile([polO, 0.0, 0.0], 1024))
    th = torch.tensor(np.tile([8.0, 0.0, 0.0], 1024))
    assert rel(pme_gen.params['Q_local'], Ql) < 1e-12 and rel(pme_gen.params['pol'], pol) < 1e-12
    E = pot_pme(positions, s.box, pairs, pme_gen.params)
    f = pme_gen.force
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, rc, 1e-5, 2, lpol=True)
    assert (f.K1, f.K2, f.K3) == (ref.K1, ref.K2, ref.K3)
    Eo = ref.get_energy(tp, s.box, pairs, Ql, pol, th, s.mScales, s.pScales, s.dScales)
    assert f.n_cycle == ref.n_cycle and f.lconverg == ref.lconverg
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item()), (E.item(), Eo.item())
    Ed = pot_disp(positions, s.box, pairs, disp_gen.params)
    idx = torch.as_tensor(disp_gen.map_atomtype)
    raw = {k: torch.tensor([DISP[t][k] for t in ('380', '381')], dtype=torch.float64) for k in DISP['380']}     # DISP = the XML's block
    a_l, b_l, q_l = raw['A'][idx] / 2625.5, raw['B'][idx] * 0.0529177249, raw['Q'][idx]
    c = torch.stack([torch.sqrt(raw['C6'][idx] * 1e6), torch.sqrt(raw['C8'][idx] * 1e8), torch.sqrt(raw['C10'][idx] * 1e10)], 1)
    kappa, K1, K2, K3 = orc.setup_ewald_parameters(rc, 1e-5, s.box)
    Edo = o_pairwise(o_tt, s.covalent_map)(tp, s.box, pairs, s.mScales, a_l, b_l, q_l, c[:, 0]) \
        - odisp.energy_disp_pme(tp, s.box, pairs, c, s.mScales, s.covalent_map, kappa, K1, K2, K3, 10)
    assert abs(Ed.item() - Edo.item()) < RTOL * abs(Edo.item()), (Ed.item(), Edo.item())
    # rot_ind_global2local (admp/multipole.py:80-89) on the device: R[zxy][:, zxy] . U
    fr = f.construct_local_frames(positions, s.box)[:5].cpu()
    U = torch.tensor(np.random.default_rng(0).normal(size=(5, 3)))
    zxy = [2, 0, 1]
    want = torch.stack([fr[k][zxy][:, zxy] @ U[k] for k in range(5)])
    assert rel(rot_ind_global2local(U, fr), want) < 1e-12
