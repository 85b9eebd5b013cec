"""Force-field front end - the ``potential_fn(positions, box, pairs, params)`` convention of admp/api.py
(SURVEY 8(f) rank 1), without OpenMM.

The reference builds its potentials by subclassing ``openmm.app.ForceField`` (admp/api.py:469-488): OpenMM parses the
XML, matches residue templates and hands every ``<ADMP...Force>`` element to a generator whose ``createForce``
prepares per-atom parameters and returns a closure. OpenMM is not installable here, so this module carries the small
part of that machinery the path needs - an XML reader for ``<AtomTypes>``, ``<Residues>``, ``<ADMPDispForce>`` and
``<ADMPPmeForce>`` (schema of examples/openmm_api/forcefield.xml), a PDB reader and residue-template matching - and
keeps the reference's object surface on top of it:

    H = Hamiltonian('forcefield.xml')
    pdb = PDBFile('water1024.pdb')
    disp_generator, pme_generator = H.getGenerators()
    pot_disp, pot_pme = H.createPotential(pdb.topology, nonbondedCutoff=4.0)          # Angstrom
    E = pot_pme(positions, box, pairs, pme_generator.params)                          # Angstrom, kJ/mol
    E.backward()  ->  pme_generator.params['mScales'].grad, ['Q_local'].grad, ...      # jax.grad(pot, argnums=3)

Units follow the reference scripts after their ``* 10`` conversions: positions and box in Angstrom (``PDBFile``
returns Angstrom directly). ``params`` are torch tensors on the GPU; every entry is differentiable through the
analytic-adjoint kernels (unit conversions A/2625.5, B*0.0529177249, sqrt(C6*1e6) ... are plain tensor arithmetic,
admp/api.py:185-192).
"""
import xml.etree.ElementTree as ET

import numpy as np
import torch

from ._ctx import device as _device
from .covalent import SparseCovalentMap
from .disp_pme import ADMPDispPmeForce
from .multipole import convert_cart2harm
from .pairwise import TT_damping_qq_c6_kernel, generate_pairwise_interaction
from .pme import ADMPPmeForce


# ----------------------------------------------------------------------------------- topology
class Atom:
    def __init__(self, index, name, element, residue):
        self.index, self.name, self.element, self.residue = index, name, element, residue


class Residue:
    def __init__(self, index, name):
        self.index, self.name, self._atoms = index, name, []

    def atoms(self):
        return iter(self._atoms)


class Topology:
    """Atoms grouped in residues plus the periodic cell (Angstrom, rows = lattice vectors). Bonds are created from
    the residue templates of the force field (the reference loads them with ``Topology.loadBondDefinitions``)."""

    def __init__(self):
        self._atoms, self._residues, self.box = [], [], None

    def addResidue(self, name):
        r = Residue(len(self._residues), name)
        self._residues.append(r)
        return r

    def addAtom(self, name, element, residue):
        a = Atom(len(self._atoms), name, element, residue)
        self._atoms.append(a)
        residue._atoms.append(a)
        return a

    def atoms(self):
        return iter(self._atoms)

    def residues(self):
        return iter(self._residues)

    def getNumAtoms(self):
        return len(self._atoms)

    def getPeriodicBoxVectors(self):
        return self.box


class PDBFile:
    """Minimal PDB reader: CRYST1 (cell), ATOM / HETATM records. ``positions`` (n, 3) and ``topology.box`` in
    Angstrom."""

    def __init__(self, path):
        top = Topology()
        pos = []
        last = None
        with open(path) as f:
            for line in f:
                rec = line[:6]
                if rec == 'CRYST1':
                    a, b, c = float(line[6:15]), float(line[15:24]), float(line[24:33])
                    al, be, ga = (np.deg2rad(float(line[33 + 7 * k:40 + 7 * k])) for k in range(3))
                    bx = b * np.cos(ga)
                    by = b * np.sin(ga)
                    cx = c * np.cos(be)
                    cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
                    cz = np.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
                    box = np.array([[a, 0, 0], [bx, by, 0], [cx, cy, cz]], dtype=np.float64)
                    box[np.abs(box) < 1e-10] = 0.0
                    top.box = box
                elif rec in ('ATOM  ', 'HETATM'):
                    name, resname = line[12:16].strip(), line[17:20].strip()
                    key = (line[21], line[22:27], resname)
                    if key != last:
                        res = top.addResidue(resname)
                        last = key
                    el = line[76:78].strip() if len(line) >= 78 else ''
                    top.addAtom(name, el or name[0], res)
                    pos.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
        self.topology = top
        self.positions = np.asarray(pos, dtype=np.float64)


class _Data:
    """What OpenMM's ForceField._SystemData gives the generators: atoms, their types, bonds."""

    def __init__(self, topology, atom_types, bonds):
        self.atoms = list(topology.atoms())
        self.atomType = {a: atom_types[a.index] for a in self.atoms}
        self.bonds = bonds                      # list of (i, j)


def build_covalent_map(data, max_neighbor):
    """admp/api.py:24-42: topological distance (1 ... max_neighbor) between bonded atoms, 0 otherwise - as a sparse map
    (the dense Na x Na matrix of the reference is 4.9 TB at 786k atoms). Breadth-first search from every atom."""
    n = len(data.atoms)
    nbrs = [[] for _ in range(n)]
    for i, j in data.bonds:
        nbrs[i].append(j)
        nbrs[j].append(i)
    ci, cj, cn = [], [], []
    for i in range(n):
        if not nbrs[i]:
            continue
        dist = {i: 0}
        frontier = [i]
        for d in range(1, max_neighbor + 1):
            nxt = []
            for a in frontier:
                for b in nbrs[a]:
                    if b not in dist:
                        dist[b] = d
                        nxt.append(b)
            frontier = nxt
            if not frontier:
                break
        for b, d in dist.items():
            if d > 0:
                ci.append(i)
                cj.append(b)
                cn.append(d)
    return SparseCovalentMap.from_pairs(n, np.asarray(ci, dtype=np.int64), np.asarray(cj, dtype=np.int64),
                                        np.asarray(cn, dtype=np.int8))


ZThenX, Bisector, ZBisect, ThreeFold, Zonly, NoAxisType = 0, 1, 2, 3, 4, 5


def set_axis_type(map_atomtypes, types, params):
    """admp/api.py:44-116 (the AMOEBA / MPID anchor convention): from the kz / kx / ky type strings of every atom's
    type (a leading '-' marks bisector-style anchors) to the local-frame axis type and the anchor TYPE names."""
    axis_types, axis_indices = [], []
    for i in map_atomtypes:
        k = []
        neg = []
        for name in ('kz', 'kx', 'ky'):
            v = params[name][i]
            if v != '':
                neg.append(v.startswith('-'))
                k.append(v[1:] if v.startswith('-') else v)
        kz, kx, ky = (k + ['', '', ''])[:3]
        kzn, kxn, kyn = (neg + [False, False, False])[:3]
        t = ZThenX
        if not kz:
            t = NoAxisType
        if kz and not kx:
            t = Zonly
        if (kz and kzn) or (kx and kxn):
            t = Bisector
        if kx and kxn and ky and kyn:
            t = ZBisect
        if kz and kzn and kx and kxn and ky and kyn:
            t = ThreeFold
        axis_types.append(t)
        axis_indices.append([types[i], kz, kx, ky])
    return np.array(axis_types), axis_indices


def _map_axis_indices(data, axis_type_names):
    """admp/api.py:398-415: anchor type names -> atom indices inside the atom's own residue (first unused atom of the
    requested type); unused / unmatched anchors are -1."""
    out = []
    for a in data.atoms:
        want = [x if x != '' else -1 for x in axis_type_names[a.index][1:]]
        for other in a.residue._atoms:
            if other is a:
                continue
            for k in range(len(want)):
                if isinstance(want[k], str) and want[k] == data.atomType[other]:
                    want[k] = other.index
                    break
        out.append([w if not isinstance(w, str) else -1 for w in want])
    return np.array(out, dtype=np.int64)


def _cutoff_angstrom(x):
    return float(x.value_in_unit_angstrom()) if hasattr(x, 'value_in_unit_angstrom') else float(x)


def _param(values, dtype=torch.float64):
    return torch.tensor(np.asarray(values, dtype=np.float64), dtype=dtype, device=_device())


# ----------------------------------------------------------------------------------- generators
class ADMPDispGenerator:
    """admp/api.py:119-224: Tang-Toennies damped short-range part minus the dispersion PME long-range part."""

    def __init__(self, hamiltonian):
        self.ff = hamiltonian
        self.params = {'A': [], 'B': [], 'Q': [], 'C6': [], 'C8': [], 'C10': []}
        self._jaxPotential = None
        self.types = []
        self.ethresh = 1.0e-5
        self.pmax = 10

    def registerAtomType(self, atom):
        self.types.append(atom['type'])
        for k in ('A', 'B', 'Q', 'C6', 'C8', 'C10'):
            self.params[k].append(float(atom[k]))

    @staticmethod
    def parseElement(element, hamiltonian):
        g = ADMPDispGenerator(hamiltonian)
        hamiltonian.registerGenerator(g)
        g.params['mScales'] = [float(element.attrib['mScale1%d' % i]) for i in range(2, 7)]
        for atomtype in element.findall('Atom'):
            g.registerAtomType(atomtype.attrib)
        g._raw = {k: np.asarray(v, dtype=np.float64) for k, v in g.params.items()}
        g.types = np.array(g.types)

    def createForce(self, data, box, rc):
        self.params = {k: _param(v) for k, v in self._raw.items()}
        n_atoms = len(data.atoms)
        map_atomtype = np.array([int(np.where(self.types == data.atomType[a])[0][0]) for a in data.atoms])
        idx = torch.as_tensor(map_atomtype, device=_device())
        covalent_map = build_covalent_map(data, 6)
        force = ADMPDispPmeForce(box, covalent_map, rc, self.ethresh, self.pmax)
        pot_fn_lr = force.get_energy
        pot_fn_sr = generate_pairwise_interaction(TT_damping_qq_c6_kernel, covalent_map, static_args={})
        self.force, self.map_atomtype, self.n_atoms = force, map_atomtype, n_atoms

        def potential_fn(positions, box, pairs, params):
            mScales = params['mScales']
            a_list = params['A'][idx] / 2625.5                       # kJ/mol -> Hartree (admp/api.py:185-187)
            b_list = params['B'][idx] * 0.0529177249                 # nm^-1 -> Bohr^-1
            q_list = params['Q'][idx]
            c6 = torch.sqrt(params['C6'][idx] * 1e6)
            c8 = torch.sqrt(params['C8'][idx] * 1e8)
            c10 = torch.sqrt(params['C10'][idx] * 1e10)
            c_list = torch.stack((c6, c8, c10), dim=1)               # (Na, 3) as ADMPDispPmeForce expects
            E_sr = pot_fn_sr(positions, box, pairs, mScales, a_list, b_list, q_list, c6)
            E_lr = pot_fn_lr(positions, box, pairs, c_list, mScales)
            return E_sr - E_lr

        self._jaxPotential = potential_fn

    def getJaxPotential(self):
        return self._jaxPotential


_PME_FIELDS = ('c0', 'dX', 'dY', 'dZ', 'qXX', 'qXY', 'qYY', 'qXZ', 'qYZ', 'qZZ', 'thole', 'polarizabilityXX',
               'polarizabilityYY', 'polarizabilityZZ')


class ADMPPmeGenerator:
    """admp/api.py:230-455: multipolar (polarizable) PME from the MPID-style ``<ADMPPmeForce>`` element."""

    def __init__(self, hamiltonian):
        self.ff = hamiltonian
        self.kStrings = {'kz': [], 'kx': [], 'ky': []}
        self._input_params = {k: [] for k in _PME_FIELDS}
        self._jaxPotential = None
        self.types = []
        self.ethresh = 1.0e-5
        self.params = {}
        self.lpol = False
        self.ref_dip = ''

    def registerAtomType(self, atom):
        atom = dict(atom)
        self.types.append(atom.pop('type'))
        for k in ('kz', 'kx', 'ky'):
            self.kStrings[k].append(atom.pop(k, ''))
        for k in _PME_FIELDS:
            self._input_params[k].append(float(atom.get(k, 0.0)))      # octupoles (oXXX ...) are parsed upstream but unused

    @staticmethod
    def parseElement(element, hamiltonian):
        g = ADMPPmeGenerator(hamiltonian)
        g.lmax = int(element.attrib.get('lmax'))
        g.pmax = int(element.attrib.get('pmax'))
        hamiltonian.registerGenerator(g)
        g._scales = {s: [float(element.attrib['%s1%d' % (s[:-1], i)]) for i in range(2, 7)] for s in ('mScales', 'pScales', 'dScales')}
        if element.findall('Polarize'):
            g.lpol = True
        for atomType in element.findall('Atom'):
            attrib = dict(atomType.attrib)
            for pol in element.findall('Polarize'):
                if pol.attrib['type'] == attrib['type']:
                    attrib.update(pol.attrib)
                    break
            g.registerAtomType(attrib)
        g._input_params = {k: np.asarray(v, dtype=np.float64) for k, v in g._input_params.items()}
        g.types = np.array(g.types)

    def createForce(self, data, box, rc):
        n_atoms = len(data.atoms)
        m = np.array([int(np.where(self.types == data.atomType[a])[0][0]) for a in data.atoms])
        p = self._input_params
        Q = np.zeros((n_atoms, 10))
        Q[:, 0] = p['c0'][m]
        for col, key in ((1, 'dX'), (2, 'dY'), (3, 'dZ')):
            Q[:, col] = p[key][m] * 10                                # e nm -> e A (admp/api.py:322-324)
        for col, key in ((4, 'qXX'), (5, 'qYY'), (6, 'qZZ'), (7, 'qXY'), (8, 'qXZ'), (9, 'qYZ')):
            Q[:, col] = p[key][m] * 300                               # nm^2 -> A^2 and MPID's Theta / 3
        # isotropic polarizability through float32, as upstream (admp/api.py:332-338)
        pol3 = np.stack([p['polarizabilityXX'][m], p['polarizabilityYY'][m], p['polarizabilityZZ'][m]], 1).astype(np.float32)
        pol = (1000 * pol3.mean(axis=1)).astype(np.float64)
        tholes = p['thole'][m].astype(np.float32).astype(np.float64)
        self.params = {k: _param(v) for k, v in self._scales.items()}
        self.params['Q_local'] = _param(convert_cart2harm(Q, 2)[:, :(self.lmax + 1) ** 2])
        self.params['pol'] = _param(pol)
        self.params['tholes'] = _param(tholes)
        covalent_map = build_covalent_map(data, 6)
        self.axis_types, names = set_axis_type(m, self.types, self.kStrings)
        self.axis_indices = _map_axis_indices(data, names)
        pme_force = ADMPPmeForce(box, self.axis_types, self.axis_indices, covalent_map, rc, self.ethresh, self.lmax, self.lpol)
        self.params['U_ind'] = pme_force.U_ind if self.lpol else None
        self.force, self.map_atomtype = pme_force, m

        def potential_fn(positions, box, pairs, params):
            if self.lpol:
                return pme_force.get_energy(positions, box, pairs, params['Q_local'], params['pol'], params['tholes'],
                                            params['mScales'], params['pScales'], params['dScales'], U_init=params.get('U_ind'))
            return pme_force.get_energy(positions, box, pairs, params['Q_local'], params['mScales'])

        self._jaxPotential = potential_fn

    def getJaxPotential(self):
        return self._jaxPotential


parsers = {'ADMPDispForce': ADMPDispGenerator.parseElement, 'ADMPPmeForce': ADMPPmeGenerator.parseElement}


# ----------------------------------------------------------------------------------- Hamiltonian
class Hamiltonian:
    """admp/api.py:469-488 with the part of ``openmm.app.ForceField`` it relies on: atom types, residue templates
    (matched by residue name and atom names) and the generator registry."""

    def __init__(self, xmlname):
        root = ET.parse(xmlname).getroot()
        self._atomTypes = {t.attrib['name']: dict(t.attrib) for t in root.findall('./AtomTypes/Type')}
        self._templates = {}
        for r in root.findall('./Residues/Residue'):
            atoms = [(a.attrib['name'], a.attrib['type']) for a in r.findall('Atom')]
            bonds = []
            for b in r.findall('Bond'):
                if 'from' in b.attrib:
                    bonds.append((int(b.attrib['from']), int(b.attrib['to'])))
                else:
                    names = [a[0] for a in atoms]
                    bonds.append((names.index(b.attrib['atomName1']), names.index(b.attrib['atomName2'])))
            self._templates[r.attrib['name']] = (atoms, bonds)
        self._forces = []
        self._potentials = []
        for child in root:
            if child.tag in parsers:
                parsers[child.tag](child, self)

    def registerGenerator(self, generator):
        self._forces.append(generator)

    def getGenerators(self):
        return self._forces

    def _match(self, topology):
        types, bonds = {}, []
        for res in topology.residues():
            if res.name not in self._templates:
                raise ValueError('no residue template named %r in the force field' % res.name)
            atoms, tbonds = self._templates[res.name]
            by_name = {a.name: a for a in res._atoms}
            if len(by_name) != len(atoms) or any(nm not in by_name for nm, _ in atoms):
                raise ValueError('residue %d (%s) does not match its template (atoms %s)' % (res.index, res.name, [a.name for a in res._atoms]))
            for nm, tp in atoms:
                types[by_name[nm].index] = tp
            for i, j in tbonds:
                bonds.append((by_name[atoms[i][0]].index, by_name[atoms[j][0]].index))
        return types, bonds

    def createPotential(self, topology, nonbondedMethod=None, nonbondedCutoff=10.0):
        """Returns one ``potential_fn(positions, box, pairs, params)`` per force element, in file order.
        ``nonbondedCutoff`` in Angstrom."""
        if topology.box is None:
            raise ValueError('the topology has no periodic cell (CRYST1 record)')
        types, bonds = self._match(topology)
        data = _Data(topology, types, bonds)
        rc = _cutoff_angstrom(nonbondedCutoff)
        self._potentials = []
        for g in self._forces:
            g.createForce(data, np.asarray(topology.box, dtype=np.float64), rc)
            self._potentials.append(g.getJaxPotential())
        return list(self._potentials)
