"""Host-side mirror of the reference's struct-matrix interface for the B200 strategy.

Names and call order follow NeoPZ so that the parity tests read like the reference's own
(UnitTest_PZ/TestStruct/StructMatrixUnitTest.cpp): build materials, build the struct matrix on a
mesh, Create() the CSR pattern, Assemble() into it, or CreateAssemble() in one go
(StrMatrix/TPZStrMatParInterface.cpp:6-26).  The mesh is a gridmesh.FlatMesh (what the C++ strategy
csrc/neopz/TPZStructMatrixB200.cpp extracts from a TPZCompMesh); all arithmetic happens in the CUDA
library behind the C ABI (include/b200asm.h).  No CPU fallback.
"""
import os

import numpy as np

from . import capi
from .gridmesh import face_outward as gridmesh_face_outward
from .gridmesh import DIM, FlatMesh

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ---------------------------------------------------------------------------------------------
# materials (constants exactly as the reference derives them)
# ---------------------------------------------------------------------------------------------
class TPZMatPoisson:
    """Material/Poisson/TPZMatPoisson.h: -scale*Laplace(u) = f.  nstate 1."""
    nstate = 1

    def __init__(self, matid, dim=3):
        self.id, self.dim = matid, dim
        self.fScale = 1.0
        self.force = 0.0                     # no forcing function set -> 0 (TPZMatPoisson.cpp:24-27)
        # TPZMaterial::fBigNumber, Material/TPZMaterial.h:125: pow(10, max_digits10)*2/3
        self.fBigNumber = (10.0 ** 17) * 2 / 3

    def SetScaleFactor(self, s):
        self.fScale = float(s)

    def SetForcingFunction(self, value):
        """Forcing function: a constant (needs no table) or a callable f(x[npts][3]) -> [npts] that the host evaluates at the
        integration points of every element (TPZMatPoisson.cpp:24-27)."""
        if callable(value):
            self.forcing, self.force = value, 0.0
        else:
            self.forcing, self.force = None, float(value)

    forcing = None

    def CreateBC(self, matid, bctype, val1, val2):
        return TPZBndCond(self, matid, bctype, val1, val2)

    def coef(self):
        return [self.fScale, self.force]

    kind = capi.POISSON


class TPZElasticity3D:
    """Material/Elasticity/TPZElasticity3D.h: constants of SetC (:183-188).  nstate 3."""
    nstate = 3
    kind = capi.ELASTICITY3D

    def __init__(self, matid, E, poisson, force, prestress=(0.0, 0.0, 0.0)):
        self.id = matid
        self.fE, self.fPoisson = float(E), float(poisson)
        self.fForce = [float(x) for x in force]
        self.fPreStress = [float(x) for x in prestress]
        nu = self.fPoisson
        self.C1 = self.fE / (2. + 2. * nu)
        self.C2 = self.fE * nu / (-1. + nu + 2. * nu * nu)
        self.C3 = self.fE * (nu - 1.) / (-1. + nu + 2. * nu * nu)
        self.forcing = None

    def SetForcingFunction(self, fn):
        """Body force as a function of x: fn(x[npts][3]) -> [npts][3] (TPZElasticity3D.cpp:271-274), host-evaluated."""
        self.forcing = fn

    def CreateBC(self, matid, bctype, val1, val2):
        return TPZBndCond(self, matid, bctype, val1, val2)

    def coef(self):
        return [self.C1, self.C2, self.C3] + self.fForce + self.fPreStress


class TPZElasticity2D:
    """Material/Elasticity/TPZElasticity2D.h: plane strain (default of the setters used here) or plane stress.  nstate 2.
    The kernel takes the three constants of TPZElasticity2D.cpp:141-199: ek(2i+a,2j+b) = a==b ? cA D_aa + cB D_a'a'
    : cC D_ab + cB D_ba."""
    nstate = 2
    kind = capi.ELASTICITY2D

    def __init__(self, matid, E, poisson, fx, fy, planestress=False):
        self.id = matid
        self.fE, self.fPoisson = float(E), float(poisson)
        self.ff = [float(fx), float(fy)]
        self.fPlaneStress = bool(planestress)
        self.fPreStress = [0.0, 0.0, 0.0]   # XX, XY, YY
        self.fBigNumber = (10.0 ** 17) * 2 / 3
        self.forcing = None

    def SetForcingFunction(self, fn):
        """Body force as a function of x: fn(x[npts][3]) -> [npts][2] (TPZElasticity2D.cpp:120-127), host-evaluated."""
        self.forcing = fn

    def SetPlaneStress(self):
        self.fPlaneStress = True

    def SetPlaneStrain(self):
        self.fPlaneStress = False

    def SetPreStress(self, sxx, syy, sxy, szz=0.0):
        self.fPreStress = [float(sxx), float(sxy), float(syy)]

    def CreateBC(self, matid, bctype, val1, val2):
        return TPZBndCond(self, matid, bctype, val1, val2)

    def coef(self):
        E, nu = self.fE, self.fPoisson
        if self.fPlaneStress:
            cA, cB, cC = E / (1 - nu * nu), E / (2. * (1 + nu)), E / (1 - nu * nu) * nu
        else:
            F = E / ((1. + nu) * (1. - 2. * nu))
            cA, cB, cC = (1. - nu) * F, (1. - 2. * nu) / 2. * F, nu * F
        return [cA, cB, cC] + self.ff + self.fPreStress


class TPZBndCond:
    """TPZBndCondT: type 0 Dirichlet (penalty), 1 Neumann; TPZMatPoisson also 2 (the Robin branch as the reference computes it:
    penalty load vector + BigNumber * Val1(0,0) * dphix(0,i) dphix(0,j), TPZMatPoisson.cpp:104-118); Elasticity3D also 2 mixed,
    3 directional null Dirichlet, 4 stress field times the face normal (TPZElasticity3D.cpp:724-737), 5-8 directional
    Dirichlet (x, y, z, x and z)."""
    kind = capi.BC

    def __init__(self, material, matid, bctype, val1, val2):
        self.material, self.id, self.type = material, matid, int(bctype)
        self.nstate = material.nstate
        self.val1 = np.zeros((3, 3))
        v1 = np.atleast_2d(np.asarray(val1, dtype=np.float64))
        self.val1[: v1.shape[0], : v1.shape[1]] = v1
        self.val2 = np.zeros(3)
        v2 = np.atleast_1d(np.asarray(val2, dtype=np.float64))
        self.val2[: len(v2)] = v2
        self.forcing = None

    def SetForcingFunctionBC(self, fn):
        """TPZBndCondT::SetForcingFunctionBC: boundary data given by a function.  fn(x[npts][3]) -> val2[npts][nstate]; the host
        evaluates it at every integration point of every boundary element (the device gets a table, like the domain forcing
        functions).  Types with val2 in the load vector only: 0, 1 (not Elasticity3D, whose Neumann data then comes from the
        face normal), 2 (Elasticity3D: val1 * fn), 5-8."""
        self.forcing = fn

    def HasForcingFunctionBC(self):
        return self.forcing is not None

    def rhs_coefficient(self, v2):
        """v2[..., nstate] (values of the forcing function) -> coefficient of phi_i * weight in ef, per state variable:
        TPZMatPoisson.cpp:79-100, TPZElasticity3D.cpp:637-697,739-772, TPZElasticity2D.cpp:256-283."""
        ns = self.nstate
        v2 = np.asarray(v2, dtype=np.float64)[..., :ns]
        if isinstance(self.material, TPZMatPoisson):
            if self.type == 0:
                return self.material.fBigNumber * v2
            if self.type == 1:
                return v2 * self.material.fScale
            if self.type == 2:
                return self.material.fBigNumber * v2
        elif isinstance(self.material, TPZElasticity2D):
            if self.type == 0:
                return self.material.fBigNumber * v2
            if self.type == 1:
                return v2
        else:
            big = 1.e12
            if self.type == 0:
                return big * v2
            if self.type == 2:   # val2loc[i] = sum_j val1(i,j) * fn_j
                out = np.zeros_like(v2)
                for i in range(3):
                    for j in range(3):
                        out[..., i] += self.val1[i, j] * v2[..., j]
                return out
            if self.type in (5, 6, 7, 8):
                on = {5: (1, 0, 0), 6: (0, 1, 0), 7: (0, 0, 1), 8: (1, 0, 1)}[self.type]
                return big * v2 * np.array(on, dtype=np.float64)
        raise ValueError("boundary condition type %d with a forcing function is not supported" % self.type)

    def coef(self):
        """(M, v) of  ek += M[a][b] phi_i phi_j w ,  ef += v[a] phi_i w  (include/b200asm.h)."""
        ns = self.nstate
        M = np.zeros((3, 3))
        v = np.zeros(3)
        if isinstance(self.material, TPZElasticity2D):
            big = self.material.fBigNumber   # TPZElasticity2D.cpp:219
            if self.type == 0:      # :256-270
                M[:ns, :ns] = np.eye(ns) * big
                v[:ns] = big * self.val2[:ns]
            elif self.type == 1:    # :273-283
                v[:ns] = self.val2[:ns]
            elif self.type == 3:    # :305-316 directional null Dirichlet
                M[:ns, :ns] = np.diag(big * self.val2[:ns])
            else:   # (type 2 crashes in the reference itself: sliced TPZMatLoadCasesBC copy, :220)
                raise ValueError("TPZElasticity2D: boundary condition type %d not supported" % self.type)
        elif isinstance(self.material, TPZMatPoisson):
            big = self.material.fBigNumber
            if self.type == 0:      # TPZMatPoisson.cpp:79-90
                M[0, 0] = big
                v[0] = big * self.val2[0]
            elif self.type == 1:    # :93-100
                v[0] = self.val2[0] * self.material.fScale
            elif self.type == 2:    # :104-118: coef[12] = coefficient of dphix(0,i) dphix(0,j) w
                v[0] = big * self.val2[0]
                return M.reshape(-1).tolist() + v.tolist() + [big * self.val1[0, 0]]
            else:
                raise ValueError("TPZMatPoisson: boundary condition type %d not supported" % self.type)
        else:
            big = 1.e12             # TPZElasticity3D.cpp:630
            if self.type == 0:      # :663-675
                M[:ns, :ns] = np.eye(ns) * big
                v[:ns] = big * self.val2[:ns]
            elif self.type == 1:    # :677-683
                v[:ns] = self.val2[:ns]
            elif self.type == 2:    # :684-697
                M[:, :] = self.val1
                v[:ns] = self.val2[:ns]
            elif self.type == 3:    # :715-723 directional null Dirichlet: penalty scaled by val2 per direction
                M[:ns, :ns] = np.diag(big * self.val2[:ns])
            elif self.type == 4:    # :724-737: load vector only, -(val1 . normal) per integration point (table, see face_normals)
                pass
            elif self.type in (5, 6, 7, 8):   # :739-772 directional Dirichlet on x / y / z / x and z
                for k in {5: (0,), 6: (1,), 7: (2,), 8: (0, 2)}[self.type]:
                    M[k, k] = big
                    v[k] = big * self.val2[k]
            else:
                raise ValueError("TPZElasticity3D: boundary condition type %d not supported" % self.type)
        return M.reshape(-1).tolist() + v.tolist()


# ---------------------------------------------------------------------------------------------
# integration rules + shape tables per topology / order
# ---------------------------------------------------------------------------------------------
def element_tables(topology, porder, key=0):
    """(qpts, qw, phi, dphi) of the rule of order 2p the reference attaches to an element
    (Mesh/pzelctemp.cpp:35-47, Material/TPZMatSingleSpace.cpp:61-72); key = side-orientation class of the
    elements (capi.orientation_keys), relevant for p >= 3."""
    order = 2 * porder
    if topology in (capi.HEX, capi.QUAD, capi.LINE):
        qpts, qw = capi.tensor_rule(topology, order)
    else:
        z = np.load(os.path.join(_DATA, "simplex_rules.npz"))
        if topology == capi.PRISM:  # TPZIntPrism3D: line rule x triangle rule (Integral/pzquad.cpp:408-436)
            qpts, qw = capi.prism_rule(order, z[f"tri_order{order}_pts"], z[f"tri_order{order}_w"])
        else:
            tag = {capi.TET: "tet", capi.TRI: "tri", capi.PYRAMID: "pyr"}[topology]
            qpts, qw = z[f"{tag}_order{order}_pts"], z[f"{tag}_order{order}_w"]
    phi, dphi = capi.shape_tables(topology, porder, qpts, key)
    return qpts, qw, phi, dphi


def points_x(topology, qpts, coords):
    """Physical coordinates of the integration points: x = sum_a N_a(qsi) x_a with the (multi)linear corner functions
    (TPZGeoEl::X, what ComputeRequiredData stores in data.x).  coords: [nel][ncorner][3] -> [nel][nq][3]."""
    geo, _ = capi.shape_tables(topology, 1, qpts)           # [nq][ncorner]
    x = np.zeros((coords.shape[0], geo.shape[0], 3))
    for a in range(geo.shape[1]):
        x += geo[None, :, a, None] * coords[:, None, a, :]
    return x


def face_normals(topology, qpts, coords, outward):
    """data.normal of boundary faces at the integration points (TPZInterpolationSpace::ComputeNormal,
    Mesh/pzinterpolationspace.cpp:300-383): the unit vector axes(0) x axes(1) of the Gram-Schmidt Jacobian, i.e. the
    normalised dx/dxi x dx/deta, turned towards `outward` (gridmesh.face_outward).  coords [nel][ncorner][3] -> [nel][nq][3]."""
    _, dgeo = capi.shape_tables(topology, 1, qpts)          # [nq][2][ncorner]
    v1 = np.einsum("qa,eak->eqk", dgeo[:, 0, :], coords)
    v2 = np.einsum("qa,eak->eqk", dgeo[:, 1, :], coords)
    n = np.cross(v1, v2)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    flip = np.einsum("eqk,ek->eq", n, outward) < 0.0
    n[flip] *= -1.0
    return n


# ---------------------------------------------------------------------------------------------
# the struct matrix
# ---------------------------------------------------------------------------------------------
class TPZStructMatrixB200:
    """TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>> (symmetric=True) or
    TPZSpStructMatrix<...> (symmetric=False) on a flattened mesh."""

    def __init__(self, mesh: FlatMesh, materials, symmetric=True, device=0, nthreads=0, engine=None, scatter=None, variant=None,
                 devices=None):
        """devices: list of CUDA devices of THIS process that share the assembly (b200asm_multi: element partition by the
        smallest destination equation, row-sharded CSR, interface rows over NVLink) - what SetNumThreads(n) means for the C++
        strategy; None / one entry: a single context on `device`."""
        self.mesh = mesh
        self.materials = {m.id: m for m in (materials.values() if isinstance(materials, dict) else materials)}
        self.symmetric = bool(symmetric)
        self.fNumThreads = nthreads
        self.multi = devices is not None and len(devices) > 1
        self.ctx = capi.MultiContext(devices) if self.multi else capi.Context(devices[0] if devices else device)
        if engine is not None:  # 0: register-tile DFMA kernels only, 1 (default): DMMA panel kernels where available
            self.ctx.set_option("engine", engine)
        if variant:  # tuning alternative of the DMMA kernels (0 = default)
            self.ctx.set_option("variant", variant)
        if scatter is not None:  # "atomic" (default) or "colored" (conflict-free element colouring, deterministic)
            self.ctx.set_option("scatter", {"atomic": 0, "colored": 1}[scatter])
        self.ia = self.ja = None
        self._flattened = False
        self.group_of_block = []    # first (for p <= 2: the only) group of every element block
        self.groups_of_block = []   # all groups of every element block (one per side-orientation class)

    def SetNumThreads(self, n):
        self.fNumThreads = n

    # -- flatten: one b200asm_group per element block --------------------------------------------
    def _flatten(self):
        if self._flattened:
            return
        mesh = self.mesh
        self.ctx.set_nodes(mesh.nodes)
        for b in mesh.blocks:
            mat = self.materials.get(b.matid)
            if mat is None:
                raise KeyError(f"no material with id {b.matid}")
            if mat.nstate != mesh.nstate:
                raise ValueError("material nstate does not match the mesh")
            meshdim = max(DIM[x.topology] for x in mesh.blocks)
            if (DIM[b.topology] == meshdim) == (mat.kind == capi.BC):
                raise ValueError(f"material {b.matid}: domain/boundary kind does not match element dimension")
            # p >= 3: the shape functions of a side depend on the orientation of the side (global corner-node
            # indices): one group per orientation class of the block; p <= 2: one group per block
            keys = capi.orientation_keys(b.topology, b.elnodes) if mesh.porder >= 3 else np.zeros(len(b.elnodes), np.int64)
            gids = []
            for key in np.unique(keys):
                sel = slice(None) if mesh.porder < 3 else np.nonzero(keys == key)[0]
                qpts, qw, phi, dphi = element_tables(b.topology, mesh.porder, key)
                force = None
                if mat.kind == capi.BC and mat.type == 4 and isinstance(mat.material, TPZElasticity3D):
                    # stress-field Neumann: ef += -(val1 . normal) phi w, the normal of every integration point tabulated here
                    coords = mesh.nodes[b.elnodes[sel]]
                    nrm = face_normals(b.topology, qpts, coords, gridmesh_face_outward(mesh, b)[sel])
                    force = -np.einsum("ab,eqb->eqa", mat.val1, nrm)
                elif mat.kind == capi.BC and mat.HasForcingFunctionBC():
                    force = mat.rhs_coefficient(mat.forcing(points_x(b.topology, qpts, mesh.nodes[b.elnodes[sel]]).reshape(-1, 3)))
                    force = force.reshape(len(b.elnodes[sel]), len(qw), mat.nstate)
                elif mat.kind != capi.BC and getattr(mat, "forcing", None) is not None:
                    force = np.asarray(mat.forcing(points_x(b.topology, qpts, mesh.nodes[b.elnodes[sel]]).reshape(-1, 3)), dtype=np.float64)
                    force = force.reshape(len(b.elnodes[sel]), len(qw), -1)[..., : mat.nstate]
                gids.append(self.ctx.add_group(b.topology, mesh.porder, mat.kind, mat.nstate, b.elnodes[sel], b.dest[sel],
                                               qpts, qw, phi, dphi, mat.coef(), force=force))
            self.groups_of_block.append(gids)
            self.group_of_block.append(gids[0])
        self._flattened = True

    # -- TPZStructMatrix::Create -------------------------------------------------------------------
    def Create(self, on_device=False, download=True):
        """CSR pattern, bit-exact with the reference's Create() (TPZSSpStructMatrix.cpp:31-193).
        on_device: build it on the GPU (b200asm_build_pattern_device) instead of the threaded host builder;
        download=False then leaves IA/JA on the device only (self.ja stays None; self.nnz is set)."""
        idx, graph = self.mesh.element_graph()
        if on_device and self.multi:
            raise ValueError("Create(on_device=True) builds the pattern on ONE GPU; a multi-device struct matrix takes the host builder")
        if on_device:
            self._flatten()
            neq, self.nnz = self.ctx.build_pattern_device(self.symmetric, idx, graph, self.mesh.block_pos, self.mesh.block_size)
            assert neq == self.mesh.neq
            self.ia, self.ja = self.ctx.get_pattern(neq, self.nnz, want_ja=download)
            return self.ia, self.ja
        self.ia, self.ja = capi.build_pattern(self.symmetric, idx, graph, self.mesh.block_pos,
                                              self.mesh.block_size, self.fNumThreads)
        self.nnz = len(self.ja)
        self._flatten()
        self.ctx.set_pattern(self.ia, self.ja, self.symmetric)
        return self.ia, self.ja

    def SetPattern(self, ia, ja):
        """Use a pattern created elsewhere (e.g. by the reference's own Create())."""
        self.ia = np.ascontiguousarray(ia, dtype=np.int64)
        self.ja = np.ascontiguousarray(ja, dtype=np.int64)
        self.nnz = len(self.ja)
        self._flatten()
        self.ctx.set_pattern(self.ia, self.ja, self.symmetric)

    # -- TPZStrMatParInterface::Assemble(stiffness, rhs) ---------------------------------------------
    def Assemble(self, a=None, rhs=None):
        """Zero + assemble into the created pattern; returns (a, rhs) host arrays."""
        if self.ia is None:
            raise RuntimeError("Assemble: call Create() first")
        if a is None:
            a = np.empty(self.nnz)
        if rhs is None:
            rhs = np.empty(self.mesh.neq)
        self.ctx.assemble(a, rhs)
        return a, rhs

    def AssembleRhs(self, rhs=None):
        """TPZStrMatParInterface::Assemble(rhs): the load vector only (the CSR values on the device stay untouched)."""
        if self.ia is None:
            raise RuntimeError("AssembleRhs: call Create() first")
        if rhs is None:
            rhs = np.empty(self.mesh.neq)
        self.ctx.assemble_rhs(rhs)
        return rhs

    def SolveCG(self, max_iter=50000, tol=1e-15, jacobi=True, x0=None, f=None):
        """TPZStepSolver::SetCG(max_iter, Jacobi(1) | TPZCopySolve, tol, FromCurrent) + Solve on the device-resident system.
        Returns (solution, iterations, relative residual)."""
        return self.ctx.cg_solve(1 if jacobi else 0, max_iter, tol, x0=x0, f=f)

    def CreateAssemble(self):
        ia, ja = self.Create()
        a, rhs = self.Assemble()
        return ia, ja, a, rhs

    def UpdateMaterials(self):
        for b, gids in zip(self.mesh.blocks, self.groups_of_block):
            for gid in gids:
                self.ctx.set_group_coef(gid, self.materials[b.matid].coef())
