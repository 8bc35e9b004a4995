/*
 * b200asm — C ABI of the B200-native global-assembly engine.
 *
 * This is the drop-in boundary underneath NeoPZ's parallel-strategy interface
 * (StrMatrix/TPZStrMatParInterface.h:44-48, the two Assemble() virtuals that
 * StrMatrix/pzstrmatrixor.cpp:41-101 implements on the CPU).  The C++ strategy
 * TPZStructMatrixB200<TVar> (neopz_b200/csrc/neopz/TPZStructMatrixB200.h) flattens a TPZCompMesh
 * once and drives these entry points; so does the Python host (neopz_b200/strmatrix.py) and the
 * benchmark.  Plain pointers and sizes only: no NeoPZ, no torch types.
 *
 * All functions return 0 on success, a negative B200ASM_E* code on failure; the message is
 * available from b200asm_last_error().  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with B200ASM_ENODEVICE.
 */
#ifndef B200ASM_H
#define B200ASM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200asm_ctx b200asm_ctx;

/* element topologies (reference: MElementType ECube/ETetraedro/EQuadrilateral/ETriangle/EOned/EPrisma/EPiramide, Common/pzeltype.h:52-62) */
enum { B200ASM_HEX = 0, B200ASM_TET = 1, B200ASM_QUAD = 2, B200ASM_TRI = 3, B200ASM_LINE = 4 /* EOned */,
       B200ASM_PRISM = 5 /* EPrisma: TPZShapePrism / TPZGeoPrism */, B200ASM_PYRAMID = 6 /* EPiramide: TPZShapePiram / TPZGeoPyramid */ };
/* weak forms:
 *  POISSON       Material/Poisson/TPZMatPoisson.cpp:19-42
 *  ELASTICITY3D  Material/Elasticity/TPZElasticity3D.cpp:85-107,269-372
 *  BC            the boundary forms of both (TPZMatPoisson.cpp:45-121, TPZElasticity3D.cpp:616-773)
 *                reduced to  ek(ns*i+a, ns*j+b) += M[a][b]*phi_i*phi_j*w ,  ef(ns*i+a) += v[a]*phi_i*w
 *  ELASTICITY2D  Material/Elasticity/TPZElasticity2D.cpp:86-203 on the quadrilaterals / triangles of a plane mesh
 *                (nstate 2; boundary = LINE elements of kind BC); POISSON on QUAD / TRI = TPZMatPoisson(dim 2) */
enum { B200ASM_POISSON = 0, B200ASM_ELASTICITY3D = 1, B200ASM_BC = 2, B200ASM_ELASTICITY2D = 3 };

enum {
    B200ASM_OK = 0,
    B200ASM_EINVAL = -1,     /* bad argument / unsupported element configuration */
    B200ASM_ENODEVICE = -2,  /* no CUDA device: the engine has no CPU path */
    B200ASM_ECUDA = -3,      /* a CUDA runtime call failed */
    B200ASM_ESTATE = -4,     /* call order violated (e.g. assemble before set_pattern) */
    B200ASM_EPATTERN = -5    /* an element entry has no slot in the CSR pattern (cf. pzsysmp.cpp:409) */
};

/* flags of b200asm_set_pattern */
#define B200ASM_SYMMETRIC 1 /* TPZSYsmpMatrix: upper triangle, Matrix/pzsysmp.h:108-127 */
#define B200ASM_FULL 0      /* TPZFYsmpMatrix: all entries, Matrix/pzysmp.h:229-237 */

/* scatter strategies (b200asm_set_option "scatter") */
#define B200ASM_SCATTER_ATOMIC 0 /* red.global.add.f64 */
#define B200ASM_SCATTER_COLORED 1 /* conflict-free element colouring, deterministic */

/*
 * A homogeneous batch of computational elements: same topology, same uniform order p, same
 * material.  Mirrors what CalcStiff sees per element (Mesh/pzinterpolationspace.cpp:404-473):
 * corner nodes, the integration rule, the shape tables and the material constants.
 */
typedef struct {
    int32_t topology; /* B200ASM_HEX ... */
    int32_t porder;   /* uniform order of every connect of the batch (hex / quad / line: 1..4; tet / tri / prism / pyramid: 1, 2) */
    int32_t kind;     /* B200ASM_POISSON ... */
    int32_t nstate;   /* TPZMaterial::NStateVariables(): 1 or 3 */
    int64_t nel;
    const int32_t *elnodes; /* [nel][ncorner] indices into the node table (TPZGeoEl::NodeIndex) */
    const int64_t *dest;    /* [nel][nshape*nstate] TPZElementMatrix::fDestinationIndex, Mesh/pzelmat.cpp:37-70; with an active
                               TPZEquationFilter the filtered (condensed) index, -1 for removed equations
                               (TPZEquationFilter::Filter, StrMatrix/TPZEquationFilter.h:120-141) */
    int32_t nqp;            /* points of the TPZIntPoints rule of order 2p (Mesh/pzelctemp.cpp:35-47) */
    int32_t nshape;         /* H1 shape functions per element.  Normally b200asm_nshape(topology, porder); a SMALLER count declares an
                               element whose sides carry different orders (TPZShapeH1<TSHAPE>::Initialize with per-side connect
                               orders, Shape/TPZShapeH1.cpp:14-37): porder is then the largest order, phi / dphi define the functions,
                               dest their equations, and the runtime-size kernels run the group */
    const double *qpts;     /* [nqp][dim]  TPZIntPoints::Point */
    const double *qwts;     /* [nqp] */
    const double *phi;      /* [nqp][nshape]       TPZShapeH1<TSHAPE>::Shape at the points (Shape/TPZShapeH1.cpp:42-116) */
    const double *dphi;     /* [nqp][dim][nshape]  master-element gradients */
    /* POISSON:      coef[0]=fScale, coef[1]=value of the (constant) forcing function
     * ELASTICITY3D: coef[0..2]=C1,C2,C3 (TPZElasticity3D.h:183-188), coef[3..5]=fForce, coef[6..8]=fPreStress
     * ELASTICITY2D: coef[0..2]=cA,cB,cC (plane strain: F(1-nu), F(1-2nu)/2, F nu with F=E/((1+nu)(1-2nu)); plane stress:
     *               E/(1-nu^2), E/(2(1+nu)), nu E/(1-nu^2)), coef[3..4]=body force, coef[5..7]=fPreStressXX, XY, YY
     * BC:           coef[0..8]=M (3x3 row-major, upper-left ns x ns used), coef[9..11]=v */
    double coef[16];
    const double *force; /* optional [nel][nqp][nstate]: forcing function evaluated by the host at the
                            integration points (std::function callbacks stay on the host); NULL = constant.
                            kind BC: boundary data given by a function (TPZBndCondT::ForcingFunctionBC): the coefficient of
                            phi_i * weight in ef at every point (e.g. BigNumber * g(x) for a Dirichlet condition,
                            TPZMatPoisson.cpp:62-90), replacing the constant coef[9..11] */
} b200asm_group;

/* ---- life cycle ------------------------------------------------------------------------- */
int b200asm_device_count(void); /* CUDA devices visible to this process (0: none - every compute entry point then fails) */
int b200asm_create(b200asm_ctx **out, int device);
void b200asm_destroy(b200asm_ctx *ctx);
const char *b200asm_last_error(const b200asm_ctx *ctx); /* ctx may be NULL: last create() error */
/* run on an existing CUDA stream (cudaStream_t passed as void*); default: a stream the context owns */
int b200asm_set_stream(b200asm_ctx *ctx, void *cuda_stream);
/* integer options: "scatter" (B200ASM_SCATTER_*), "engine" (0 register-tile DFMA kernels, 1 DMMA / closed-form kernels where one
 * exists, 2 the generic runtime-size kernel for every volume group - the kernel that otherwise only runs the groups without a
 * specialised one: other orders, sides of different order), "timing" (1: CUDA events around every group's kernel launches, read by b200asm_group_time_ms),
 * "affine" (default 1: hexahedral groups of order <= 2 whose elements are ALL parallelepipeds - measured on the device from
 * the node coordinates, to 1e-13 of the shortest edge vector, again after every b200asm_set_nodes - run the closed-form
 * kernel: constant Jacobian, no quadrature loop; 0: always the Gram / DMMA kernels),
 * "locality" (default 1, before add_group: volume groups are stored along a Morton curve through the element centroids,
 * inside every overlap chunk, when b200asm_set_nodes was called first: the rows of A an element shares with its neighbours
 * are then revisited while they are still in L2),
 * "overlap" (default 1: b200asm_assemble with a host matrix copies the finished rows of A back while later element
 * chunks are still being assembled; "overlap_min_elements" (before add_group) and "overlap_min_bytes" tune the chunking),
 * "variant" (before add_group: 0 default kernels; 7 DMMA kernels on tetrahedra p <= 2; hexahedra p = 2 Poisson: 20 sum factorisation
 * with one warp per element (= the default), 13 with one CTA per element, 16 the one-warp DMMA kernel; hexahedra p = 2
 * elasticity: 31 a pair of warps per element (= the default), 30 one warp, 34 the team of ten warps; 21 register-tile kernel on
 * tetrahedra p = 3, 4),
 * "gather" (default 0; 1: on one GPU in atomic mode the closed-form groups are assembled row by row by the warp that owns the
 * node - every CSR row written once, fixed summation order, no atomics on those rows; slower than the scatter kernels),
 * "drop_tiny" (default 0; 1: element entries below 1e-12 are skipped like TPZSYsmpMatrix::AddKel does, Matrix/pzsysmp.cpp:381),
 * "staging_lo" / "staging_hi" (before add_group, row-sharded assembly: see b200asm_exchange_*), "exchange_timeout_ms" */
int b200asm_set_option(b200asm_ctx *ctx, const char *name, int64_t value);

/* ---- flattened mesh ---------------------------------------------------------------------- */
/* node coordinates xyz[nnodes][3] (TPZGeoNode::Coord).  May be called again to move the nodes. */
int b200asm_set_nodes(b200asm_ctx *ctx, int64_t nnodes, const double *xyz);
/* copies the batch to the device; returns the group index (>=0) or an error code */
int b200asm_add_group(b200asm_ctx *ctx, const b200asm_group *g);
/* replaces coef[] of a group (material constants may change between assemblies) */
int b200asm_set_group_coef(b200asm_ctx *ctx, int group, const double coef[16]);
/* replaces the force table of a group ([nel][nqp][nstate] in the element order of add_group): forcing functions and boundary
 * data that change between assemblies (time-dependent sources; the reference calls the std::function again at every
 * CalcStiff, Material/Poisson/TPZMatPoisson.cpp:24-27).  The group must have been added with a table. */
int b200asm_set_group_force(b200asm_ctx *ctx, int group, const double *force);
int b200asm_clear_groups(b200asm_ctx *ctx);

/* ---- CSR pattern --------------------------------------------------------------------------
 * ia[neq+1], ja[nnz] exactly as TPZSYsmpMatrix::IA()/JA() (or TPZFYsmpMatrix) hold them after
 * TPZSSpStructMatrix::Create() (StrMatrix/TPZSSpStructMatrix.cpp:31-193).  Uploads the pattern and
 * builds, on the device, the element-entry -> CSR-position scatter map of every group. */
int b200asm_set_pattern(b200asm_ctx *ctx, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric);

/* The same pattern built ON THE DEVICE from the element->connect graph (the inputs of b200asm_build_pattern below):
 * replaces TPZSSpStructMatrix::Create / TPZSpStructMatrix::Create (StrMatrix/TPZSSpStructMatrix.cpp:31-193,
 * StrMatrix/TPZSpStructMatrix.cpp:53-190; External/TPZRenumbering.cpp:30-110) bit-exactly, without ever holding the
 * pattern in host memory, makes it the current pattern of the context and builds the scatter maps. */
int b200asm_build_pattern_device(b200asm_ctx *ctx, int symmetric, int64_t nel, const int64_t *elgraphindex,
                                 const int64_t *elgraph, int64_t nblock, const int64_t *blockpos, const int64_t *blocksize,
                                 int64_t *neq_out, int64_t *nnz_out);
/* copies the current pattern to the host as the reference stores it (int64 IA/JA); either pointer may be NULL */
int b200asm_get_pattern(b200asm_ctx *ctx, int64_t *ia_host, int64_t *ja_host);
/* column indices [first, first+count) of the resident pattern (the row-sharded setup only needs the interface rows) */
int b200asm_get_ja_range(b200asm_ctx *ctx, int64_t first, int64_t count, int64_t *ja_host);

/* ---- assembly -----------------------------------------------------------------------------
 * Zeroes A and rhs on the device, runs every group (CalcStiff + AddKel + AddFel of
 * StrMatrix/pzstrmatrixor.cpp:157-250 for all elements at once) and, when the host pointers are
 * non-NULL, copies the result back: a_host[nnz] in CSR order, rhs_host[neq].  Synchronous.  With atomic scatter the
 * download of A overlaps the kernels: large groups run in element chunks and every row that no later chunk can touch
 * (rows below the smallest destination equation of the remaining elements) is copied on a second stream at once
 * (only when a_host is page-locked: cudaHostAlloc / cudaHostRegister; pageable memory takes the plain path). */
int b200asm_assemble(b200asm_ctx *ctx, double *a_host, double *rhs_host);
/* Load vector only: TPZStrMatParInterface::Assemble(rhs) (StrMatrix/TPZStrMatParInterface.h:47-48; the reference's
 * CalcResidual path Mesh/pzinterpolationspace.cpp:476-527, which for these linear materials evaluates the same ef as
 * CalcStiff).  Zeroes rhs on the device, runs every group without the Gram products and without touching the CSR
 * values, copies rhs back when rhs_host != NULL. */
int b200asm_assemble_rhs(b200asm_ctx *ctx, double *rhs_host);
/* asynchronous, device resident (no copies, no synchronisation): enqueue on the context stream */
int b200asm_assemble_async(b200asm_ctx *ctx);
int b200asm_synchronize(b200asm_ctx *ctx);
/* copy the device-resident result to the host (synchronous) */
int b200asm_download(b200asm_ctx *ctx, double *a_host, double *rhs_host);
/* Page-lock / release a host buffer the caller owns (cudaHostRegister / cudaHostUnregister), e.g. the storage of a
 * TPZSYsmpMatrix (Matrix/pzsysmp.h:223-229): b200asm_assemble then downloads into it at PCIe speed and overlapped with the
 * kernels.  The buffer must be released before its owner frees it. */
int b200asm_pin_host(b200asm_ctx *ctx, void *ptr, size_t bytes);
int b200asm_unpin_host(b200asm_ctx *ctx, void *ptr);
/* device pointers of the resident CSR values / rhs (for a GPU solver downstream) */
int b200asm_device_pointers(b200asm_ctx *ctx, double **a_dev, double **rhs_dev);
/* ---- solve (device resident) ------------------------------------------------------------------
 * Conjugate gradients on the resident CSR values, statement by statement the reference's
 * CG(A, x, b, M, residual, max_iter, tol, FromCurrent) (Solvers/LinearSolvers/cg.h:44-120) with the product of
 * TPZSYsmpMatrix::MultAdd (Matrix/pzsysmp.cpp:190-232) / TPZFYsmpMatrix::MultAdd.
 * precond: 0 identity (TPZCopySolve), 1 one Jacobi sweep (TPZStepSolver::SetJacobi(1, 0., 0)): z = D^-1 r.
 * f_host: right-hand side (NULL: the assembled load vector on the device).  x_host: initial guess when from_current != 0,
 * receives the solution (NULL: leave it on the device, b200asm_cg_solution_device).  Stops when ||r||/||b|| <= tol
 * (the reference's criterion) or after max_iter iterations; reports both. */
int b200asm_cg_solve(b200asm_ctx *ctx, int precond, int64_t max_iter, double tol, int from_current, const double *f_host,
                     double *x_host, int64_t *iters_out, double *resid_out);
int b200asm_cg_solution_device(b200asm_ctx *ctx, double **x_dev);
/* Multi-GPU interface exchange (row-sharded assembly, neopz_b200/distributed.py): adds n values received from the
 * neighbouring rank into the resident CSR values (target 0) or rhs (target 1) at precomputed, distinct
 * positions: dst[positions[k]] += values[k].  positions/values are DEVICE pointers; asynchronous. */
int b200asm_scatter_add(b200asm_ctx *ctx, int target, const int32_t *positions_dev, const double *values_dev, int64_t n);
/* ---- multi-GPU, row-sharded assembly: interface exchange over peer memory -------------------------------
 * One context per GPU.  Every context holds a LOCAL system: the rows it owns plus "staging" rows that another GPU owns but its
 * elements contribute to (options "staging_lo" / "staging_hi", local row range, before add_group: the elements that touch
 * them are stored and launched first).  At every assembly the staged values are added straight into the owner's CSR values and
 * load vector over NVLink (P2P stores into the peer's memory, mapped with cudaDeviceEnablePeerAccess inside one process or
 * cudaIpcOpenMemHandle between the processes of a one-process-per-GPU job) on a side stream, while the interior elements are
 * assembled; GPUs order themselves with step counters in device memory, the host never synchronises.  The reference's only
 * parallel knob is the thread count of the strategy (StrMatrix/TPZStrMatParInterface.h:62-69); this is the GPU counterpart.
 * Every context of the job must run the same sequence of assemble calls.  A new pattern drops the links. */
#define B200ASM_MAX_PEERS 16
typedef struct { unsigned char handle[64]; int64_t offset; } b200asm_ipc_mem; /* cudaIpcMemHandle_t + offset in the allocation */
/* handles of this context's CSR values, load vector and flag block, for b200asm_exchange_add_peer in ANOTHER process */
int b200asm_exchange_export(b200asm_ctx *ctx, b200asm_ipc_mem out[3]);
/* Declares a peer.  push != 0: this context pushes staged contributions into it (then b200asm_exchange_set_map); push == 0: the
 * peer pushes into this context, incoming_min_row = smallest local row it touches (-1: unknown).  slot_there = index of the link
 * to THIS context in the peer's own table (the value its b200asm_exchange_add_peer returns / returned).  Give either the
 * exported handles of a context in another process (mem) or a context of this process (peer).  Returns the link index. */
int b200asm_exchange_add_peer(b200asm_ctx *ctx, int push, int slot_there, const b200asm_ipc_mem mem[3], b200asm_ctx *peer,
                              int64_t incoming_min_row);
/* what a push link sends: the CSR values [a_src0, a_src0 + n_a) of this context are added at a_dst[k] of the peer's values,
 * rhs[rhs_src[k]] at rhs_dst[k] of the peer's load vector (host arrays, copied) */
int b200asm_exchange_set_map(b200asm_ctx *ctx, int link, int64_t n_a, int64_t a_src0, const int32_t *a_dst, int64_t n_rhs,
                             const int32_t *rhs_src, const int32_t *rhs_dst);
int b200asm_exchange_clear(b200asm_ctx *ctx);
/* ---- multi-GPU in ONE process: the same calls as above on a set of devices ------------------------------------------------
 * The counterpart of TPZStrMatParInterface::SetNumThreads (StrMatrix/TPZStrMatParInterface.h:62-69) for GPUs.  The caller gives
 * the flattened mesh and the GLOBAL pattern exactly as to one context; b200asm_multi_set_pattern partitions the elements by
 * their smallest destination equation into chunks of equal work, GPU g owns the row block [row_begin[g],
 * row_begin[g+1]) of the global CSR, holds its slice of IA / JA / A only, and the contributions of its elements to rows of
 * other GPUs travel over NVLink through b200asm_exchange_* while the interior elements are assembled.  b200asm_multi_assemble
 * writes the caller's GLOBAL a_host[nnz] / rhs_host[neq]: every GPU its own slice, concurrently. */
typedef struct b200asm_multi b200asm_multi;
int b200asm_multi_create(b200asm_multi **out, int ndev, const int *devices /* NULL: 0 .. ndev-1 */);
void b200asm_multi_destroy(b200asm_multi *m);
const char *b200asm_multi_last_error(const b200asm_multi *m);
int b200asm_multi_num_devices(const b200asm_multi *m);
int b200asm_multi_context(b200asm_multi *m, int k, b200asm_ctx **ctx); /* the context of the k-th device (timing, counters, pointers) */
int b200asm_multi_set_option(b200asm_multi *m, const char *name, int64_t value);
int b200asm_multi_set_nodes(b200asm_multi *m, int64_t nnodes, const double *xyz);
int b200asm_multi_add_group(b200asm_multi *m, const b200asm_group *g); /* global destination indices; copied */
int b200asm_multi_set_group_coef(b200asm_multi *m, int group, const double coef[16]);
int b200asm_multi_set_group_force(b200asm_multi *m, int group, const double *force);
int b200asm_multi_clear_groups(b200asm_multi *m);
int b200asm_multi_set_pattern(b200asm_multi *m, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric);
/* the partition: row_begin[ndev + 1], elements[ndev], staged_entries[ndev] (CSR entries that travel per assembly); NULL = skip */
int b200asm_multi_partition(const b200asm_multi *m, int64_t *row_begin, int64_t *elements, int64_t *staged_entries);
int b200asm_multi_assemble(b200asm_multi *m, double *a_host, double *rhs_host);
int b200asm_multi_assemble_rhs(b200asm_multi *m, double *rhs_host);
int b200asm_multi_assemble_async(b200asm_multi *m);
int b200asm_multi_synchronize(b200asm_multi *m);
int b200asm_multi_counters(const b200asm_multi *m, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes);
/* b200asm_cg_solve on the row-sharded matrix of the last b200asm_multi_assemble: every GPU multiplies its row block, the halos of
 * the direction vector and of the product travel over NVLink (peer loads / reductions), the scalars of the iteration are summed
 * on the host in GPU order.  f_host / x_host are GLOBAL vectors (f_host == NULL: the assembled load vector); x_host == NULL leaves
 * the solution slices on the devices. */
int b200asm_multi_cg_solve(b200asm_multi *m, int precond, int64_t max_iter, double tol, int from_current, const double *f_host,
                           double *x_host, int64_t *iters_out, double *resid_out);
/* duration (ms, CUDA events on the context stream) of the kernel launches of one group in the LAST assembly;
 * needs option "timing" = 1.  Waits for that group's launches to finish. */
int b200asm_group_time_ms(b200asm_ctx *ctx, int group, double *ms);
/* kernel family that assembles the matrix part of a group (chosen when the scatter maps are built, i.e. valid after the first
 * assembly): "sumfact", "dmma", "closed_form", "gather_rows", "register_tile", "generic", "plane", "boundary".  Diagnostic: the
 * reference has no counterpart (its one code path is Mesh/pzinterpolationspace.cpp:404-473). */
int b200asm_group_kernel(const b200asm_ctx *ctx, int group, char *name, int len);
/* number of kernels launched by this context so far, and bytes moved H2D/D2H */
int b200asm_counters(const b200asm_ctx *ctx, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- host-side helpers (pure CPU, no device needed) ----------------------------------------
 * Gauss-Legendre rule the reference uses for order `order` (Integral/tpzgaussrule.cpp:171-243):
 * npts = (int)(0.51*(order+2)) points in the reference's interleaved -z,+z order. Returns npts. */
int b200asm_gauss_legendre(int order, double *loc, double *w);
/* tensor rules for hexahedra / quadrilaterals of order 2p in the reference's point order
 * (Integral/pzquad.cpp:153-169,268-284).  Returns the number of points. */
int b200asm_tensor_rule(int topology, int order, double *qpts, double *qw);
/* prism rule of the reference (TPZIntPrism3D, Integral/pzquad.cpp:408-436): the Gauss-Legendre line rule of `order` in zeta times
 * a triangle rule in (xi, eta) supplied by the caller (the reference's triangle tables are data, Integral/tpzintrulet.cpp);
 * triangle point fastest, weight = line weight * triangle weight.  Returns the number of points. */
int b200asm_prism_rule(int order, int ntri, const double *tripts, const double *triw, double *qpts, double *qw);
/* H1 shape tables (uniform p<=2) at given master-element points.  Returns nshape.  Prisms: Shape/pzshapeprism.cpp:42-205;
 * pyramids: Shape/pzshapepiram.cpp:47-119,331-392 (rational corner functions; points must stay off the apex). */
int b200asm_shape_tables(int topology, int porder, int nqp, const double *qpts, double *phi, double *dphi);
/* number of H1 shape functions of an element of uniform order p (TSHAPE::NShapeF): hex (p+1)^3, quad (p+1)^2, tetrahedra
 * (p+1)(p+2)(p+3)/6, triangles (p+1)(p+2)/2, lines p+1; prisms (6, 18) / pyramids (5, 14) for p <= 2. */
int b200asm_nshape(int topology, int porder);
/* Side-orientation key of every element from the GLOBAL indices of its corner nodes (what
 * ComputeTransforms / GetTransformId derive per side: Shape/pzgenericshape.cpp:57-68, Topology/tpzcube.cpp:1059-1111,
 * Topology/tpzquadrilateral.cpp:591-618, Topology/tpztetrahedron.cpp:1128-1170, Topology/tpztriangle.cpp:599-658): 1 bit per edge,
 * 3 bits per quadrilateral or triangular side (hexahedra: edges in bits 0-11, faces from bit 12; tetrahedra: edges 0-5, faces from
 * bit 6; quadrilaterals: edges 0-3, interior from bit 4; triangles: edges 0-2, interior from bit 3; lines: bit 0).  For p >= 3 the shape
 * functions of a side depend on it; elements with equal keys share their tables (one b200asm_group per key).
 * elnodes[nel][ncorner]. */
int b200asm_orientation_keys(int topology, int64_t nel, const int32_t *elnodes, int64_t *keys);
/* H1 shape tables of uniform order p (any p for hex / quad / line / tet / tri; p <= 2 for prisms / pyramids) for the orientation class `key`:
 * TPZShapeH1<TSHAPE>::Shape (Shape/TPZShapeH1.cpp:42-116) at the given points.  Returns nshape. */
int b200asm_shape_tables_oriented(int topology, int porder, int64_t key, int nqp, const double *qpts, double *phi,
                                  double *dphi);
/* CSR pattern of the reference from the element->connect graph (Mesh/pzcmesh.cpp:1223-1267,
 * External/TPZRenumbering.cpp:76-110, TPZSSpStructMatrix.cpp:50-193 / TPZSpStructMatrix.cpp:53-190).
 * elgraphindex[nel+1], elgraph[]: sequence numbers of each element's connects; blockpos/blocksize
 * per sequence number (TPZBlock::Position/Size).  Two calls: ja==NULL returns nnz and fills ia. */
int64_t b200asm_build_pattern(int symmetric, int64_t nel, const int64_t *elgraphindex, const int64_t *elgraph,
                              int64_t nblock, const int64_t *blockpos, const int64_t *blocksize, int64_t *ia,
                              int64_t *ja, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* B200ASM_H */
