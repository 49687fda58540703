/*
 * optcuts_b200.h — C-ABI of liboptcuts_b200.so: the B200 (sm_100a, fp64) implementation of
 * OptCuts' geometry-and-topology inner loop.
 *
 * The reference has no FFI; its "plugin surface" is two C++ abstract classes plus the concrete
 * Optimizer.  The entry points below are what thin C++ subclasses
 *   CudaSymDirichletEnergy : OptCuts::Energy            (src/Energy/Energy.hpp:16-46)
 *   CudaLinSysSolver       : OptCuts::LinSysSolver<..>  (src/LinSysSolver/LinSysSolver.hpp:22-256)
 * and a device-resident mirror of OptCuts::Optimizer (src/Optimizer.hpp:22-143) bind to; the
 * stubs are shown in INTEGRATION.md.  Each declaration cites the reference interface it replaces.
 *
 * Conventions
 *  - plain C, no exceptions; every call returns 0 on success or a negative ocb_status; the text of
 *    the last error is kept per context (ocb_last_error).
 *  - all pointers are HOST pointers owned by the caller unless the name ends in `_dev`; data is
 *    copied synchronously (the call returns after the context's stream has drained).
 *  - matrices are column-major exactly as Eigen hands them over (MatrixXd V: all u then all v;
 *    MatrixXi F: |F|x3 col-major); gradient / search direction / solver vectors are interleaved
 *    [u0 v0 u1 v1 ...] (Optimizer.cpp:666-668, SymDirichletEnergy.cpp:287).
 *  - indices are int32, scalars IEEE fp64.  The one dense contraction on the path (the exact inverse of the coarse
 *    Galerkin matrix of the preconditioner) runs on the FP64 tensor cores (mma.sync m8n8k4); everything else is CUDA-core
 *    streaming work.
 *  - one CUDA stream per context; distinct contexts may be used from distinct threads.
 *  - ocb_create does no CUDA work beyond selecting the device lazily on first upload, so the C++
 *    shim objects are cheap to construct (nested optimizers create thousands: SURVEY H7).
 *
 * System layout ("whole mesh", Scaffold.cpp:179-184): global vertex ids [0,nV) are the mesh's UV
 * vertices; air-mesh vertex i maps to localVI2Global[i] (< nV for the first nBnd, which alias mesh
 * boundary vertices; nV + (i - nBnd) for the rest).  nSys = 2 * (nV + nVa - nBnd) DOFs.
 */
#ifndef OPTCUTS_B200_H
#define OPTCUTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ocb_ctx ocb_ctx;

typedef enum {
    OCB_OK = 0,
    OCB_ERR_CUDA = -1,        /* a CUDA runtime call failed (text in ocb_last_error) */
    OCB_ERR_ARG = -2,         /* bad argument / call order */
    OCB_ERR_STATE = -3,       /* required data (mesh, uv, pattern, matrix ...) not set */
    OCB_ERR_INVERTED = -4,    /* element with negative signed UV area (TriMesh::checkInversion, Optimizer.cpp:65-67) */
    OCB_ERR_NOT_CONVERGED = -5, /* PCG hit max_it (solution is still written) */
    OCB_ERR_BREAKDOWN = -6    /* PCG breakdown: pAp <= 0 (matrix not SPD) */
} ocb_status;

/* ---- context ------------------------------------------------------------------------------ */
int  ocb_create(ocb_ctx** out, int device);
void ocb_destroy(ocb_ctx* ctx);
const char* ocb_last_error(const ocb_ctx* ctx);
const char* ocb_version(void);
/* run everything of this context on an existing cudaStream_t (e.g. torch's current stream).  The handle is used as
 * given: NULL means the legacy default stream, so the library's work stays ordered with the caller's work on it.  Without
 * this call (or after ocb_use_own_stream) the context creates a private non-blocking stream on first use. */
int  ocb_set_stream(ocb_ctx* ctx, void* cuda_stream);
int  ocb_use_own_stream(ocb_ctx* ctx);
int  ocb_synchronize(ocb_ctx* ctx);
/* tuning switches: "pcg_scaled_norm" (0 / 1): 0 (default) = ocb_solve stops on ||r||_2 <= rel_tol ||b||_2; 1 = on the
 * block-Jacobi-scaled norm sqrt(r^T D^-1 r) (D = the 2x2 diagonal blocks), in which soft rows (the scaffold's) converge
 * relative to their own stiffness.  Measured against an LU of the same matrices (profiles/r2_pcg_norm.txt): no gain in
 * the accuracy of the direction, ~3 % more iterations -- the differences to the reference's LDL^T on badly conditioned
 * states (kappa ~ 1e12 at a Tutte start) are rounding, in both solvers, not the stopping test.
 * "mas_equilibrate" (default 1): the preconditioner's group / coarse inversions work on the diagonally equilibrated blocks (keeps the
 * preconditioner positive definite on badly scaled systems: one degenerate triangle, diagonal 3e-3 .. 1e10).
 * "scale_system" (default 0): symmetric diagonal scaling of the whole Newton system (measured: 7-17x more CG iterations).
 * "force_direct" (default 0, tests): ocb_solve goes to the direct safety net (block-tridiagonal Cholesky, see ocb_solve) instead of CG. */
int  ocb_set_option(ocb_ctx* ctx, const char* key, double value);
/* CUDA-event stopwatch on the context's stream (what bench.py times kernels with) */
int  ocb_timer_start(ocb_ctx* ctx);
int  ocb_timer_stop_ms(ocb_ctx* ctx, double* ms);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t ocb_launch_count(const ocb_ctx* ctx);

/* per-kernel-class CUDA-event profile on the context's stream (bench.py's roofline object): when enabled,
 * every launch is bracketed by events; ocb_profile_get returns accumulated ms and launch counts per class
 * (ocb_profile_classes() entries, names from ocb_profile_name). */
int ocb_profile_enable(ocb_ctx* ctx, int on);
int ocb_profile_get(ocb_ctx* ctx, double* ms, int64_t* counts);
const char* ocb_profile_name(int kernel_class);
int ocb_profile_classes(void);

/* ---- a1: rest-frame features — TriMesh::computeFeatures, TriMesh.cpp:343-398 ---------------
 * rest8_soa (8 x nF, row k contiguous): triArea, triAreaSq, e0SqLen, e1SqLen, e0dote1,
 * e0SqLen_div_dbAreaSq, e1SqLen_div_dbAreaSq, e0dote1_div_dbAreaSq.  Triangles with area below
 * areaThres_AM get the equilateral surrogate (TriMesh.cpp:373-383).  scalars3 = surfaceArea,
 * avgEdgeLen (igl::avg_edge_length), virtualRadius = sqrt(surfaceArea/pi).  Returns
 * OCB_ERR_ARG if a triangle has zero area (the reference exits: TriMesh.cpp:368-402). */
int ocb_rest_features(ocb_ctx* ctx, int nV, int nF, const double* V_rest_colmajor /* nV x 3 */,
                      const int32_t* F_colmajor, double areaThres_AM,
                      double* rest8_soa, double* scalars3);

/* ---- problem data --------------------------------------------------------------------------
 * ocb_set_mesh: TriMesh::F + features + fixedVert (TriMesh.hpp:30-68).  Call when topology changes. */
int ocb_set_mesh(ocb_ctx* ctx, int nV, int nF, const int32_t* F_colmajor, const double* rest8_soa,
                 double surfaceArea, const int32_t* fixed, int nFixed);
/* ocb_set_air: Scaffold::airMesh + localVI2Global + airMesh.fixedVert (Scaffold.cpp:169-199);
 * w_scaf_over_Fa = w_scaf / |F_air| (Optimizer.cpp:776,795,839).  nFa = 0 removes the scaffold.
 * Call after every Scaffold rebuild (Optimizer.cpp:236-239). */
int ocb_set_air(ocb_ctx* ctx, int nVa, int nFa, const int32_t* Fa_colmajor, const double* rest8_soa,
                const int32_t* localVI2Global, int nBnd, const int32_t* fixedAir, int nFixedAir,
                double w_scaf_over_Fa);
/* UV coordinates: TriMesh::V (nV x 2) and airMesh.V (nVa x 2, NULL when no scaffold) */
int ocb_set_uv(ocb_ctx* ctx, const double* V_colmajor, const double* Va_colmajor);
int ocb_get_uv(ocb_ctx* ctx, double* V_colmajor, double* Va_colmajor);
/* device-side snapshot / rollback of the UV state (what the reference does with whole-TriMesh copies:
 * data_findExtrema = result, Optimizer.cpp:381,452; triSoup_bestFeasible + setConfig, main.cpp:711-723) */
int ocb_save_uv(ocb_ctx* ctx);
int ocb_restore_uv(ocb_ctx* ctx);
/* sizes: nV, nF, nVa, nFa, nBnd, nSys, nnz_upper (reference CSR), nnz_blocks (BSR 2x2) */
int ocb_get_sizes(const ocb_ctx* ctx, int64_t* sizes8);

/* ---- a2/a13: energy — SymDirichletEnergy::getEnergyValPerElem (SymDirichletEnergy.cpp:24-46),
 * Energy::computeEnergyVal (Energy.cpp:35-40), Optimizer::computeEnergyVal (Optimizer.cpp:764-782).
 * E_total = energyParam0 * E_sd + E_scaf,  E_scaf = w_scaf/|Fa| * sum_air E_t(uniform).
 * Returns OCB_ERR_INVERTED (values still written, possibly inf/nan) if any signed area < 0. */
int ocb_energy(ocb_ctx* ctx, double energyParam0, double* E_total, double* E_sd, double* E_scaf);
/* per-element values of the mesh term, w = triArea/surfaceArea or 1 (uniformWeight) */
int ocb_energy_per_elem(ocb_ctx* ctx, int uniformWeight, double* out_nF);
/* a3: SymDirichletEnergy::getEnergyValByElemID (:48-68): the value of triangle triI alone */
int ocb_energy_by_elem(ocb_ctx* ctx, int triI, int uniformWeight, double* E);

/* ---- a4/a13: gradient — SymDirichletEnergy::computeGradient (:258-304), Optimizer::computeGradient
 * (Optimizer.cpp:783-797), Scaffold::augmentGradient (Scaffold.cpp:210-229).  Result (nSys,
 * interleaved, fixed vertices zeroed) stays on the device as the solver's right-hand side; g_out
 * may be NULL.  sqnorm = ||g||^2 (Optimizer.cpp:210). */
int ocb_gradient(ocb_ctx* ctx, double energyParam0, double* g_out, double* sqnorm);

/* ---- a10: pattern — LinSysSolver::set_pattern (LinSysSolver.hpp:37-135) on the merged adjacency
 * (Scaffold::mergeVNeighbor / mergeFixedV, Scaffold.cpp:295-313).  adj is the vNeighbor sets in CSR
 * form (each row ascending, as std::set iterates).  Builds the device BSR(2x2) pattern, the
 * element->block scatter map and the reference's upper-triangular scalar CSR layout. */
int ocb_set_pattern(ocb_ctx* ctx, int nVtot, const int32_t* adj_ptr, const int32_t* adj_idx,
                    const int32_t* fixed, int nFixed);
/* convenience: derive the adjacency from the element lists already uploaded (mesh + air) */
int ocb_set_pattern_from_elements(ocb_ctx* ctx);

/* ---- a5/a9/a11/a18: Hessian — SymDirichletEnergy::computeHessian (:429-549) + IglUtils::makePD
 * (IglUtils.hpp:71-90) + addBlockToMatrix/addDiagonalToMatrix (IglUtils.cpp:361-451) +
 * Optimizer::computeHessian (Optimizer.cpp:798-843) + Scaffold::augmentProxyMatrix
 * (Scaffold.cpp:231-248) + LinSysSolver::update_a (LinSysSolver.hpp:138-159): element blocks are
 * projected and scattered straight into the device matrix; no triplets exist. */
int ocb_hessian_assemble(ocb_ctx* ctx, double energyParam0);
/* compatibility path = Energy::computeHessian(data,&V,&I,&J,uniformWeight) of the MESH term
 * (unscaled by energyParam0): same triplet order as the reference.  Call with V==NULL to get *n. */
int ocb_hessian_triplets(ocb_ctx* ctx, int uniformWeight, double* V, int32_t* I, int32_t* J, int64_t* n);
/* projected 6x6 element blocks of the mesh term (row-major 36 per triangle, w applied, unscaled) */
int ocb_hessian_blocks(ocb_ctx* ctx, int uniformWeight, double* out_nFx36);
/* a6: Energy::computeHessian(data, MatrixXd&, uniformWeight) of the MESH term, the dense flavour of the nested optimizers
 * (SymDirichletEnergy.cpp:306-427): H_out is 2nV x 2nV (symmetric; fixed vertices: zero row / column, unit diagonal),
 * element blocks added in triangle order.  For local stencils: nV <= 4096. */
int ocb_hessian_dense(ocb_ctx* ctx, int uniformWeight, double* H_out);
/* mirrors LinSysSolver::update_a(I,J,S): zero, then accumulate triplets with i<=j */
int ocb_update_values_triplets(ocb_ctx* ctx, int64_t nT, const int32_t* I, const int32_t* J, const double* S);
/* reference layout read-back: 1-based upper-triangular CSR (LinSysSolver.hpp:27-29, get_ia/ja/a) */
int ocb_download_csr(ocb_ctx* ctx, int32_t* ia, int32_t* ja, double* a);
/* change counter of the device matrix (pattern or values): lets a caller keep its read-back copy (LinSysSolver::coeffMtr,
 * LinSysSolver.hpp:183-196, asks entry by entry) until the next assembly; -1 for a null context */
long long ocb_matrix_version(const ocb_ctx* ctx);
/* y = A x with the device matrix (fixes the reference's broken LinSysSolver::multiply, :172-187) */
int ocb_multiply(ocb_ctx* ctx, const double* x, double* y);

/* ---- a12: solve — EigenLibSolver::analyze_pattern/factorize/solve (EigenLibSolver.cpp:71-107),
 * replaced by a preconditioned CG in one persistent cooperative kernel (two-level additive Schwarz: 8-vertex leaves along
 * a Hilbert curve through the UVs, affine coarse spaces, exact dense coarse inverse; 2x2 block-Jacobi when no geometry is
 * known).  rhs==NULL solves A x = -gradient
 * (Optimizer.cpp:557-563) with the gradient left on the device by ocb_gradient; the solution stays
 * on the device as the search direction; x_out may be NULL.  rel_tol <= 0 -> 1e-12 (relative residual
 * ||r|| / ||b||; see ocb_set_option for the scaled variant), max_it <= 0 -> 20*n.
 * Safety net: when CG cannot handle a system of <= 40 000 unknowns (the two-level preconditioner comes out indefinite; inside
 * ocb_newton_step also a breakdown d.Ad <= 0 or the iteration cap) the system is solved directly, by a block-tridiagonal Cholesky
 * over the breadth-first levels of the matrix graph whose block steps are cuSOLVER / cuBLAS calls (opened with dlopen at first use;
 * without the libraries the solve is repeated with block-Jacobi CG).  ocb_precond_info()[14] counts those solves, [15] the rejected
 * preconditioners.  The libraries' host part uses OpenMP: a process started with OMP_NUM_THREADS=1 pays ~7x for such a solve. */
int ocb_factorize(ocb_ctx* ctx);   /* builds the preconditioner (Galerkin products, group inverses, coarse inverse); OCB_ERR_BREAKDOWN if a diagonal 2x2 block is not SPD */
int ocb_solve(ocb_ctx* ctx, const double* rhs, double* x_out, double rel_tol, int max_it,
              int* iters, double* rel_res);
/* host only, no CUDA call (tests): the breadth-first level order and the blocks the direct safety net's block-tridiagonal Cholesky
 * would use for a SYMMETRIC vertex pattern in CSR form (unsorted rows and diagonal entries allowed).  pos[v] = position of v in the
 * level order, blkOf[v] = its block, blkBeg[k] = first position of block k (up to n + 1 entries).  Every coupling joins equal or
 * adjacent blocks.  Returns the number of blocks (> 0) or an OCB_ERR_* code. */
int ocb_direct_level_blocks(int n, const int32_t* rowPtr, const int32_t* colIdx, int target_vertices_per_block,
                            int32_t* pos, int32_t* blkOf, int32_t* blkBeg);
int ocb_get_search_dir(ocb_ctx* ctx, double* p_out);
int ocb_set_search_dir(ocb_ctx* ctx, const double* p);

/* ---- a7: step bound — SymDirichletEnergy::initStepSize (:551-610) over mesh AND air mesh
 * (Optimizer::initStepSize, Optimizer.cpp:692-704; Scaffold::wholeSearchDir2airMesh).  searchDir
 * NULL = the device search direction.  *alpha is in/out (start value, usually 1.0). */
int ocb_step_bound(ocb_ctx* ctx, const double* searchDir, double* alpha);

/* ---- a14: line search — Optimizer::lineSearch + stepForward (Optimizer.cpp:575-673),
 * Scaffold::stepForward (Scaffold.cpp:282-293), TriMesh::checkInversion (TriMesh.cpp:1710-1756).
 * Uses the device search direction; alpha0 is the starting step (already x0.99).  E_last is
 * recomputed when a scaffold is present (Optimizer.cpp:590) or when E_last <= 0 is passed.  Outputs: accepted alpha, new total
 * energy, its scaffold part, its mesh SD part (unscaled), lastEDec (scaffold change excluded,
 * :631-634), number of halvings, stop flag (:635, honoured only if allowEDecRelTol). */
typedef struct {
    double alpha, E_new, E_scaf_new, E_sd_new, E_last, lastEDec;
    int n_halvings, stopped;
} ocb_linesearch_result;
int ocb_line_search(ocb_ctx* ctx, double energyParam0, double E_last, double alpha0,
                    int allowEDecRelTol, ocb_linesearch_result* out);
/* x = x0 + alpha * p without evaluation (Optimizer::stepForward) */
int ocb_step_forward(ocb_ctx* ctx, double alpha);

/* ---- one whole Newton iteration of Optimizer::solve(1) (Optimizer.cpp:203-261, 505-573):
 * gradient (+ energy at x, same pass) -> convergence test -> Hessian -> preconditioner -> PCG -> step bound -> line
 * search, device resident, three host round trips.  ocb_newton_step_ex lets the host program's own control flow keep
 * the parts it already did:
 *   OCB_STEP_REUSE_GRADIENT        ocb_gradient ran at this x with this energyParam0 (Optimizer::solve computes the
 *                                  gradient and tests convergence itself, Optimizer.cpp:209-221)
 *   OCB_STEP_REUSE_MATRIX          ocb_hessian_assemble ran at this x (fractureInitiated: the topology step assembled
 *                                  the matrix for the Newton step that follows it, Optimizer.cpp:428, 512-514, 567)
 *   OCB_STEP_SKIP_CONVERGENCE_TEST the caller has tested ||g||^2 < targetGRes already
 * pcg_max_it <= 0: max(10000, 40 sqrt(n)) iterations, then the current iterate is used (inexact Newton).
 * ms_solve / ms_line_search: host wall clock of {assembly, set-up, PCG} and {step bound, line search} (the reference's
 * timer_step activities 0-4 and 5). */
#define OCB_STEP_REUSE_GRADIENT 1
#define OCB_STEP_REUSE_MATRIX 2
#define OCB_STEP_SKIP_CONVERGENCE_TEST 4
typedef struct {
    double sqn_g, targetGRes, alpha, E_new, E_scaf_new, E_sd_new, lastEDec, pcg_rel_res;
    int converged, stopped, n_halvings, pcg_iters;
    double alpha_init, E_last, ms_solve, ms_line_search;
    int pcg_status, reserved;    /* pcg_status: 0, OCB_ERR_NOT_CONVERGED (max_it) or OCB_ERR_BREAKDOWN (truncated CG: d.Ad <= 0, the iterate reached
                                  * so far was used); reserved: how often the diagonal was lifted after a breakdown (0 on healthy systems) */
} ocb_newton_result;
int ocb_newton_step(ocb_ctx* ctx, double energyParam0, double targetGRes, double pcg_rel_tol,
                    int pcg_max_it, int allowEDecRelTol, ocb_newton_result* out);
int ocb_newton_step_ex(ocb_ctx* ctx, double energyParam0, double targetGRes, double pcg_rel_tol,
                       int pcg_max_it, int allowEDecRelTol, int flags, ocb_newton_result* out);
/* what the last gradient pass left besides g: ||g||^2, ||g_mesh||^2 (the unscaled mesh term, gradient_ET[0] of
 * Optimizer::writeGradL2NormToFile), E_sd and E_scaf at x.  OCB_ERR_STATE when x moved since. */
int ocb_gradient_info(ocb_ctx* ctx, double* info4);

/* ---- a15: seam energy — TriMesh::computeSeamSparsity (TriMesh.cpp:1542-1558); returns
 * (sum + initSeamLen) / virtualRadius in *E_se (Optimizer.cpp:709-710). */
int ocb_seam_energy(ocb_ctx* ctx, int nCoh, const int32_t* cohE_colmajor /* nCoh x 4 */,
                    const double* edgeLen, const int32_t* boundaryEdge, double initSeamLen,
                    double virtualRadius, double avgEdgeLen, int triSoup, double* E_se);

/* ---- a8: candidate filter score — SymDirichletEnergy::computeLocalGradient (:215-256) +
 * computeDivGradPerVert (:108-149): per-vertex sample std-dev of incident corner gradients. */
int ocb_divgrad_scores(ocb_ctx* ctx, double* perVert_nV);

/* ---- a16/a17 (non-bijective local stencils): batched evaluation of candidate operations —
 * TriMesh::computeLocalLDec -> computeLocalEdDec_* -> nested dense Optimizer (TriMesh.cpp:2105-2794).
 * The caller builds the local meshes exactly as the reference does (local TriMesh after the local split /
 * merge, 1-2 free vertices, TriMesh.cpp:2270-2300, 2330) and packs them into a ragged batch; one thread
 * block per stencil runs Optimizer::precompute + setRelGL2Tol(relGL2Tol) + solve(maxIter) (dense LDL^T in
 * shared memory, energyParams = {1}, weights = local area fractions) and returns E_init and E_final of the
 * local energy (the caller forms eDec = (initE - E_final) * A_local / A_total, TriMesh.cpp:2380-2381).
 * score[i] = score_scale[i] * (E_init[i] - E_final[i]) + score_offset[i] (NULL: 1 and 0), e.g.
 * (1 - lambda) * A_local/A_total and -lambda * seInc (TriMesh.cpp:2184-2186); *argmax = first maximum of
 * score (TriMesh.cpp:726-731).  Limits: 64 local vertices, 96 triangles, 16 free vertices per stencil
 * (status -2 beyond); status -4 = inverted input. */
typedef struct {
    int nStencil;
    const int32_t* vert_ptr;   /* nStencil+1: local vertex ranges */
    const int32_t* tri_ptr;    /* nStencil+1: local triangle ranges */
    const double*  V_rest;     /* 3 per local vertex (x y z interleaved) */
    const double*  UV;         /* 2 per local vertex (interleaved) */
    const int32_t* F;          /* 3 per local triangle, LOCAL vertex ids */
    const uint8_t* is_free;    /* per local vertex: 1 = free DOF, 0 = fixed */
    const double*  score_scale;   /* nStencil or NULL */
    const double*  score_offset;  /* nStencil or NULL */
} ocb_stencil_batch;
int ocb_eval_stencils(ocb_ctx* ctx, const ocb_stencil_batch* batch, int maxIter, double relGL2Tol,
                      double* E_init, double* E_final, double* UV_out, int32_t* iters, double* score,
                      int32_t* status, int* argmax);

/* ---- a16/a17 with bijectivity on: ONE Newton iteration of the nested Optimizer of every candidate in the batch.
 * The nested Optimizer re-triangulates the candidate's local air mesh after every iteration (Optimizer.cpp:252-257 ->
 * Scaffold.cpp:153-199, Triangle "qYQ"): Triangle stays with the host, so the caller advances all candidates in lock
 * step -- triangulate the air region of every active candidate, pack local mesh + air mesh, call this, read the new UVs
 * back (shim/CudaCandidates.cpp does exactly that inside TriMesh::querySplit / queryMerge).  A stencil's vertices are the
 * local mesh's first, then the air mesh's own vertices (outer loop + Steiner points); its triangles the mesh's first
 * (area weights), then the air mesh's (uniform weights, scaled by w_scaf / #air triangles; rest shape = the positions
 * handed in, clamped by area_thres: TriMesh.cpp:373-383).  Per stencil: result = 0 step taken, 1 converged before a step
 * (||g||^2 < target_gres, Optimizer.cpp:215-221), 2 step taken and the line search says stop (:635), -2 over the limits
 * (128 vertices, 192 triangles, 32 free vertices), -4 inverted input, -5 non-finite result; out6 = E_SD of the mesh at the returned UVs,
 * E incl. scaffold, ||g||^2, accepted step, E before the step, lastEDec. */
typedef struct {
    int nStencil;
    const int32_t* vert_ptr;     /* nStencil+1 */
    const int32_t* tri_ptr;      /* nStencil+1 */
    const int32_t* n_mesh_vert;  /* nStencil */
    const int32_t* n_mesh_tri;   /* nStencil */
    const double*  V_rest;       /* 3 per vertex (x y z interleaved; ignored for air vertices) */
    const double*  UV;           /* 2 per vertex (interleaved) */
    const int32_t* F;            /* 3 per triangle, LOCAL vertex ids */
    const uint8_t* is_free;      /* per vertex: 1 = free DOF */
    const double*  area_thres;   /* nStencil: areaThres_AM of the local air mesh (Scaffold.cpp:156,176) */
    const double*  target_gres;  /* nStencil: Optimizer::updateTargetGRes of the local problem (Optimizer.cpp:675-678) */
    double w_scaf;               /* 0.01 for the nested optimizers (Optimizer.cpp:87 with energyParams = {1}) */
} ocb_stencil_step_batch;
int ocb_stencil_newton_step(ocb_ctx* ctx, const ocb_stencil_step_batch* batch, double* UV_out, double* out6, int32_t* result);

/* ---- a12 (solver set-up): the multilevel additive Schwarz preconditioner that replaces the numeric
 * factorisation (EigenLibSolver.cpp:80-93) is rebuilt by ocb_factorize; its hierarchy (row order along a Hilbert curve
 * through the UVs, leaves of <= 8 vertices, groups of 8, 6 affine DOFs per node) is built with
 * the sparsity pattern.  info[0] = 1 if active (0: block-Jacobi only, no UV was known at pattern time),
 * info[1] = levels L, info[2] = CTA-local levels, info[3] = persistent CTAs, info[4..4+L) = nodes per level,
 * info[15] = solves that were repeated with block-Jacobi because the preconditioner came out indefinite. */
int ocb_precond_info(const ocb_ctx* ctx, int32_t* info16);
/* A bare LinSysSolver (ocb_set_pattern + ocb_update_values_triplets, no mesh on this context) has no geometry: the
 * caller may hand over positions of the first n vertices of the NEXT ocb_set_pattern (Eigen column-major n x 2: all x,
 * then all y — e.g. TriMesh::V of the mesh being optimised); vertices beyond n (interior air-mesh vertices) are placed
 * at the mean of their neighbours.  Purely a hint for the preconditioner hierarchy: results do not depend on it,
 * without it the solver falls back to block-Jacobi.  n = 0 clears the hint. */
int ocb_set_coordinate_hint(ocb_ctx* ctx, int n, const double* xy_colmajor);
/* Host-only (no CUDA work; unit tests and tools/mas_proto.py): the hierarchy for n free points xy (2 per point)
 * and `grid` CTAs.  vert_of[n] = point of every solver row; child_beg = the per-level child ranges, level after
 * level (nodes_l + 1 entries each, `cap` entries available).  Returns the entries written or an error. */
int ocb_precond_hierarchy(ocb_ctx* ctx, int n, const double* xy, int grid, int32_t* vert_of, int32_t* info16,
                          int32_t* child_beg, int cap);

#ifdef __cplusplus
}
#endif
#endif /* OPTCUTS_B200_H */
