/* nfa.h -- C ABI of the B200 finite-element assembly library (libnfa.so).
 *
 * Drop-in boundary for the assembly hot path of OmarDuran/nfempy (R = the reference tree).  The reference is
 * pure Python; the interface each entry point replaces is cited as R/<file>:<lines>.  All pointers named
 * d_* are DEVICE pointers owned by the caller (the Python host passes torch tensors' data_ptr()); the library
 * allocates nothing the caller does not pass in except the opaque plan objects.  Every function returns 0 on
 * success and a non-zero status otherwise, with the text available from nfa_last_error().  All work is
 * stream-ordered on the caller's stream (a cudaStream_t passed as void*; NULL = default stream).
 *
 * Index types: int32 for dof and node indices (PETSc AIJ indices are 32-bit, R/.../md_primal_cosserat.py:415),
 * int64 for counts.  All floating point is IEEE double.
 */
#ifndef NFA_H
#define NFA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfa_plan nfa_plan;   /* scalar sparsity plan: CSR pattern + deterministic gather lists */
typedef struct nfa_bplan nfa_bplan; /* node-block plan for the fused H1 vector/scalar assembly */
typedef void *nfa_stream;           /* cudaStream_t */

int nfa_version(void);
const char *nfa_last_error(void);
/* number of kernels launched by this library since load (bench.py reports it as gpu_launches) */
int64_t nfa_launch_count(void);
/* when enabled, the fused assembly kernel is bracketed by CUDA events on its own stream; nfa_timing_read
 * synchronises, returns the summed device time (ms) and the number of bracketed launches, and resets */
void nfa_timing_enable(int on);
int nfa_timing_read(double *total_ms, int64_t *count);

/* ---- K1 geometry: replaces evaluate_mapping + _compute_det_and_pseudo_inverse, R/src/geometry/mapping.py:23-55.
 * Simplex cells are affine, so J, |det| and J^+ are per cell (the reference recomputes them per quadrature
 * point by QR).  dim = 1, 2, 3 (cell embedded in 3-space).  jac [n,3,dim], det [n] (always >= 0),
 * inv [n,dim,3] = Moore-Penrose inverse.  Any output pointer may be NULL. */
int nfa_geom_simplex(int dim, const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells,
                     double *d_jac, double *d_det, double *d_inv, nfa_stream stream);

/* x_q = sum_v phi_v(xi_q) X_v: physical quadrature points handed to the Python coefficient callables,
 * R/src/geometry/mapping.py:46.  shape1 [n_q, dim+1] = P1 shape values at the reference points; x [n,n_q,3]. */
int nfa_map_points(int dim, int n_q, const double *d_shape1, const double *d_coords,
                   const int32_t *d_cell_nodes, int64_t n_cells, double *d_x, nfa_stream stream);

/* ---- K2 H1 volume forms, general coefficients (quadrature loop as in the reference).
 * tab [1+dim, n_q, n_phi]: reference values (slab 0) and reference gradients (slabs 1..dim) of the scalar
 * Lagrange basis, weights [n_q].  Coefficients are addressed as c[cell*stride_cell + q*stride_q] so that a
 * constant (0,0), a per-cell (1,0) and a per-point (n_q,1) array all fit.  d_f [n, n_q, n_comp] or NULL.
 * d_alpha (global vector) + d_dest [n, n_loc] or NULL for alpha = 0.  Outputs: Ke [n, n_loc, n_loc] row-major,
 * re [n, n_loc]; local dof = basis * n_comp + comp.
 *
 * nfa_local_h1_elasticity: LEPrimalWeakForm.evaluate_form, R/src/weak_forms/le_primal_weak_form.py:13-129
 *   (r = +int phi.f + int sigma(u_h):grad phi, J = dr/dalpha), n_comp = dim.
 * nfa_local_h1_laplace: LaplacePrimalWeakForm.evaluate_form, R/src/weak_forms/laplace_primal_weak_form.py:10-74
 *   (r = -int phi f + int kappa grad u_h . grad phi), n_comp = 1. */
int nfa_local_h1_elasticity(int dim, int n_phi, int n_q, const double *d_tab, const double *d_weights,
                            const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells,
                            const double *d_lambda, int64_t lambda_stride_cell, int64_t lambda_stride_q,
                            const double *d_mu, int64_t mu_stride_cell, int64_t mu_stride_q,
                            const double *d_f, const double *d_alpha, const int32_t *d_dest,
                            double *d_Ke, double *d_re, nfa_stream stream);
int nfa_local_h1_laplace(int dim, int n_phi, int n_q, const double *d_tab, const double *d_weights,
                         const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells,
                         const double *d_kappa, int64_t kappa_stride_cell, int64_t kappa_stride_q,
                         const double *d_f, const double *d_alpha, const int32_t *d_dest,
                         double *d_Ke, double *d_re, nfa_stream stream);

/* load vector only: re [n, n_phi*n_comp] = sign * int phi_a f_i (the load term of the two forms above; used by
 * the fused path, whose matrix kernel never forms local data).  tab0 [n_q, n_phi], f [n, n_q, n_comp]. */
int nfa_local_h1_load(int dim, int n_phi, int n_q, int n_comp, const double *d_tab0, const double *d_weights,
                      const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_f,
                      double sign, double *d_re, nfa_stream stream);
/* the same with padded node records: re [n, n_phi, comp_pitch] (comp_pitch = 4 for vector P2 on tetrahedra: one 32-byte
 * sector per (cell, node), what nfa_bplan_gather_load and nfa_assemble_p2tet_elasticity_load gather) */
int nfa_local_h1_load_pitched(int dim, int n_phi, int n_q, int n_comp, const double *d_tab0, const double *d_weights,
                              const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_f,
                              double sign, double *d_re, int comp_pitch, nfa_stream stream);

/* ---- K5 H1 facet (boundary) forms on (dim-1)-simplices embedded in 3-space.
 * J = beta * int phi phi^T per component, r = r_sign * beta * int g_c phi  (g [n, n_q, n_comp] host-evaluated).
 * Dirichlet penalty (beta = 1e12, r_sign = -1): LEPrimalWeakFormBCDirichlet R/.../le_primal_weak_form.py:131-176,
 * LaplacePrimalWeakFormBCDirichlet R/.../laplace_primal_weak_form.py:77-157.  Neumann load (with_matrix = 0,
 * beta = 1, r_sign = +1): LEPrimalWeakFormBCNeumann R/.../le_primal_weak_form.py:179-221.
 * d_comp_scale [n, n_comp] or NULL scales the (c, c) matrix block of each facet -- n_c^2 for the normal-only penalty
 * of LEPrimalWeakFormBCNormalDirichlet, R/.../le_primal_weak_form.py:224-286 (whose load is g_c = (u_D . n) n_c).
 * facet_dim = dim-1 (0 = point facets of 1-D cells); facet_dim = dim makes it the Lagrange mass matrix and load of
 * L2ProjectorWeakForm on volume cells (R/src/weak_forms/l2_projector_weak_form.py:40-68). */
int nfa_local_h1_facet(int facet_dim, int n_phi, int n_q, int n_comp, const double *d_tab0,
                       const double *d_weights, const double *d_coords, const int32_t *d_facet_nodes,
                       int64_t n_facets, const double *d_g, const double *d_comp_scale, double beta, double r_sign,
                       int with_matrix, double *d_Ke, double *d_re, nfa_stream stream);

/* ---- K3 H(div) x L2 forms (mixed Poisson).  n_u vector basis functions whose first (dim+1)*nfd are the facet dofs
 * (nfd per facet, facet-major), n_p scalar L2 functions; local dofs = [u | p].
 * u_val [n_q, n_u, dim], u_div [n_q, n_u]: REFERENCE values and reference divergence tr(grad_hat phi_hat);
 * p_val [n_q, n_p]; Tcases [n_cases, nfd, nfd]: orientation matrices (R/src/basis/permute_and_transform.py);
 * cases int8 [n_cells, dim+1]: which matrix each facet of each cell takes (0 = identity).
 * nfa_local_hdiv_l2: LaplaceDualWeakForm.evaluate_form, R/src/weak_forms/laplace_dual_weak_form.py:10-87:
 *   J = [[int kappa^-1 psi_j.psi_i, -int w_l div psi_i], [int div psi_j w_k, 0]], r = [0, -int f w]  (the reference
 *   zeroes alpha, :50).  kappa stride-addressed like the H1 coefficients; f [n, n_q] or NULL.  With n_p = 0, kappa = 1
 *   and the vector load f_u [n, n_q, dim] (r_u = -int f_u . psi) this is the H(div) block of L2ProjectorWeakForm,
 *   R/src/weak_forms/l2_projector_weak_form.py:44-74.
 * nfa_local_hdiv_facet: LaplaceDualWeakFormBCDirichlet, R/...:90-154: r_k = int_F p_D (psi_k . n) over the facet dofs
 *   of the first neighbour cell; trace_tab [n_tabs, n_q, nfd, dim] = reference values of the facet's dofs at the facet
 *   quadrature points mapped into the neighbour (R/src/basis/parametric_transformation.py:58-98), one table per
 *   (local facet, vertex permutation); tidx int32 [n] picks the table; cases int8 [n]; normals [n,3]; pd [n, n_q]. */
int nfa_local_hdiv_l2(int dim, int n_u, int n_p, int n_q, int nfd, const double *d_u_val, const double *d_u_div,
                      const double *d_p_val, const double *d_weights, const double *d_Tcases, const int8_t *d_cases,
                      const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_kappa,
                      int64_t kappa_stride_cell, int64_t kappa_stride_q, const double *d_f, const double *d_f_u,
                      double *d_Ke, double *d_re, nfa_stream stream);
int nfa_local_hdiv_facet(int dim, int nfd, int n_q, const double *d_trace_tab, const int32_t *d_tidx,
                         const double *d_weights, const double *d_Tcases, const int8_t *d_cases, const double *d_coords,
                         const int32_t *d_facet_nodes, const int32_t *d_neigh_nodes, const double *d_normals,
                         const double *d_pd, int64_t n_facets, double *d_re, nfa_stream stream);

/* Generalised H(div) x L2 block form for the Arbogast-Taicher family: DegenerateEllipticWeakForm
 * (R/src/weak_forms/degenerate_elliptic_weak_form.py:41-144), UnscaledEllipticWeakForm
 * (R/src/weak_forms/unscaled_elliptic_weak_form.py:41-153) and ScaledToPhysicalL2Projection
 * (R/src/weak_forms/scaled_to_physical_l2_projection.py:27-105).  With D_i = c_d div psi_i + g . psi_i:
 *   J = [[int c_uu psi_j.psi_i, -int w_l D_i], [int D_j w_k, int c_pp w_l w_k]],
 *   r = J alpha - [int s_u (beta_u.psi).psi_i ; int s_p (beta_p w) w_k] - [0 ; int load_p w_k].
 * Per-point arrays [n, n_q] (g: [n, n_q, dim]); c_d NULL = 1, g / c_pp / load_p / s_u / s_p NULL = absent;
 * alpha, beta: global vectors gathered through d_dest [n, n_u + n_p] (NULL = 0). */
int nfa_local_hdiv_l2_general(int dim, int n_u, int n_p, int n_q, int nfd, const double *d_u_val,
                              const double *d_u_div, const double *d_p_val, const double *d_weights,
                              const double *d_Tcases, const int8_t *d_cases, const double *d_coords,
                              const int32_t *d_cell_nodes, int64_t n_cells, const double *d_c_uu, const double *d_c_d,
                              const double *d_g, const double *d_c_pp, const double *d_load_p, const double *d_alpha,
                              const double *d_beta, const int32_t *d_dest, const double *d_s_u, const double *d_s_p,
                              double *d_Ke, double *d_re, nfa_stream stream);

/* Two-flux mixed transport with reaction, fields [mp | mc | p | c] (config 5): SundusDualWeakForm.evaluate_form,
 * R/examples/parabolic_diffusive_problem/SundusWeakForm.py:11-177.  mp, mc share the H(div) tables, p, c the L2 table.
 * Residual and tangent at the current iterate d_alpha (global vector, gathered through d_dest [n, 2 n_u + 2 n_p]) with
 * the previous time level d_alpha_prev; R = 1 + eta c_h^2 gives the only nonlinear tangent term 2 eta c_h (w, w).
 * kappa, delta, eta stride-addressed; f_f, f_r [n, n_q] or NULL. */
int nfa_local_sundus(int dim, int n_u, int n_p, int n_q, int nfd, const double *d_u_val, const double *d_u_div,
                     const double *d_p_val, const double *d_weights, const double *d_Tcases, const int8_t *d_cases,
                     const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_kappa,
                     int64_t kappa_stride_cell, int64_t kappa_stride_q, const double *d_delta, int64_t delta_stride_cell,
                     int64_t delta_stride_q, const double *d_eta, int64_t eta_stride_cell, int64_t eta_stride_q,
                     double delta_t, const double *d_f_f, const double *d_f_r, const double *d_alpha,
                     const double *d_alpha_prev, const int32_t *d_dest, double *d_Ke, double *d_re, nfa_stream stream);

/* ---- K6 sparsity plan: replaces the implicit pattern growth of PETSc Mat.setValue(..., addv=True) without
 * preallocation (R/src/assembly/ScatterFormData.py:5-11): the pattern is the union of
 * dest x dest over all items of all segments ("full" policy: explicit zeros kept), columns sorted per row.
 * A segment is one form's destination table d_dest[s] = int32 [n_items[s], n_loc[s]].
 * base == NULL: the pattern is defined by these segments.  base != NULL: the segments are merged INTO the
 * pattern of `base` (every pair must already be present, else an error) -- used to add boundary forms to a
 * pattern produced by the fused path.  seg_flags[s] & 1: the segment contributes matrix entries (otherwise
 * right-hand-side rows only, e.g. the volume load of the fused path); bit 1 set: the segment's right-hand side is summed
 * elsewhere (nfa_bplan_gather_load) and stays out of the plan's row lists; seg_flags == NULL means matrix + rhs for all. */
int nfa_plan_create(int64_t n_dof, int n_segments, const int32_t *const *d_dest, const int64_t *n_items,
                    const int32_t *n_loc, const int32_t *seg_flags, const nfa_plan *base, nfa_plan **out,
                    nfa_stream stream);
/* adopt the pattern of a node-block plan (no gather lists) */
/* Value-dependent ("filtered") patterns of the example drivers' inline scatter closures
 * (R/examples/elasticity/cg_elasticity_example_1.py:92-145, R/examples/cosserat_elasticity/
 * mfe_cosserat_elasticity_example_3.py:150-243): d_masks[k] is a uint8 [n_items, n_loc, n_loc] table, 0 = the local
 * entry is left out of the pattern and of the sums.  The host layer builds the masks from the local matrices
 * (~isclose(v, 1e-16), optional row <= col for the symmetric "sbaij" storage, optional explicit diagonal). */
int nfa_plan_create_filtered(int64_t n_dof, int n_segments, const int32_t *const *d_dest, const int64_t *n_items,
                             const int32_t *n_loc, const int32_t *seg_flags, const uint8_t *const *d_masks,
                             nfa_plan **out, nfa_stream stream);
/* masks of nfa_plan_create_filtered from the local matrices, in one pass: kept = |v - 1e-16| > tol (~np.isclose(v, 1e-16));
 * d_mask = kept, and dest[a] <= dest[b] when `upper`, and 1 on the local diagonal when `force_diag`; entries that fail the
 * value test are zeroed in d_Ke (a forced diagonal entry then adds 0.0). */
int nfa_filter_local_entries(int64_t n_items, int n_loc, const int32_t *d_dest, double tol, int upper, int force_diag,
                             double *d_Ke, uint8_t *d_mask, nfa_stream stream);
int nfa_plan_from_bplan(const nfa_bplan *bp, nfa_plan **out, nfa_stream stream);
void nfa_plan_destroy(nfa_plan *plan);
int64_t nfa_plan_nnz(const nfa_plan *plan);
int64_t nfa_plan_n_dof(const nfa_plan *plan);
int64_t nfa_plan_bytes(const nfa_plan *plan);
/* copy the pattern out: indptr int32 [n_dof+1], indices int32 [nnz] (device) -- A.getValuesCSR() index arrays */
int nfa_plan_csr(const nfa_plan *plan, int32_t *d_indptr, int32_t *d_indices, nfa_stream stream);

/* ---- K7 deterministic reduce: replaces the n_loc^2 setValue calls and res_g[dest[k]] += res_l[k]
 * (R/src/assembly/ScatterFormData.py:5-25).  Every CSR value / rhs row sums its contributions in ascending
 * (segment, item, local row, local col) order = the reference's call order; no atomics.
 * d_Ke[s] [n_items[s], n_loc[s]^2], d_re[s] [n_items[s], n_loc[s]] (either array may be NULL to skip).
 * accumulate != 0 adds to the existing d_vals / d_rhs (PETSc ADD_VALUES across scatter_forms calls). */
int nfa_plan_reduce(const nfa_plan *plan, const double *const *d_Ke, const double *const *d_re,
                    int accumulate_vals, double *d_vals, int accumulate_rhs, double *d_rhs,
                    nfa_stream stream);

/* y (+)= A x on the plan's pattern, one warp per row, fixed order (residual term J.alpha of the fused path) */
int nfa_csr_spmv(const nfa_plan *plan, const double *d_vals, const double *d_x, int accumulate, double *d_y,
                 nfa_stream stream);

/* ---- streamed scatter (csrc/nfa_stream.cu): "full"-pattern assembly of forms whose local matrices cannot all be resident
 * (BASELINE config 4: LCEScaledDualWeakForm at BDM_3 x RT_3 x P_2 x P_2, 348 local dofs, R/src/weak_forms/
 * lce_scaled_dual_weak_form.py:445-699) in spaces where a dof belongs to at most two cells.  The CSR position of the local
 * entry (cell, a, b) is d_indptr[d_dest[cell][a]] + (d_tabs[d_tab_of[cell][a]][b] & 0x7fff): a rank table per cell (cell-
 * local rows: rank inside the cell's sorted dofs) and per ordered (cell, partner) pair (rank inside the sorted union of
 * both cells' dofs).  Bit 15 of a table entry: the entry is ADDED in pass 1 (the row's lower cell stored it in pass 0);
 * clear: STORED in pass 0.  Chunks of cells in ascending order, pass 0 then pass 1 for every chunk: each value is the sum
 * of its at most two contributions in ascending cell order (R/src/assembly/ScatterFormData.py:5-11), no atomics.
 * d_indptr is int64 (such matrices exceed 2^31 stored entries).  nfa_rows_copy_i32 fills the column indices: row r
 * receives d_src[d_src_off[r] ..) (its cell's sorted dofs or its pair's sorted union). */
int nfa_stream_scatter(int64_t n_cells, int n_loc, const int32_t *d_dest, const int32_t *d_tab_of, const uint16_t *d_tabs,
                       const int64_t *d_indptr, const double *d_Ke, double *d_vals, int pass, nfa_stream stream);
int nfa_rows_copy_i32(int64_t n_rows, const int64_t *d_indptr, const int64_t *d_src_off, const int32_t *d_src,
                      int32_t *d_out, nfa_stream stream);

/* ---- lowest-order H(div) x L2 forms by row owner (csrc/nfa_hdivrows.cu): no local matrices in HBM.
 * Replaces, for LaplaceDualWeakForm on RT_1 x P_0 (R/src/weak_forms/laplace_dual_weak_form.py:10-87; form = 0, local order
 * [u | p]) and SundusDualWeakForm on RT_1 x RT_1 x P_0 x P_0 (R/examples/parabolic_diffusive_problem/SundusWeakForm.py:
 * 11-177; form = 1, local order [mp | mc | p | c]), the per-cell evaluate_form + n_loc^2 Mat.setValue calls of
 * R/src/assembly/ScatterFormData.py:5-25.
 * nfa_hdiv_rt1_records: one thread per cell writes the cell's distinct numbers as a record of `stride` doubles
 *   (stride a multiple of 4): for flux field F and local facet a, at (F n_u + a) w, w = n_u + 2 rounded up to even:
 *   M_F[a][0..n_u), C[a], r_F[a]
 *   (M = int w psi_a.psi_b / kappa resp. / delta, C = int div psi_a w, r = residual row at the iterate); then the block the
 *   scalar rows read: form 0: r_p, C[0..n_u); form 1: K_pp, K_cc, r_p, r_c, C[0..n_u).  Arguments as nfa_local_hdiv_l2 /
 *   nfa_local_sundus (d_f_f is the load `rhs` of form 0; delta, eta, delta_t, f_r, alpha, alpha_prev, dest are read by
 *   form 1 only).
 *   form = 2: the generalised mixed elliptic forms (DegenerateEllipticWeakForm, UnscaledEllipticWeakForm,
 *   R/src/weak_forms/degenerate_elliptic_weak_form.py:41-144, unscaled_elliptic_weak_form.py:41-153) on RT_1 x P_0, local
 *   order [u | p]: d_kappa = c_uu (multiplies the flux mass), d_delta = c_d (scales div psi; NULL = 1), d_eta = c_pp (weight
 *   of the (p, p) mass; NULL = no such block), d_f_f = load_p, d_f_r = g [n_cells][n_q][dim] (adds g . psi to the coupling),
 *   r = J alpha - [0; int load_p w]; record: flux rows as above, then K_pp, r_p, C[0..n_u).
 * Runs: local row a of an item reads the record doubles [d_roff[a], d_roff[a] + 8) (d_roff even); d_codes int8 [n_loc][8]
 *   says what entry k of the run is for that row: 0 = nothing, +-(b + 1) = local column b with that sign (each column at
 *   most once; columns without a code are structural zeros, stored as 0.0), 127 = the row's residual.
 * nfa_rows_locate (one-off): d_rows int32 [n_rows] lists the rows the items feed (ascending), d_row_src int32 [n_rows][2]
 *   names the one or two sources of each as item * 32 + local row (-1 = none; ascending item order), d_dest int32
 *   [n_items][n_loc] the items' global dofs; writes d_slots uint8 [n_rows][2][8] = position of every run entry's column
 *   inside the row's CSR segment.
 * nfa_rows_expand: d_blocks int32 [n_blocks][4] = (first row, number of CONSECUTIVE rows <= 256, index of the first row
 *   in d_rows / d_row_src / d_slots, 0); one CTA pass per block, one thread per row; every CSR value of the listed rows
 *   is overwritten with the sum of its sources' entries, lower item first; the right-hand side row is the sum of the
 *   sources' residual entries (rhs_add: added to d_rhs instead of stored).  Rows outside the blocks are not touched. */
int nfa_hdiv_rt1_records(int dim, int form, int n_q, const double *d_u_val, const double *d_u_div, const double *d_p_val,
                         const double *d_weights, const double *d_Tcases, const int8_t *d_cases, const double *d_coords,
                         const int32_t *d_cell_nodes, int64_t n_cells, const double *d_kappa, int64_t ksc, int64_t ksq,
                         const double *d_delta, int64_t dsc, int64_t dsq, const double *d_eta, int64_t esc, int64_t esq,
                         double delta_t, const double *d_f_f, const double *d_f_r, const double *d_alpha,
                         const double *d_alpha_prev, const int32_t *d_dest, int stride, double *d_rec, nfa_stream stream);
int nfa_rows_locate(int64_t n_rows, const int32_t *d_rows, const int32_t *d_indptr, const int32_t *d_indices,
                    const int32_t *d_row_src, int n_loc, const int8_t *d_codes, const int32_t *d_dest, uint8_t *d_slots,
                    nfa_stream stream);
int nfa_rows_expand(int64_t n_blocks, const int32_t *d_blocks, int max_row_len, const int32_t *d_indptr, const int32_t *d_row_src,
                    const uint8_t *d_slots, int n_loc, const int32_t *d_roff, const int8_t *d_codes, const double *d_rec,
                    int stride, double *d_vals, double *d_rhs, int rhs_add, nfa_stream stream);
/* pattern-only plan over a caller-owned CSR (kept alive by the caller): base of nfa_plan_create for the other segments */
int nfa_plan_from_csr(int64_t n_dof, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices, nfa_plan **out);

/* ---- K8 multi-GPU interface rows: after the local reduce a rank holds partial sums on rows owned by a
 * lower rank.  nfa_pack_rows copies the CSR segments of `rows` (int32 global row ids, ascending) into a contiguous
 * send buffer (buf_off[k] = prefix sum of the row lengths, int64); the owner calls nfa_unpack_add_rows on the
 * received buffer, once per sender in ascending rank order (fixed order => deterministic).  Both sides hold the same
 * pattern for these rows, so segments line up entry by entry.  The transfer itself is ncclSend/ncclRecv issued by the
 * host layer (torch.distributed) between the two calls. */
int nfa_pack_rows(const int32_t *d_indptr, const double *d_vals, const int32_t *d_rows, const int64_t *d_buf_off,
                  int64_t n_rows, double *d_buf, nfa_stream stream);
int nfa_unpack_add_rows(const int32_t *d_indptr, double *d_vals, const int32_t *d_rows, const int64_t *d_buf_off,
                        int64_t n_rows, const double *d_buf, nfa_stream stream);

/* ---- fused H1 path (affine cells, cell-wise constant coefficients): local B^T D B contraction and the global
 * reduce in ONE kernel, no local matrices in HBM.  The node-block plan groups the n_comp x n_comp blocks of the
 * CSR by (row node, col node); d_node_dest int32 [n_cells, n_phi] = dest[:, ::n_comp] / n_comp. */
int nfa_bplan_create(int64_t n_nodes, int n_phi, int n_comp, const int32_t *d_node_dest, int64_t n_cells,
                     nfa_bplan **out, nfa_stream stream);
void nfa_bplan_destroy(nfa_bplan *bp);
/* Right-hand side of the fused form by ROW OWNER: row n_comp N + c = sum over the (cell, local node) pairs that are node N,
 * in ascending cell order, of d_re[(cell n_phi + a) comp_pitch + c] -- res_g[dest] += r_el of
 * R/src/assembly/ScatterFormData.py:14-18 without the per-row entry lists of nfa_plan_reduce (one 32-bit entry per node
 * and cell instead of one per row and cell; comp_pitch = 4 with three components: one 32-byte sector per entry).
 * overwrite != 0: the sums replace the rows (rows of nodes without cells become 0), else they are added. */
int nfa_bplan_gather_load(const nfa_bplan *bp, const double *d_re, int comp_pitch, int overwrite, double *d_rhs,
                          nfa_stream stream);
/* Load vector AND its scatter in two kernels, without the local load vectors in HBM ("tile-merged"): tiles of 64
 * consecutive cells compute their local load vectors into shared memory (d_f [n_cells, n_q, n_comp] streamed by bulk
 * copies, like nfa_local_h1_load), sum them per node of the tile (ascending cell) and write one record per (tile, node)
 * pair -- about 3 per cell instead of n_phi; the second kernel adds a node's pair sums in ascending tile order into rows
 * n_comp N + c of d_rhs.  Same terms as res_g[dest] += r_el (R/src/assembly/ScatterFormData.py:14-18), associated tile by
 * tile.  Vector P2 / P1 on tetrahedra, vector P2 on triangles; n_q <= 64.  Only the rows of nodes that HAVE cells in this
 * plan are visited (nfa_bplan_n_active_nodes of them): overwrite != 0 replaces those rows and leaves every other row of
 * d_rhs alone -- a rank's local plan over the global numbering never touches the other ranks' rows. */
int nfa_bplan_load_rhs(const nfa_bplan *bp, int dim, int n_q, const double *d_tab0, const double *d_weights,
                       const double *d_coords, const int32_t *d_cell_nodes, const double *d_f, double sign, int overwrite,
                       double *d_rhs, nfa_stream stream);
int64_t nfa_bplan_n_active_nodes(const nfa_bplan *bp);
int64_t nfa_bplan_nnz(const nfa_bplan *bp);
int64_t nfa_bplan_bytes(const nfa_bplan *bp);
/* Mref [n_phi, n_phi, dim, dim] = int_ref d_k phi_a d_l phi_b (exact reference stiffness tensor).
 * lambda, mu, kappa: per cell (stride 1) or constant (stride 0).  Writes ALL nnz values of the pattern. */
int nfa_assemble_h1_elasticity_affine(const nfa_bplan *bp, int dim, const double *d_Mref,
                                      const double *d_coords, const int32_t *d_cell_nodes,
                                      const double *d_lambda, int64_t lambda_stride,
                                      const double *d_mu, int64_t mu_stride,
                                      double *d_vals, nfa_stream stream);
int nfa_assemble_h1_laplace_affine(const nfa_bplan *bp, int dim, const double *d_Mref,
                                   const double *d_coords, const int32_t *d_cell_nodes,
                                   const double *d_kappa, int64_t kappa_stride,
                                   double *d_vals, nfa_stream stream);


/* ---- mixed "tensor" forms: row-wise H(div) stresses with vector L2 multipliers (config 4).  One local kernel for the
 * block structure of the reference's vectorised forms LEDualWeakForm (R/src/weak_forms/le_dual_weak_form.py:354-510),
 * LCEDualWeakForm (lce_dual_weak_form.py:410-648), LCEScaledDualWeakForm (lce_scaled_dual_weak_form.py:445-699) and the
 * Riesz-map forms (le_riesz_map_weak_form.py:258-360, lce_riesz_map_weak_form.py:372-546,
 * lce_scaled_riesz_map_weak_form.py:415-598).  Local dofs [H(div) 0 | H(div) 1 | L2 0 | L2 1], basis-major,
 * component-minor.  All arrays are device pointers; per-point coefficient arrays are [n_cells, n_q] (g: [.., dim],
 * load: [.., n_comp]) and optional (NULL = term absent; h_div_cd NULL = 1).
 *   (F,F)  int a_id (psi_a.psi_b) d_cc' + a_tr psi_b[c] psi_a[c'] + a_vol psi_a[c] psi_b[c'] + a_dd div psi_a div psi_b d_cc'
 *   (U,F)  div_coupling: d_cc' int w phi_k (cd div^ psi_a + det g.psi_a);  asym_coupling: int w det phi_k eps(c,c',d) psi_a[d]
 *   (U,U)  d_cc' int a_mass phi_k phi_l;   r_U = -int load_c phi_k;   r += J alpha[dest] when alpha is given. */
typedef struct nfa_tensor_form {
    int32_t dim, n_q, n_hdiv, n_l2;
    const double *weights;                             /* [n_q] */
    int32_t h_n_phi[2], h_n_comp[2], h_nfd[2];
    const double *h_val[2], *h_div[2], *h_Tcases[2];   /* reference values [n_q,n_phi,dim], reference divergence, T cases */
    const int8_t *h_cases[2];                          /* [n_cells, dim+1] orientation case of every facet */
    const double *h_a_id[2], *h_a_tr[2], *h_a_vol[2], *h_a_dd[2];
    const double *h_div_cd[2], *h_div_g[2];
    int32_t l_n_phi[2], l_n_comp[2];
    const double *l_val[2], *l_a_mass[2], *l_load[2];
    int32_t div_coupling[2][2], asym_coupling[2][2];   /* [L2 field][H(div) field] */
    int32_t h_dd_general[2];                           /* a_dd block on (cd div psi + g.psi) instead of div psi */
} nfa_tensor_form;
int nfa_local_hdiv_tensor(const nfa_tensor_form *form, const double *d_coords, const int32_t *d_cell_nodes,
                          int64_t n_cells, const double *d_alpha, const int32_t *d_dest, double *d_Ke, double *d_re,
                          nfa_stream stream);
/* Natural boundary load of one H(div) field with n_comp rows: re[facet, re_off + a n_comp + c] = sign int_F g_c (psi_a.n)
 * over the facet dofs of the first neighbour (the vector-valued form of nfa_local_hdiv_facet; LCEDualWeakFormBCDirichlet
 * R/src/weak_forms/lce_dual_weak_form.py:651-756, LEDualWeakFormBCDirichlet le_dual_weak_form.py:513-578).
 * g [n_facets, n_q, n_comp]; re rows have re_stride entries. */
int nfa_local_hdiv_facet_vec(int dim, int nfd, int n_q, int n_comp, const double *d_trace_tab, const int32_t *d_tidx,
                             const double *d_weights, const double *d_Tcases, const int8_t *d_cases,
                             const double *d_coords, const int32_t *d_facet_nodes, const int32_t *d_neigh_nodes,
                             const double *d_normals, const double *d_g, int64_t n_facets, double sign, int re_off,
                             int re_stride, double *d_re, nfa_stream stream);

/* ---- error norms of a discrete solution: l2_error (R/src/postprocess/l2_error_post_processor.py:7-63) and div_error
 * (:185-237).  kind 0: Lagrange field (val [n_q, n_phi]); kind 1: values of a row-wise H(div) field (val
 * [n_q, n_phi, dim] reference values, exact [n, n_q, n_comp, dim]); kind 2: its divergence (dv [n_q, n_phi] reference
 * divergence, exact [n, n_q, n_comp]).  d_alpha is the global vector, read through d_dest[cell, dest_off + a n_comp + c]
 * (rows of dest_stride entries).  d_out [n_cells] receives each cell's squared-error contribution. */
int nfa_field_error(int dim, int kind, int n_phi, int n_comp, int n_q, int nfd, const double *d_val,
                    const double *d_div, const double *d_weights, const double *d_Tcases, const int8_t *d_cases,
                    const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_alpha,
                    const int32_t *d_dest, int dest_stride, int dest_off, const double *d_exact, double *d_out,
                    nfa_stream stream);
/* Same with kind 3: error of the SCALED divergence  l div psi + grad l . psi  of an H(div) field against d_exact
 * (div_scaled_error, R/src/postprocess/l2_error_post_processor.py:240-305); d_scale [n_cells][n_q], d_gscale
 * [n_cells][n_q][dim] are the host-evaluated l and grad l.  kinds 0..2 ignore them (nfa_field_error forwards here). */
int nfa_field_error_scaled(int dim, int kind, int n_phi, int n_comp, int n_q, int nfd, const double *d_val,
                           const double *d_div, const double *d_weights, const double *d_Tcases, const int8_t *d_cases,
                           const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_alpha,
                           const int32_t *d_dest, int dest_stride, int dest_off, const double *d_exact,
                           const double *d_scale, const double *d_gscale, double *d_out, nfa_stream stream);

/* ---- row-item fused path for vector-P2 elasticity on affine tetrahedra (config 2's headline kernel).
 * Same plan object, same output (all nnz values of the pattern, written once, deterministic) as
 * nfa_assemble_h1_elasticity_affine, i.e. the matrix part of LEPrimalWeakForm.evaluate_form
 * (R/src/weak_forms/le_primal_weak_form.py:91-126) + the scatter (R/src/assembly/ScatterFormData.py:5-11); the
 * local blocks are evaluated from the physical gradients of the barycentric coordinates, one thread per set of CSR
 * blocks fed by the same cells.  The item lists are built by the host layer (nfempy_b200/assembly/p2_items.py) from
 * the plan's block pattern and attached once:
 *   nfa_bplan_block_pattern  copies bptr int32 [n_nodes + 1] / bcols int32 [n_blocks] to caller buffers (device);
 *   nfa_bplan_set_p2_items   copies the lists (device pointers; layouts in nfa_plan.cuh / p2_items.py) into the plan.
 * Requires mu > 0 in every cell (the caller checks); lambda, mu per cell (stride 1) or constant (stride 0). */
int nfa_bplan_block_pattern(const nfa_bplan *bp, int32_t *d_bptr, int32_t *d_bcols, nfa_stream stream);
int64_t nfa_bplan_n_blocks(const nfa_bplan *bp);
int nfa_bplan_set_p2_items(nfa_bplan *bp, int64_t n_chunks, const int32_t *d_chunk_hdr, int64_t n_tasks,
                           const int32_t *d_task_hdr, int64_t n_slots, const uint16_t *d_item_hdr, int64_t ell_size,
                           const uint32_t *d_ell, int max_tile_blocks, nfa_stream stream);
/* Bank-aware order of the items (= lanes) inside their (chunk, type) groups: d_group_ptr int64 [n_groups + 1] item ranges,
 * d_hdr6 int32 [n_items][6] = three tile offsets, row pitch, two opposite-edge tile offsets (0xFFFF = unused);
 * d_perm int64 [n_items] receives the new order (new position -> old index) within every group.  A thread per group deals
 * the items into half warps whose shared-memory tile stores hit distinct banks (greedy over a window of `window` items). */
int nfa_p2_lane_order(int64_t n_groups, const int64_t *d_group_ptr, const int32_t *d_hdr6, int window, int64_t *d_perm,
                      nfa_stream stream);
int nfa_bplan_has_p2_items(const nfa_bplan *bp);
/* Work-balanced chunk ranges: CTA b of a grid of n_cta CTAs (one per SM) walks chunks d_cta_ptr[b] .. d_cta_ptr[b+1]
 * (int32 [n_cta + 1], non-decreasing, from 0 to n_chunks) instead of an equal COUNT of chunks -- chunks hold equal
 * numbers of CSR blocks, but rows on the domain boundary have fewer cells to walk.  Used when the launch has exactly
 * n_cta CTAs; n_cta = 0 removes the split. */
int nfa_bplan_set_p2_split(nfa_bplan *bp, int n_cta, const int32_t *d_cta_ptr, nfa_stream stream);
/* Optional cell mask of the row-item kernel (uint8 per cell of the plan, device pointer; NULL clears it): cells with
 * mask 0 contribute exact zeros.  The sharded assembler switches off the halo cells of a rank's local plan this way,
 * so that mesh-constant lambda, mu stay single numbers (stride 0) on every rank. */
int nfa_bplan_set_cell_mask(nfa_bplan *bp, const uint8_t *d_mask, nfa_stream stream);
int nfa_assemble_p2tet_elasticity(const nfa_bplan *bp, const double *d_coords, const int32_t *d_cell_nodes,
                                  const double *d_lambda, int64_t lambda_stride, const double *d_mu,
                                  int64_t mu_stride, double *d_vals, nfa_stream stream);
/* Matrix AND load vector in the same pass: the items that own a node's diagonal block walk exactly the cells the node's
 * right-hand-side rows need, so they also sum the cells' local load vectors d_re4 [n_cells, 10, 4] (from
 * nfa_local_h1_load_pitched, comp_pitch = 4) in ascending cell order -- the order of res_g[dest] += r_el,
 * R/src/assembly/ScatterFormData.py:14-18 -- and write rows 3 N .. 3 N + 2 of d_rhs.  rhs_overwrite = 1: the sums
 * replace the rows (d_rhs is zero on entry, or every node of the plan has cells); 0: they are added to what d_rhs holds. */
int nfa_assemble_p2tet_elasticity_load(const nfa_bplan *bp, const double *d_coords, const int32_t *d_cell_nodes,
                                       const double *d_lambda, int64_t lambda_stride, const double *d_mu,
                                       int64_t mu_stride, double *d_vals, const double *d_re4, double *d_rhs,
                                       int rhs_overwrite, nfa_stream stream);

/* ---- penalised normal trace of a row-wise H(div) field on boundary facets: LCEDualWeakFormBCNeumann
 * (R/src/weak_forms/lce_dual_weak_form.py:759-871; the same class in le_dual_weak_form.py:581-644 and
 * lce_scaled_dual_weak_form.py:814-917).  Same facet inputs as nfa_local_hdiv_facet_vec; writes the field's block
 * J[(k,c),(l,c)] = beta int (psi_k.n)(psi_l.n) at [off, off + nfd * n_comp) of the facet's n_loc x n_loc local matrix
 * (zeroed by the caller) and r[(k,c)] = -beta int g_c (psi_k.n). */
int nfa_local_hdiv_facet_penalty(int dim, int nfd, int n_q, int n_comp, const double *d_trace_tab,
                                 const int32_t *d_tidx, const double *d_weights, const double *d_Tcases,
                                 const int8_t *d_cases, const double *d_coords, const int32_t *d_facet_nodes,
                                 const int32_t *d_neigh_nodes, const double *d_normals, const double *d_g,
                                 int64_t n_facets, double beta, int off, int n_loc, double *d_Ke, double *d_re,
                                 nfa_stream stream);

/* ---- generic multi-field Lagrange volume form: the local kernel of LCEPrimalWeakForm
 * (R/src/weak_forms/lce_primal_weak_form.py:9-220), LETaylorHoodWeakForm (R/src/weak_forms/le_taylor_hood_weak_form.py:9-186)
 * and LEPrimalStressConstraintWeakForm (R/src/weak_forms/le_primal_stress_constraint_weak_form.py:10-126).
 * Local dof i = (basis dof_a[i] of the concatenated table d_tab [1 + dim][n_q][n_phi], component row dof_cr[i]);
 * J[i][j] = sum_q w det sum_{terms of (cr_i, cr_j)} val * coef_m(q) * D^s phi_a(q) * D^t phi_b(q) with terms packed as
 * s | t << 4 | m << 8 in d_term_code[d_term_ptr[cr_i * n_crow + cr_j] ..), D^0 = value, D^{1 + d} = d/dx_d;
 * r = J alpha[dest] + sum_q w det sum_l load[cell][q][l] D^{s_l} phi_a (rows with cr_l = cr_i; d_load_code = cr | s << 8).
 * coef_ptrs / strides are HOST arrays of n_coef (<= 8) device pointers, cell strides and point strides. */
int nfa_local_h1_multifield(int dim, int n_q, int n_phi, int n_loc, int n_crow, const double *d_tab,
                            const double *d_weights, const int32_t *d_dof_a, const int32_t *d_dof_cr,
                            const int32_t *d_term_ptr, const int32_t *d_term_code, const double *d_term_val,
                            int n_coef, const double *const *coef_ptrs, const int64_t *coef_cell_stride,
                            const int64_t *coef_point_stride, int n_load, const int32_t *d_load_code,
                            const double *d_load, const double *d_coords, const int32_t *d_cell_nodes,
                            int64_t n_cells, const double *d_alpha, const int32_t *d_dest, double *d_Ke, double *d_re,
                            nfa_stream stream);

/* ---- mixed-dimensional Robin coupling across a fracture cell: CouplingWeakForm.evaluate_form
 * (R/examples/md_elliptic_scalar/CouplingWeakForm.py:11-151), scattered by the drivers' interface closure
 * (R/examples/md_elliptic_scalar/md_elliptic_scalar.py:284-342; library types R/src/assembly/ScatterFormData.py:59-76).
 * Per interface cell: the boundary facet of the higher-dimensional domain on its + side (arrays *_a) and on its - side
 * (*_b), each described like the facets of nfa_local_hdiv_facet (trace table index, orientation case, facet / neighbour
 * vertices, outward normal of the neighbour); d_p_val [n_q][n_p] the fracture's potential basis at the facet points;
 * d_ainv [n_items][n_q] = (delta / 2) / kappa_normal.  Local dofs [u_+ (nfd) | u_- (nfd) | p (n_p)]; writes
 * Ke [n_items][n_loc][n_loc] (zero blocks stored) and re = Ke alpha_l (0 without a state vector). */
int nfa_local_hdiv_coupling(int dim, int nfd, int n_p, int n_q, const double *d_trace_tab, const double *d_weights,
                            const double *d_Tcases, const double *d_coords, const int32_t *d_tidx_a,
                            const int8_t *d_cases_a, const int32_t *d_facet_nodes_a, const int32_t *d_neigh_nodes_a,
                            const double *d_normals_a, const int32_t *d_tidx_b, const int8_t *d_cases_b,
                            const int32_t *d_facet_nodes_b, const int32_t *d_neigh_nodes_b, const double *d_normals_b,
                            const double *d_p_val, const double *d_ainv, int64_t n_items, const double *d_alpha,
                            const int32_t *d_dest, double *d_Ke, double *d_re, nfa_stream stream);

/* ---- H(curl) fields (first-kind Nedelec "N1E", covariant Piola map phi = J^{-T} phi_hat) of L2ProjectorWeakForm
 * (R/src/weak_forms/l2_projector_weak_form.py:8-93; the family reaches it through FiniteElement.evaluate_basis,
 * R/src/basis/finite_element.py:156-186, and permute_and_transform, R/src/basis/permute_and_transform.py:8-54) and of
 * l2_error (R/src/postprocess/l2_error_post_processor.py:7-63).  d_val [n_q][n_phi][dim] reference values, d_signs int8
 * [n_cells][n_phi] = the edge-orientation factor of every dof, d_f / d_exact [n_cells][n_q][3] host-evaluated data.
 *   Ke[a][b] = sum_q w det phi_a . phi_b,  re[a] = r_sign sum_q w det f . phi_a;   out[cell] = int |exact - u_h|^2. */
int nfa_local_hcurl_mass(int dim, int n_phi, int n_q, const double *d_val, const double *d_weights, const int8_t *d_signs,
                         const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_f,
                         double r_sign, double *d_Ke, double *d_re, nfa_stream stream);
int nfa_hcurl_field_error(int dim, int n_phi, int n_q, const double *d_val, const double *d_weights, const int8_t *d_signs,
                          const double *d_coords, const int32_t *d_cell_nodes, int64_t n_cells, const double *d_alpha,
                          const int32_t *d_dest, int dest_stride, int dest_off, const double *d_exact, double *d_out,
                          nfa_stream stream);

/* ---- interface-row exchange over NVLink peer memory (multi-GPU extension; the reference is one process).
 * nfa_p2p_alloc: cudaMalloc'ed, zeroed area + its 64-byte CUDA IPC handle; nfa_p2p_open maps a peer's area into this
 * process (peer access enabled lazily); nfa_p2p_close / nfa_p2p_free release them.
 * nfa_push_rows: the CSR segments of `d_rows` (d_buf_off int64 [n_rows + 1] prefix sums of their lengths) and their
 * right-hand-side entries are stored into the OWNER's area (d_remote_buf: [segments | rhs entries]) once the owner has
 * acknowledged the previous payload (d_local_ack, written by the owner); then d_remote_flag receives the epoch.
 * nfa_pull_add_rows: waits for the flag of the next epoch in its own area, adds the payload to vals / rhs and publishes
 * the acknowledgement at the sender (d_remote_ack).  d_state: uint64 [2] per (peer, direction) in LOCAL memory, zeroed
 * once.  All stream-ordered device work (capturable in a CUDA graph); call the pulls in ascending sender rank. */
int nfa_p2p_alloc(int64_t bytes, void **d_ptr, void *handle64);
int nfa_p2p_open(const void *handle64, void **d_peer_ptr);
int nfa_p2p_close(void *d_peer_ptr);
int nfa_p2p_free(void *d_ptr);
/* number of exchange waits that gave up (a peer never published its epoch); 0 in a healthy run, -1 on a CUDA error */
int64_t nfa_p2p_timeouts(void);
int nfa_push_rows(const int32_t *d_indptr, const double *d_vals, const double *d_rhs, const int32_t *d_rows,
                  const int64_t *d_buf_off, int64_t n_rows, double *d_remote_buf, uint64_t *d_remote_flag,
                  const uint64_t *d_local_ack, uint64_t *d_state, nfa_stream stream);
int nfa_pull_add_rows(const int32_t *d_indptr, double *d_vals, double *d_rhs, const int32_t *d_rows,
                      const int64_t *d_buf_off, int64_t n_rows, const double *d_local_buf, const uint64_t *d_local_flag,
                      uint64_t *d_remote_ack, uint64_t *d_state, nfa_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* NFA_H */
