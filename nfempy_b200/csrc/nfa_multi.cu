// libnfa: generic multi-field H1 (Lagrange) volume form -- the local kernel behind the reference's remaining primal
// forms: LCEPrimalWeakForm (R/src/weak_forms/lce_primal_weak_form.py:9-220), LETaylorHoodWeakForm
// (R/src/weak_forms/le_taylor_hood_weak_form.py:9-186) and LEPrimalStressConstraintWeakForm
// (R/src/weak_forms/le_primal_stress_constraint_weak_form.py:10-126).
//
// All of them have the shape the reference writes as ``grad_phi @ S.T`` / ``phi @ s``: the residual row of local dof
// (field f, basis a, component c) is  sum_q w det sum_s D^s phi_a(x_q) * Sigma[(f, c), s](x_q),  with D^0 = value,
// D^{1..dim} = physical gradient, and the generalised stress Sigma linear in the generalised strain
// E[(f', c'), t] = sum_b alpha_{f', b, c'} D^t phi_b:   Sigma = sum_m coef_m(x) B_m E  (+ a load term).
// The constant matrices B_m are sparse; the host lists their non-zeros per pair of component rows (cr, cr') as terms
// (s, t, m, value), so an entry of the local matrix costs a handful of DFMA per quadrature point.
// Local dof order: field-major, basis-major, component-minor (the reference's ``alpha[:, c:n:ncomp]`` slices).
#include <algorithm>

#include "nfa_common.cuh"

using namespace nfa;

namespace {

constexpr int MF_MAX_COEF = 8;

struct MfCoefs {
    const double *p[MF_MAX_COEF];
    int64_t sc[MF_MAX_COEF], sq[MF_MAX_COEF];
};

// One CTA per cell.
template <int DIM>
__global__ void __launch_bounds__(128)
    local_h1_multifield_kernel(int n_q, int n_phi, int n_loc, int n_crow, const double *__restrict__ tab,
                               const double *__restrict__ weights, const int32_t *__restrict__ dof_a,
                               const int32_t *__restrict__ dof_cr, const int32_t *__restrict__ term_ptr,
                               const int32_t *__restrict__ term_code, const double *__restrict__ term_val, int n_coef,
                               MfCoefs coefs, int n_load, const int32_t *__restrict__ load_code,
                               const double *__restrict__ load, const double *__restrict__ coords,
                               const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                               const double *__restrict__ alpha, const int32_t *__restrict__ dest,
                               double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NOP = DIM + 1;
    extern __shared__ double smem[];
    double *s_d = smem;                          // [n_q][n_phi][NOP] value and physical gradient of every basis function
    double *s_cw = s_d + n_q * n_phi * NOP;      // [n_q][n_coef] w det coef_m
    double *s_wd = s_cw + n_q * n_coef;          // [n_q] w det
    double *s_a = s_wd + n_q;                    // [n_loc] alpha (local)
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        __syncthreads();  // the previous cell's readers are done with shared memory
        for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
            const double wd = weights[q] * g.det;
            s_wd[q] = wd;
            for (int m = 0; m < n_coef; ++m) s_cw[q * n_coef + m] = wd * coefs.p[m][cell * coefs.sc[m] + q * coefs.sq[m]];
        }
        for (int i = threadIdx.x; i < n_loc; i += blockDim.x) s_a[i] = (alpha && dest) ? alpha[dest[cell * n_loc + i]] : 0.0;
        for (int idx = threadIdx.x; idx < n_q * n_phi; idx += blockDim.x) {
            const int q = idx / n_phi, a = idx - q * n_phi;
            double *d = s_d + (size_t)idx * NOP;
            d[0] = tab[(size_t)q * n_phi + a];
            double gh[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) gh[k] = tab[((size_t)(1 + k) * n_q + q) * n_phi + a];
#pragma unroll
            for (int x = 0; x < DIM; ++x) {  // physical gradient: d/dx_x = sum_k Jinv[k][x] d/dxi_k
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) v += g.Jinv[k][x] * gh[k];
                d[1 + x] = v;
            }
        }
        __syncthreads();
        double *K = Ke + cell * (int64_t)n_loc * n_loc;
        for (int e = threadIdx.x; e < n_loc * n_loc; e += blockDim.x) {
            const int i = e / n_loc, j = e - i * n_loc;
            const int a = dof_a[i], b = dof_a[j], pair = dof_cr[i] * n_crow + dof_cr[j];
            const int t0 = term_ptr[pair], t1 = term_ptr[pair + 1];
            double acc = 0.0;
            if (t1 > t0) {
                for (int q = 0; q < n_q; ++q) {
                    const double *da = s_d + ((size_t)q * n_phi + a) * NOP, *db = s_d + ((size_t)q * n_phi + b) * NOP;
                    const double *cw = s_cw + q * n_coef;
                    double sq = 0.0;
                    for (int t = t0; t < t1; ++t) {
                        const int code = term_code[t];  // s | t << 4 | m << 8
                        sq = fma(term_val[t] * cw[code >> 8], da[code & 15] * db[(code >> 4) & 15], sq);
                    }
                    acc += sq;
                }
            }
            K[e] = acc;
        }
        __syncthreads();  // K of this cell is complete in global memory for the row products below
        for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
            double acc = 0.0;
            if (alpha && dest)
                for (int j = 0; j < n_loc; ++j) acc += K[i * n_loc + j] * s_a[j];
            if (load) {
                const int a = dof_a[i], cr = dof_cr[i];
                for (int l = 0; l < n_load; ++l) {
                    const int code = load_code[l];  // cr | s << 8
                    if ((code & 255) != cr) continue;
                    const int s = code >> 8;
                    for (int q = 0; q < n_q; ++q)
                        acc += s_wd[q] * load[((size_t)cell * n_q + q) * n_load + l] * s_d[((size_t)q * n_phi + a) * NOP + s];
                }
            }
            re[cell * n_loc + i] = acc;
        }
    }
}

}  // namespace

extern "C" int nfa_local_h1_multifield(int dim, int n_q, int n_phi, int n_loc, int n_crow, const double *tab,
                                       const double *weights, const int32_t *dof_a, const int32_t *dof_cr,
                                       const int32_t *term_ptr, const int32_t *term_code, const double *term_val,
                                       int n_coef, const double *const *coef_ptrs, const int64_t *coef_cell_stride,
                                       const int64_t *coef_point_stride, int n_load, const int32_t *load_code,
                                       const double *load, const double *coords, const int32_t *cell_nodes,
                                       int64_t n_cells, const double *alpha, const int32_t *dest, double *Ke,
                                       double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (n_coef < 1 || n_coef > MF_MAX_COEF) return fail("nfa_local_h1_multifield: between 1 and 8 coefficient arrays");
    if (n_crow < 1 || n_crow > 255 || dim + 1 > 15) return fail("nfa_local_h1_multifield: too many component rows");
    if (n_load > 0 && (!load || !load_code)) return fail("nfa_local_h1_multifield: load arrays missing");
    if (alpha && !dest) return fail("nfa_local_h1_multifield: alpha given without dest");
    cudaStream_t s = (cudaStream_t)stream;
    MfCoefs cf;
    for (int m = 0; m < MF_MAX_COEF; ++m) {
        cf.p[m] = m < n_coef ? coef_ptrs[m] : nullptr;
        cf.sc[m] = m < n_coef ? coef_cell_stride[m] : 0;
        cf.sq[m] = m < n_coef ? coef_point_stride[m] : 0;
        if (m < n_coef && !cf.p[m]) return fail("nfa_local_h1_multifield: NULL coefficient array");
    }
    const size_t smem = sizeof(double) * ((size_t)n_q * n_phi * (dim + 1) + (size_t)n_q * n_coef + n_q + n_loc);
    if (smem > 200 * 1024) return fail("nfa_local_h1_multifield: element too large for the shared-memory staging");
    const int grid = (int)std::min<int64_t>(n_cells, 148 * 8);
    const double *ld = n_load > 0 ? load : nullptr;
#define NFA_MF(D)                                                                                                   \
    {                                                                                                               \
        auto kern = local_h1_multifield_kernel<D>;                                                                  \
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<grid, 128, smem, s>>>(n_q, n_phi, n_loc, n_crow, tab, weights, dof_a, dof_cr, term_ptr, term_code,   \
                                     term_val, n_coef, cf, n_load, load_code, ld, coords, cell_nodes, n_cells,      \
                                     alpha, dest, Ke, re);                                                          \
    }
    if (dim == 3)
        NFA_MF(3)
    else if (dim == 2)
        NFA_MF(2)
    else if (dim == 1)
        NFA_MF(1)
    else
        return fail("nfa_local_h1_multifield: dim must be 1, 2 or 3");
#undef NFA_MF
    NFA_LAUNCH_CHECK();
    return 0;
}
