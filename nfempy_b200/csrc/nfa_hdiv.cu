// libnfa: H(div) x L2 local kernels (K3): mixed Poisson volume form and its natural (Dirichlet) boundary load.
// Reference: LaplaceDualWeakForm / LaplaceDualWeakFormBCDirichlet, R/src/weak_forms/laplace_dual_weak_form.py:10-154.
// Basis handling restates FiniteElement.evaluate_basis (R/src/basis/finite_element.py:156-186): contravariant Piola
// (1/|det|) J phi_hat on the values, divergence = tr(grad_hat phi_hat)/|det|, then the orientation transformation of
// permute_and_transform (R/src/basis/permute_and_transform.py:110-140) as a small matrix per facet chosen by a
// per-cell case id computed on the host.
#include "nfa_common.cuh"

using namespace nfa;

namespace {

struct CoefH {
    const double *p;
    int64_t sc, sq;
};

// One CTA per cell.
template <int DIM>
__global__ void __launch_bounds__(128)
    local_hdiv_l2_kernel(int n_u, int n_p, int n_q, int nfd, const double *__restrict__ u_val,
                         const double *__restrict__ u_div, const double *__restrict__ p_val,
                         const double *__restrict__ weights, const double *__restrict__ Tcases,
                         const int8_t *__restrict__ cases, const double *__restrict__ coords,
                         const int32_t *__restrict__ cell_nodes, int64_t n_cells, CoefH kappa,
                         const double *__restrict__ f, double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    extern __shared__ double smem[];
    double *s_psi = smem;                         // [n_q][n_u][DIM]
    double *s_div = s_psi + n_q * n_u * DIM;      // [n_q][n_u]
    double *s_wk = s_div + n_q * n_u;             // [n_q] w det / kappa
    double *s_wd = s_wk + n_q;                    // [n_q] w det
    double *s_T = s_wd + n_q;                     // [NF][nfd][nfd]
    const int n_loc = n_u + n_p;
    const int n_fdofs = NF * nfd;
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const double idet = 1.0 / g.det;
        __syncthreads();  // previous cell's readers are done with shared memory
        for (int i = threadIdx.x; i < NF * nfd * nfd; i += blockDim.x) {
            int fc = i / (nfd * nfd);
            s_T[i] = Tcases[(int)cases[cell * NF + fc] * nfd * nfd + (i - fc * nfd * nfd)];
        }
        for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
            double wd = weights[q] * g.det;
            s_wd[q] = wd;
            s_wk[q] = wd * (1.0 / kappa.p[cell * kappa.sc + q * kappa.sq]);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < n_q * n_u; idx += blockDim.x) {
            int q = idx / n_u, j = idx - q * n_u;
            double vh[DIM], dh;
            if (j < n_fdofs) {
                int fc = j / nfd, a = j - fc * nfd;
#pragma unroll
                for (int k = 0; k < DIM; ++k) vh[k] = 0.0;
                dh = 0.0;
                for (int b = 0; b < nfd; ++b) {
                    double t = s_T[(fc * nfd + a) * nfd + b];
                    const double *src = u_val + ((int64_t)q * n_u + fc * nfd + b) * DIM;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) vh[k] += t * src[k];
                    dh += t * u_div[q * n_u + fc * nfd + b];
                }
            } else {
#pragma unroll
                for (int k = 0; k < DIM; ++k) vh[k] = u_val[((int64_t)q * n_u + j) * DIM + k];
                dh = u_div[q * n_u + j];
            }
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) v += g.J[c][k] * vh[k];
                s_psi[(q * n_u + j) * DIM + c] = v * idet;
            }
            s_div[q * n_u + j] = dh * idet;
        }
        __syncthreads();
        double *K = Ke + cell * (int64_t)n_loc * n_loc;
        for (int e = threadIdx.x; e < n_loc * n_loc; e += blockDim.x) {
            int ra = e / n_loc, rb = e - ra * n_loc;
            double acc = 0.0;
            if (ra < n_u && rb < n_u) {
                for (int q = 0; q < n_q; ++q) {
                    const double *a = s_psi + (q * n_u + ra) * DIM, *b = s_psi + (q * n_u + rb) * DIM;
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) dot += a[c] * b[c];
                    acc += s_wk[q] * dot;
                }
            } else if (ra < n_u) {
                for (int q = 0; q < n_q; ++q) acc -= s_wd[q] * p_val[q * n_p + (rb - n_u)] * s_div[q * n_u + ra];
            } else if (rb < n_u) {
                for (int q = 0; q < n_q; ++q) acc += s_wd[q] * s_div[q * n_u + rb] * p_val[q * n_p + (ra - n_u)];
            }
            K[e] = acc;
        }
        for (int ra = threadIdx.x; ra < n_loc; ra += blockDim.x) {
            double acc = 0.0;
            if (ra >= n_u && f)
                for (int q = 0; q < n_q; ++q) acc -= s_wd[q] * p_val[q * n_p + (ra - n_u)] * f[cell * n_q + q];
            re[cell * n_loc + ra] = acc;
        }
    }
}

// Natural boundary load: r_k = sum_q det_F w_q p_D(x_q) (psi_k(x_q) . n), psi = trace of the neighbour's basis.
template <int DIM>
__global__ void __launch_bounds__(128)
    local_hdiv_facet_kernel(int nfd, int n_q, const double *__restrict__ trace_tab, const int32_t *__restrict__ tidx,
                            const double *__restrict__ weights, const double *__restrict__ Tcases,
                            const int8_t *__restrict__ cases, const double *__restrict__ coords,
                            const int32_t *__restrict__ facet_nodes, const int32_t *__restrict__ neigh_nodes,
                            const double *__restrict__ normals, const double *__restrict__ pd, int64_t n_facets,
                            double *__restrict__ re) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t fc = (int64_t)blockIdx.x * 4 + warp; fc < n_facets; fc += (int64_t)gridDim.x * 4) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, neigh_nodes, fc, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        double det_f = 1.0;
        if constexpr (DIM > 1) {
            double XF[DIM][3];
            load_cell_coords<DIM - 1>(coords, facet_nodes, fc, XF);
            Geom<DIM - 1> gf;
            compute_geom<DIM - 1>(XF, gf);
            det_f = gf.det;
        }
        const double *T = Tcases + (int)cases[fc] * nfd * nfd;
        const double *tab = trace_tab + (int64_t)tidx[fc] * n_q * nfd * DIM;
        double jn[DIM];  // (J^T n) / det: psi.n = (1/det) n^T J psi_hat
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) v += g.J[c][k] * normals[fc * 3 + c];
            jn[k] = v / g.det;
        }
        for (int k = lane; k < nfd; k += 32) {
            double acc = 0.0;
            for (int q = 0; q < n_q; ++q) {
                double flux = 0.0;
                for (int b = 0; b < nfd; ++b) {
                    const double *src = tab + ((int64_t)q * nfd + b) * DIM;
                    double t = 0.0;
#pragma unroll
                    for (int kk = 0; kk < DIM; ++kk) t += jn[kk] * src[kk];
                    flux += T[k * nfd + b] * t;
                }
                acc += det_f * weights[q] * pd[fc * n_q + q] * flux;
            }
            re[fc * nfd + k] = acc;
        }
    }
}

}  // namespace

extern "C" int nfa_local_hdiv_l2(int dim, int n_u, int n_p, int n_q, int nfd, const double *u_val,
                                 const double *u_div, const double *p_val, const double *weights,
                                 const double *Tcases, const int8_t *cases, const double *coords,
                                 const int32_t *cell_nodes, int64_t n_cells, const double *kappa, int64_t ksc,
                                 int64_t ksq, const double *f, double *Ke, double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if ((dim + 1) * nfd > n_u) return fail("nfa_local_hdiv_l2: facet dofs exceed the basis size");
    cudaStream_t s = (cudaStream_t)stream;
    CoefH kp{kappa, ksc, ksq};
    size_t smem = sizeof(double) * ((size_t)n_q * n_u * (dim + 1) + 2 * n_q + (size_t)(dim + 1) * nfd * nfd);
    if (smem > 200 * 1024) return fail("nfa_local_hdiv_l2: element too large for the shared-memory staging");
    int grid = (int)(n_cells < 148 * 8 ? n_cells : 148 * 8);
    if (dim == 3) {
        auto kern = local_hdiv_l2_kernel<3>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,
                                     cell_nodes, n_cells, kp, f, Ke, re);
    } else if (dim == 2) {
        auto kern = local_hdiv_l2_kernel<2>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,
                                     cell_nodes, n_cells, kp, f, Ke, re);
    } else
        return fail("nfa_local_hdiv_l2: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_hdiv_facet(int dim, int nfd, int n_q, const double *trace_tab, const int32_t *tidx,
                                    const double *weights, const double *Tcases, const int8_t *cases,
                                    const double *coords, const int32_t *facet_nodes, const int32_t *neigh_nodes,
                                    const double *normals, const double *pd, int64_t n_facets, double *re,
                                    nfa_stream stream) {
    if (n_facets == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t ctas = (n_facets + 3) / 4;
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (dim == 3)
        local_hdiv_facet_kernel<3><<<grid, 128, 0, s>>>(nfd, n_q, trace_tab, tidx, weights, Tcases, cases, coords,
                                                        facet_nodes, neigh_nodes, normals, pd, n_facets, re);
    else if (dim == 2)
        local_hdiv_facet_kernel<2><<<grid, 128, 0, s>>>(nfd, n_q, trace_tab, tidx, weights, Tcases, cases, coords,
                                                        facet_nodes, neigh_nodes, normals, pd, n_facets, re);
    else
        return fail("nfa_local_hdiv_facet: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}
