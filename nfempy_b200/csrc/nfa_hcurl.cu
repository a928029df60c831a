// libnfa: H(curl) fields (first-kind Nedelec, "N1E") of the L2 projector and of the error norms (sm_100a).
//
// The reference pushes N1E basis functions forward with the covariant Piola map (basix push_forward inside
// FiniteElement.evaluate_basis, R/src/basis/finite_element.py:156-186) and fixes the edge orientation with
// permute_and_transform (R/src/basis/permute_and_transform.py:8-19, 38-54: an edge whose local vertices are not in
// ascending global order gets transformations["interval"][0], which is [-1] for the lowest degree); its only consumer is
// the projector (R/src/weak_forms/l2_projector_weak_form.py:8-93, driven by R/tests/test_projectors/test_projectors.py:
// 187-216).  Here:  phi_a(x_q) = s_a J^{-T} phi_hat_a(xi_q),  s_a = +-1 per (cell, edge dof) from the host,
//   J[a][b] = sum_q w_q det phi_a . phi_b,     r[a] = sign sum_q w_q det f(x_q) . phi_a.
// One warp per cell; a lane owns entries of the local matrix.
#include "nfa_plan.cuh"

using namespace nfa;

namespace {

template <int DIM>
__device__ __forceinline__ void covariant(const Geom<DIM> &g, const double *ref, double s, double (&out)[3]) {
#pragma unroll
    for (int x = 0; x < 3; ++x) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) v += g.Jinv[k][x] * ref[k];
        out[x] = s * v;
    }
}

template <int DIM>
__global__ void __launch_bounds__(128)
    local_hcurl_mass_kernel(int n_phi, int n_q, const double *__restrict__ val, const double *__restrict__ weights,
                            const int8_t *__restrict__ signs, const double *__restrict__ coords,
                            const int32_t *__restrict__ cell_nodes, int64_t n_cells, const double *__restrict__ f,
                            double r_sign, double *__restrict__ Ke, double *__restrict__ re) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t cell = (int64_t)blockIdx.x * 4 + warp; cell < n_cells; cell += (int64_t)gridDim.x * 4) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const int8_t *sg = signs + cell * n_phi;
        for (int e = lane; e < n_phi * n_phi + n_phi; e += 32) {
            double acc = 0.0;
            if (e < n_phi * n_phi) {
                const int a = e / n_phi, b = e - a * n_phi;
                for (int q = 0; q < n_q; ++q) {
                    double pa[3], pb[3];
                    covariant<DIM>(g, val + ((int64_t)q * n_phi + a) * DIM, (double)sg[a], pa);
                    covariant<DIM>(g, val + ((int64_t)q * n_phi + b) * DIM, (double)sg[b], pb);
                    acc += weights[q] * g.det * (pa[0] * pb[0] + pa[1] * pb[1] + pa[2] * pb[2]);
                }
                if (Ke) Ke[cell * n_phi * n_phi + e] = acc;
            } else {
                const int a = e - n_phi * n_phi;
                if (f) {
                    for (int q = 0; q < n_q; ++q) {
                        double pa[3];
                        covariant<DIM>(g, val + ((int64_t)q * n_phi + a) * DIM, (double)sg[a], pa);
                        const double *fq = f + ((int64_t)cell * n_q + q) * 3;
                        acc += weights[q] * g.det * (pa[0] * fq[0] + pa[1] * fq[1] + pa[2] * fq[2]);
                    }
                }
                re[cell * n_phi + a] = r_sign * acc;
            }
        }
    }
}

// out[cell] = int_cell |f_exact - f_h|^2,  f_h = sum_a alpha[dest[a]] phi_a  (l2_error of an H(curl) field)
template <int DIM>
__global__ void __launch_bounds__(128)
    hcurl_field_error_kernel(int n_phi, int n_q, const double *__restrict__ val, const double *__restrict__ weights,
                             const int8_t *__restrict__ signs, const double *__restrict__ coords,
                             const int32_t *__restrict__ cell_nodes, int64_t n_cells, const double *__restrict__ alpha,
                             const int32_t *__restrict__ dest, int dest_stride, int dest_off,
                             const double *__restrict__ exact, double *__restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t cell = (int64_t)blockIdx.x * 4 + warp; cell < n_cells; cell += (int64_t)gridDim.x * 4) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const int8_t *sg = signs + cell * n_phi;
        const int32_t *dd = dest + cell * dest_stride + dest_off;
        double err = 0.0;
        for (int q = lane; q < n_q; q += 32) {
            double uh[3] = {0.0, 0.0, 0.0};
            for (int a = 0; a < n_phi; ++a) {
                double pa[3];
                covariant<DIM>(g, val + ((int64_t)q * n_phi + a) * DIM, (double)sg[a], pa);
                const double al = alpha[dd[a]];
                uh[0] += al * pa[0];
                uh[1] += al * pa[1];
                uh[2] += al * pa[2];
            }
            const double *ex = exact + ((int64_t)cell * n_q + q) * 3;
            double loc = 0.0;
#pragma unroll
            for (int x = 0; x < 3; ++x) loc += (ex[x] - uh[x]) * (ex[x] - uh[x]);
            err += weights[q] * g.det * loc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if (lane == 0) out[cell] = err;
    }
}

}  // namespace

extern "C" int nfa_local_hcurl_mass(int dim, int n_phi, int n_q, const double *val, const double *weights,
                                    const int8_t *signs, const double *coords, const int32_t *cell_nodes,
                                    int64_t n_cells, const double *f, double r_sign, double *Ke, double *re,
                                    nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!val || !weights || !signs || !re) return fail("nfa_local_hcurl_mass: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ctas = (n_cells + 3) / 4;
    const int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (dim == 3)
        local_hcurl_mass_kernel<3><<<grid, 128, 0, s>>>(n_phi, n_q, val, weights, signs, coords, cell_nodes, n_cells, f, r_sign, Ke, re);
    else if (dim == 2)
        local_hcurl_mass_kernel<2><<<grid, 128, 0, s>>>(n_phi, n_q, val, weights, signs, coords, cell_nodes, n_cells, f, r_sign, Ke, re);
    else
        return fail("nfa_local_hcurl_mass: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_hcurl_field_error(int dim, int n_phi, int n_q, const double *val, const double *weights,
                                     const int8_t *signs, const double *coords, const int32_t *cell_nodes,
                                     int64_t n_cells, const double *alpha, const int32_t *dest, int dest_stride,
                                     int dest_off, const double *exact, double *out, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!val || !weights || !signs || !alpha || !dest || !exact || !out) return fail("nfa_hcurl_field_error: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ctas = (n_cells + 3) / 4;
    const int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (dim == 3)
        hcurl_field_error_kernel<3><<<grid, 128, 0, s>>>(n_phi, n_q, val, weights, signs, coords, cell_nodes, n_cells, alpha, dest, dest_stride, dest_off, exact, out);
    else if (dim == 2)
        hcurl_field_error_kernel<2><<<grid, 128, 0, s>>>(n_phi, n_q, val, weights, signs, coords, cell_nodes, n_cells, alpha, dest, dest_stride, dest_off, exact, out);
    else
        return fail("nfa_hcurl_field_error: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}
