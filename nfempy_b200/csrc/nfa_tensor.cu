// libnfa: local matrices of the mixed "tensor" forms -- row-wise H(div) stresses with vector L2 multipliers (sm_100a).
//
// One kernel covers the block structure shared by the reference's vectorised weakly-symmetric mixed forms
//   LEDualWeakForm.evaluate_form_vectorized          R/src/weak_forms/le_dual_weak_form.py:354-510        (s, u, t)
//   LCEDualWeakForm.evaluate_form_vectorized         R/src/weak_forms/lce_dual_weak_form.py:410-648       (s, m, u, t)
//   LCEScaledDualWeakForm.evaluate_form_vectorized   R/src/weak_forms/lce_scaled_dual_weak_form.py:445-699
//   LERieszMapWeakForm / LCE(Scaled)RieszMapWeakForm .evaluate_form_vectorized (block-diagonal norm matrices)
// Local dofs: [H(div) field 0 | H(div) field 1 | L2 field 0 | L2 field 1], basis-major, component-minor; dof (a, c) of
// an H(div) field is row c of the tensor built from the vector function psi_a (contravariant Piola + orientation).
//   (F, F)   int A_id (psi_a.psi_b) d_cc' + A_tr psi_b[c] psi_a[c'] + A_vol psi_a[c] psi_b[c'] + A_dd div psi_a div psi_b d_cc'
//   (U, F)   d_cc' int w phi_k (c_d div^ psi_a + det g.psi_a)        generalised divergence coupling (c_d = 1, g = 0: plain)
//            + int w det phi_k sum_d eps(c, c', d) psi_a[d]           asymmetry coupling (rotation x stress)
//   (U, U)   d_cc' int A_m phi_k phi_l
//   r_U      -int load_c phi_k  (+ J alpha when a state is given)
// All coefficients are host-evaluated per quadrature point.  One CTA per cell: the mapped bases are staged in shared
// memory, threads stride the n_loc^2 entries.  Output goes to the two-stage reduce (nfa_plan_reduce).
#include <algorithm>
#include <cstdlib>

#include "nfa_common.cuh"

using namespace nfa;

namespace {

struct TensorDev {  // device copy of nfa_tensor_form, pointers already device pointers
    nfa_tensor_form f;
};

__device__ __forceinline__ double levi_civita(int a, int b, int c) {
    if (a == b || b == c || a == c) return 0.0;
    return ((b - a + 3) % 3 == 1) ? 1.0 : -1.0;  // (a, b, c) cyclic -> +1
}

// THREADS: 128 for elements whose staged bases leave room for several CTAs per SM; 512 for the large ones (BDM_3 x RT_3
// stages 125 KB per CTA: one CTA per SM, which with 128 threads would be four warps per SM)
template <int DIM, int THREADS>
__global__ void __launch_bounds__(THREADS)
    local_hdiv_tensor_kernel(const nfa_tensor_form P, int n_qc, const double *__restrict__ coords,
                             const int32_t *__restrict__ cell_nodes, int64_t n_cells, const double *__restrict__ alpha,
                             const int32_t *__restrict__ dest, double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    extern __shared__ double smem[];
    const int n_q_all = P.n_q;
    // the quadrature points are processed in chunks of n_qc (all of them when the staged bases fit in shared memory;
    // BDM_3 x RT_3 with 64 points does not): the first chunk writes the local matrix, the following ones add to it --
    // every entry is owned by the same thread in every chunk
    const int n_q = n_qc;
    // shared-memory layout (sized for one chunk)
    double *s_psi[2], *s_div[2], *s_gd[2];
    double *cur = smem;
    int off_h[2], off_l[2], n_loc = 0;
    for (int h = 0; h < 2; ++h) {
        const int np = h < P.n_hdiv ? P.h_n_phi[h] : 0;
        s_psi[h] = cur;
        cur += (size_t)n_q * np * DIM;  // mapped values [q][a][c]
        // physical divergence [q][a]: only the plain div-div term reads it (Riesz maps); the dual forms stage nothing here
        const bool need_div = h < P.n_hdiv && P.h_a_dd[h] != nullptr && !P.h_dd_general[h];
        s_div[h] = need_div ? cur : nullptr;
        cur += need_div ? (size_t)n_q * np : 0;
        s_gd[h] = cur;
        cur += (size_t)n_q * np;  // generalised divergence x det [q][a]
        off_h[h] = n_loc;
        n_loc += np * (h < P.n_hdiv ? P.h_n_comp[h] : 0);
    }
    for (int l = 0; l < 2; ++l) {
        off_l[l] = n_loc;
        n_loc += l < P.n_l2 ? P.l_n_phi[l] * P.l_n_comp[l] : 0;
    }
    double *s_wd = cur;  // [n_q] w det
    cur += n_q;
    double *s_T = cur;  // orientation matrices of this cell, per H(div) field [NF][nfd][nfd]
    int t_off[2] = {0, 0};
    {
        int t = 0;
        for (int h = 0; h < P.n_hdiv; ++h) {
            t_off[h] = t;
            t += NF * P.h_nfd[h] * P.h_nfd[h];
        }
        cur += t;
    }
    double *s_a = cur;  // [n_loc] local state
    cur += n_loc;
    double *s_cq = cur;  // [2][n_q][4] w det (A_id, A_tr, A_vol, A_dd) of the two H(div) fields at this cell's points
    // the (F, F) blocks are computed per PAIR of vector basis functions (all row components at once) when every stress
    // field has DIM rows; the per-entry loop below then only handles the couplings and the L2 masses
    bool pair_pass = true;
    for (int h = 0; h < P.n_hdiv; ++h) pair_pass = pair_pass && P.h_n_comp[h] == DIM;

    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const double idet = 1.0 / g.det;
        double *K = Ke + cell * (int64_t)n_loc * n_loc;
        for (int q0 = 0; q0 < n_q_all; q0 += n_qc) {
        const int nq = min(n_qc, n_q_all - q0);  // points of this chunk; global point index = q0 + q
        const int64_t cq_base = cell * (int64_t)n_q_all + q0;
        __syncthreads();
        for (int h = 0; h < P.n_hdiv; ++h) {
            const int nfd = P.h_nfd[h];
            for (int i = threadIdx.x; i < NF * nfd * nfd; i += blockDim.x) {
                const int fc = i / (nfd * nfd);
                s_T[t_off[h] + i] = P.h_Tcases[h][(int)P.h_cases[h][cell * NF + fc] * nfd * nfd + (i - fc * nfd * nfd)];
            }
        }
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            const double wd = P.weights[q0 + q] * g.det;
            s_wd[q] = wd;
            for (int h = 0; h < P.n_hdiv; ++h) {
                double *cq = s_cq + ((size_t)h * n_q + q) * 4;
                cq[0] = P.h_a_id[h] ? wd * P.h_a_id[h][cq_base + q] : 0.0;
                cq[1] = P.h_a_tr[h] ? wd * P.h_a_tr[h][cq_base + q] : 0.0;
                cq[2] = P.h_a_vol[h] ? wd * P.h_a_vol[h][cq_base + q] : 0.0;
                cq[3] = P.h_a_dd[h] ? wd * P.h_a_dd[h][cq_base + q] : 0.0;
            }
        }
        for (int i = threadIdx.x; i < n_loc; i += blockDim.x) s_a[i] = (alpha && dest) ? alpha[dest[cell * n_loc + i]] : 0.0;
        __syncthreads();
        for (int h = 0; h < P.n_hdiv; ++h) {
            const int np = P.h_n_phi[h], nfd = P.h_nfd[h], n_fdofs = NF * nfd;
            const double *val = P.h_val[h] + (size_t)q0 * np * DIM, *dv = P.h_div[h] + (size_t)q0 * np;
            const double *cd = P.h_div_cd[h], *gv = P.h_div_g[h];
            for (int idx = threadIdx.x; idx < nq * np; idx += blockDim.x) {
                const int q = idx / np, j = idx - q * np;
                double vh[DIM], dh = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) vh[k] = 0.0;
                if (j < n_fdofs) {
                    const int fc = j / nfd, a = j - fc * nfd;
                    for (int b = 0; b < nfd; ++b) {
                        const double t = s_T[t_off[h] + (fc * nfd + a) * nfd + b];
                        const double *src = val + ((int64_t)q * np + fc * nfd + b) * DIM;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) vh[k] += t * src[k];
                        dh += t * dv[q * np + fc * nfd + b];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < DIM; ++k) vh[k] = val[((int64_t)q * np + j) * DIM + k];
                    dh = dv[q * np + j];
                }
                double gd = (cd ? cd[cq_base + q] : 1.0) * dh;  // c_d div^ psi  (reference divergence: no 1/det)
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    double v = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) v += g.J[c][k] * vh[k];
                    v *= idet;
                    s_psi[h][(q * np + j) * DIM + c] = v;
                    if (gv) gd += g.det * gv[(cq_base + q) * DIM + c] * v;
                }
                if (s_div[h]) s_div[h][q * np + j] = dh * idet;
                s_gd[h][q * np + j] = gd;
            }
        }
        __syncthreads();
        if (pair_pass && q0 == 0) {
            // everything outside the blocks written below (stress-stress of one field, couplings, L2 masses) is structurally
            // zero: one coalesced sweep instead of a decode-and-test per entry
            for (int e = threadIdx.x; e < n_loc * n_loc; e += blockDim.x) K[e] = 0.0;
            __syncthreads();
        }
        if (pair_pass) {
            for (int h = 0; h < P.n_hdiv; ++h) {
                const int np = P.h_n_phi[h], n_pairs = np * (np + 1) / 2;
                const double *dd = (P.h_dd_general[h] || !s_div[h]) ? s_gd[h] : s_div[h];  // (unused without a_dd: cq[3] = 0)
                const double dds = P.h_dd_general[h] ? idet : 1.0;  // s_gd carries det (cd div psi + g.psi)
                const double *cqh = s_cq + (size_t)h * n_q * 4;
                for (int pr = threadIdx.x; pr < n_pairs; pr += blockDim.x) {
                    int a = 0, rem = pr;  // pair (a, b), a <= b, from the linear index over the upper triangle
                    while (rem >= np - a) {
                        rem -= np - a;
                        ++a;
                    }
                    const int b = a + rem;
                    double acc[DIM][DIM];
#pragma unroll
                    for (int r = 0; r < DIM; ++r)
#pragma unroll
                        for (int c = 0; c < DIM; ++c) acc[r][c] = 0.0;
                    for (int q = 0; q < nq; ++q) {
                        const double *pa = s_psi[h] + (q * np + a) * DIM, *pb = s_psi[h] + (q * np + b) * DIM;
                        const double *cq = cqh + q * 4;
                        double va[DIM], vb[DIM], dot = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) {
                            va[d] = pa[d];
                            vb[d] = pb[d];
                            dot = fma(va[d], vb[d], dot);
                        }
                        const double diag = cq[0] * dot + cq[3] * (dd[q * np + a] * dds) * (dd[q * np + b] * dds);
                        const double ctr = cq[1], cvol = cq[2];
#pragma unroll
                        for (int r = 0; r < DIM; ++r)
#pragma unroll
                            for (int c = 0; c < DIM; ++c) {
                                // row (a, r), column (b, c): A_tr psi_b[r] psi_a[c] + A_vol psi_a[r] psi_b[c]
                                double t = ctr * vb[r] * va[c] + cvol * va[r] * vb[c];
                                if (r == c) t += diag;
                                acc[r][c] += t;
                            }
                    }
                    const int ra0 = off_h[h] + a * DIM, rb0 = off_h[h] + b * DIM;
#pragma unroll
                    for (int r = 0; r < DIM; ++r)
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            const double v = q0 == 0 ? acc[r][c] : K[(ra0 + r) * n_loc + rb0 + c] + acc[r][c];
                            K[(ra0 + r) * n_loc + rb0 + c] = v;
                            if (a != b) K[(rb0 + c) * n_loc + ra0 + r] = v;  // the block is symmetric
                        }
                }
            }
        }
        if (pair_pass) {
            // L2 x H(div) couplings, one (L2 basis k, H(div) basis a) product per thread: the divergence coupling is one
            // number for the DIM equal-component entries, the asymmetry coupling one vector B_d = int w det phi_k psi_a[d]
            // placed with the Levi-Civita signs; both blocks (and their transposes) are written from registers
            for (int l = 0; l < P.n_l2; ++l)
                for (int h = 0; h < P.n_hdiv; ++h) {
                    const bool dv = P.div_coupling[l][h] != 0, as = P.asym_coupling[l][h] != 0;
                    if (!dv && !as) continue;
                    const int npl = P.l_n_phi[l], nph = P.h_n_phi[h], ncl = P.l_n_comp[l];
                    const double *lv = P.l_val[l] + (size_t)q0 * npl;
                    for (int it = threadIdx.x; it < npl * nph; it += blockDim.x) {
                        const int k = it / nph, a = it - k * nph;
                        double cdiv = 0.0, B[DIM];
#pragma unroll
                        for (int d = 0; d < DIM; ++d) B[d] = 0.0;
                        for (int q = 0; q < nq; ++q) {
                            const double lvq = lv[q * npl + k];
                            if (dv) cdiv += P.weights[q0 + q] * lvq * s_gd[h][q * nph + a];
                            if (as) {
                                const double wl = s_wd[q] * lvq;
#pragma unroll
                                for (int d = 0; d < DIM; ++d) B[d] += wl * s_psi[h][(q * nph + a) * DIM + d];
                            }
                        }
                        for (int lc = 0; lc < ncl; ++lc)
                            for (int hc = 0; hc < DIM; ++hc) {
                                double v = 0.0;
                                bool any = false;
                                if (dv && lc == hc) {
                                    v = cdiv;
                                    any = true;
                                }
                                if (as) {
                                    const int ce = ncl == 1 ? 2 : lc;  // the single rotation of a 2-D problem is the z one
                                    double t = 0.0;
                                    for (int d = 0; d < DIM; ++d) {
                                        const double eps = levi_civita(ce, hc, d);
                                        if (eps != 0.0) t += eps * B[d];
                                    }
                                    v += t;
                                    any = true;
                                }
                                if (!any) continue;  // stays the structural zero
                                const int rl = off_l[l] + k * ncl + lc, ch = off_h[h] + a * DIM + hc;
                                if (q0 == 0) {
                                    K[rl * n_loc + ch] = v;
                                    K[ch * n_loc + rl] = v;
                                } else {
                                    K[rl * n_loc + ch] += v;
                                    K[ch * n_loc + rl] += v;
                                }
                            }
                    }
                }
            // L2 masses: one (k, k') product per thread, repeated on the component diagonal
            for (int l = 0; l < P.n_l2; ++l) {
                const double *am = P.l_a_mass[l];
                if (!am) continue;
                const int npl = P.l_n_phi[l], ncl = P.l_n_comp[l];
                const double *lv = P.l_val[l] + (size_t)q0 * npl;
                for (int it = threadIdx.x; it < npl * npl; it += blockDim.x) {
                    const int k = it / npl, k2 = it - k * npl;
                    double m = 0.0;
                    for (int q = 0; q < nq; ++q) m += s_wd[q] * am[cq_base + q] * lv[q * npl + k] * lv[q * npl + k2];
                    for (int c = 0; c < ncl; ++c) {
                        const int e = (off_l[l] + k * ncl + c) * n_loc + off_l[l] + k2 * ncl + c;
                        K[e] = q0 == 0 ? m : K[e] + m;
                    }
                }
            }
        }
        for (int e = pair_pass ? n_loc * n_loc : threadIdx.x; e < n_loc * n_loc; e += blockDim.x) {
            int row = e / n_loc, col = e - row * n_loc;
            // decode (kind, field, basis, comp)
            int rk, rf, ra, rc, ck, cf, ca, cc;
            {
                int r = row;
                if (r < off_l[0]) {
                    rk = 0;
                    rf = (P.n_hdiv > 1 && r >= off_h[1]) ? 1 : 0;
                    r -= off_h[rf];
                    ra = r / P.h_n_comp[rf];
                    rc = r - ra * P.h_n_comp[rf];
                } else {
                    rk = 1;
                    rf = (P.n_l2 > 1 && r >= off_l[1]) ? 1 : 0;
                    r -= off_l[rf];
                    ra = r / P.l_n_comp[rf];
                    rc = r - ra * P.l_n_comp[rf];
                }
                int c = col;
                if (c < off_l[0]) {
                    ck = 0;
                    cf = (P.n_hdiv > 1 && c >= off_h[1]) ? 1 : 0;
                    c -= off_h[cf];
                    ca = c / P.h_n_comp[cf];
                    cc = c - ca * P.h_n_comp[cf];
                } else {
                    ck = 1;
                    cf = (P.n_l2 > 1 && c >= off_l[1]) ? 1 : 0;
                    c -= off_l[cf];
                    ca = c / P.l_n_comp[cf];
                    cc = c - ca * P.l_n_comp[cf];
                }
            }
            double acc = 0.0;
            if (pair_pass && rk == 0 && ck == 0 && rf == cf) continue;  // written by the pair pass above
            if (rk == 0 && ck == 0) {
                if (rf == cf) {
                    const int h = rf, np = P.h_n_phi[h];
                    const double *aid = P.h_a_id[h], *atr = P.h_a_tr[h], *avol = P.h_a_vol[h], *add = P.h_a_dd[h];
                    for (int q = 0; q < nq; ++q) {
                        const double *pa = s_psi[h] + (q * np + ra) * DIM, *pb = s_psi[h] + (q * np + ca) * DIM;
                        double t = 0.0;
                        if (rc == cc) {
                            double dot = 0.0;
#pragma unroll
                            for (int d = 0; d < DIM; ++d) dot += pa[d] * pb[d];
                            if (aid) t += aid[cq_base + q] * dot;
                            if (add) {
                                if (P.h_dd_general[h])  // s_gd carries det (cd div psi + g.psi)
                                    t += add[cq_base + q] * (s_gd[h][q * np + ra] * idet) * (s_gd[h][q * np + ca] * idet);
                                else
                                    t += add[cq_base + q] * s_div[h][q * np + ra] * s_div[h][q * np + ca];
                            }
                        }
                        if (rc < DIM && cc < DIM) {
                            if (atr) t += atr[cq_base + q] * pb[rc] * pa[cc];
                            if (avol) t += avol[cq_base + q] * pa[rc] * pb[cc];
                        }
                        acc += s_wd[q] * t;
                    }
                }
            } else if (rk == 1 && ck == 1) {
                if (rf == cf && rc == cc && P.l_a_mass[rf]) {
                    const int l = rf, np = P.l_n_phi[l];
                    const double *am = P.l_a_mass[l], *lv = P.l_val[l] + (size_t)q0 * np;
                    for (int q = 0; q < nq; ++q) acc += s_wd[q] * am[cq_base + q] * lv[q * np + ra] * lv[q * np + ca];
                }
            } else {
                // L2 x H(div) coupling (symmetric): put the L2 dof in (l, k, lc) and the H(div) dof in (h, a, hc)
                const int l = rk == 1 ? rf : cf, k = rk == 1 ? ra : ca, lc = rk == 1 ? rc : cc;
                const int h = rk == 1 ? cf : rf, a = rk == 1 ? ca : ra, hc = rk == 1 ? cc : rc;
                const int npl = P.l_n_phi[l], nph = P.h_n_phi[h];
                const double *lv = P.l_val[l] + (size_t)q0 * npl;
                if (P.div_coupling[l][h] && lc == hc) {
                    for (int q = 0; q < nq; ++q) acc += P.weights[q0 + q] * lv[q * npl + k] * s_gd[h][q * nph + a];
                }
                if (P.asym_coupling[l][h]) {
                    const int ce = P.l_n_comp[l] == 1 ? 2 : lc;  // the single rotation of a 2-D problem is the z one
                    for (int d = 0; d < DIM; ++d) {
                        const double eps = levi_civita(ce, hc, d);
                        if (eps != 0.0)
                            for (int q = 0; q < nq; ++q)
                                acc += eps * s_wd[q] * lv[q * npl + k] * s_psi[h][(q * nph + a) * DIM + d];
                    }
                }
            }
            K[e] = q0 == 0 ? acc : K[e] + acc;
        }
        }  // quadrature chunks
        __syncthreads();  // K of this cell is complete in global memory for the row products below
        // one warp per row: the lanes stride the row of K (coalesced) and the partial products meet in a shuffle tree
        for (int row = threadIdx.x >> 5; row < n_loc; row += blockDim.x >> 5) {
            const int lane = threadIdx.x & 31;
            double acc = 0.0;
            if (alpha && dest) {
                for (int c = lane; c < n_loc; c += 32) acc = fma(K[row * n_loc + c], s_a[c], acc);
            }
            if (row >= off_l[0]) {
                const int l = (P.n_l2 > 1 && row >= off_l[1]) ? 1 : 0;
                const int r = row - off_l[l], k = r / P.l_n_comp[l], c = r - k * P.l_n_comp[l];
                const double *ld = P.l_load[l];
                if (ld) {
                    const int np = P.l_n_phi[l], ncp = P.l_n_comp[l];
                    for (int q = lane; q < n_q_all; q += 32)
                        acc -= P.weights[q] * g.det * ld[(cell * (int64_t)n_q_all + q) * ncp + c] * P.l_val[l][q * np + k];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane != 0) continue;
            re[cell * n_loc + row] = acc;
        }
    }
}

}  // namespace

extern "C" int nfa_local_hdiv_tensor(const nfa_tensor_form *form, const double *coords, const int32_t *cell_nodes,
                                     int64_t n_cells, const double *alpha, const int32_t *dest, double *Ke, double *re,
                                     nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!form) return fail("nfa_local_hdiv_tensor: form is NULL");
    const nfa_tensor_form &P = *form;
    if (P.n_hdiv < 1 || P.n_hdiv > 2 || P.n_l2 < 0 || P.n_l2 > 2) return fail("nfa_local_hdiv_tensor: 1-2 H(div) and 0-2 L2 fields");
    if (P.dim != 2 && P.dim != 3) return fail("nfa_local_hdiv_tensor: dim must be 2 or 3");
    cudaStream_t s = (cudaStream_t)stream;
    size_t per_q = 1 + 2 * 4, fixed = 0;  // doubles per staged quadrature point / independent of the chunk size
    int n_loc = 0;
    for (int h = 0; h < P.n_hdiv; ++h) {
        if (!P.h_val[h] || !P.h_div[h] || !P.h_Tcases[h] || !P.h_cases[h]) return fail("nfa_local_hdiv_tensor: H(div) tables missing");
        const bool need_div = P.h_a_dd[h] != nullptr && !P.h_dd_general[h];
        per_q += (size_t)P.h_n_phi[h] * (P.dim + 1 + (need_div ? 1 : 0));
        fixed += (size_t)(P.dim + 1) * P.h_nfd[h] * P.h_nfd[h];
        n_loc += P.h_n_phi[h] * P.h_n_comp[h];
    }
    for (int l = 0; l < P.n_l2; ++l) {
        if (!P.l_val[l]) return fail("nfa_local_hdiv_tensor: L2 table missing");
        n_loc += P.l_n_phi[l] * P.l_n_comp[l];
    }
    fixed += n_loc;
    // all points at once when the staged bases fit in shared memory, else equal chunks of points
    const size_t budget = 222 * 1024 / sizeof(double);  // of the 227 KB a CTA may have
    if (fixed + per_q > budget) return fail("nfa_local_hdiv_tensor: element too large for the shared-memory staging");
    int n_qc = (int)std::min<size_t>((size_t)P.n_q, (budget - fixed) / per_q);
    if (const char *e = getenv("NFA_TENSOR_QCHUNK")) n_qc = std::max(1, std::min(n_qc, atoi(e)));  // tests
    const int n_chunks = (P.n_q + n_qc - 1) / n_qc;
    n_qc = (P.n_q + n_chunks - 1) / n_chunks;
    const size_t smem = sizeof(double) * (fixed + per_q * (size_t)n_qc);
    const int grid = (int)(n_cells < 148 * 8 ? n_cells : 148 * 8);
    // CTAs that stage more than 56 KB leave room for at most three per SM: they run 512 threads instead of 128
    const bool big = smem > 56 * 1024;
    using kern_t = void (*)(const nfa_tensor_form, int, const double *, const int32_t *, int64_t, const double *,
                            const int32_t *, double *, double *);
    kern_t kern = P.dim == 3 ? (big ? (kern_t)local_hdiv_tensor_kernel<3, 512> : (kern_t)local_hdiv_tensor_kernel<3, 128>)
                             : (big ? (kern_t)local_hdiv_tensor_kernel<2, 512> : (kern_t)local_hdiv_tensor_kernel<2, 128>);
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, big ? 512 : 128, smem, s>>>(P, n_qc, coords, cell_nodes, n_cells, alpha, dest, Ke, re);
    NFA_LAUNCH_CHECK();
    return 0;
}
