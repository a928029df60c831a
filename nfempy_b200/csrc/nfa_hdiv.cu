// libnfa: H(div) x L2 local kernels (K3): mixed Poisson volume form and its natural (Dirichlet) boundary load.
// Reference: LaplaceDualWeakForm / LaplaceDualWeakFormBCDirichlet, R/src/weak_forms/laplace_dual_weak_form.py:10-154.
// Basis handling restates FiniteElement.evaluate_basis (R/src/basis/finite_element.py:156-186): contravariant Piola
// (1/|det|) J phi_hat on the values, divergence = tr(grad_hat phi_hat)/|det|, then the orientation transformation of
// permute_and_transform (R/src/basis/permute_and_transform.py:110-140) as a small matrix per facet chosen by a
// per-cell case id computed on the host.
#include <algorithm>
#include <cstdlib>

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
                         const double *__restrict__ f, const double *__restrict__ f_u, double *__restrict__ Ke,
                         double *__restrict__ re) {
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
            if (ra < n_u && f_u)  // vector load -int f . psi (L2ProjectorWeakForm on an H(div) field)
                for (int q = 0; q < n_q; ++q) {
                    const double *ps = s_psi + (q * n_u + ra) * DIM;
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) dot += ps[c] * f_u[(cell * n_q + q) * DIM + c];
                    acc -= s_wd[q] * dot;
                }
            re[cell * n_loc + ra] = acc;
        }
    }
}

// Register-tiled version of local_hdiv_l2_kernel.  The per-entry loop above reads two shared-memory vectors per DFMA
// group (7 LDS for 4 DFMA: bound by shared-memory bandwidth, 37 ms for RT2 on 2 M tets); here a lane owns a T x T tile
// of the symmetric flux mass block (tiles on and above the diagonal only, mirrored on store: 2T LDS for T*T DFMA per
// point and component) or T rows of one pressure column of the divergence coupling (stored with both signs).
// G lanes share a cell (8 / 16 / 32: sub-warp groups synchronised with __syncwarp on the group's mask, several cells per
// CTA; 128: one CTA per cell), so that small elements do not idle a whole CTA.
// Staging: pass 1 writes the Piola-mapped reference basis J psi_hat / |det| and div_hat / |det| ([q][c][j], rows padded
// with zeros to a multiple of T), pass 2 applies the facet orientation matrices in place.
constexpr int HT_MAXFD = 10;  // facet dofs per facet: RT_k 1 / 3 / 6, BDM_k 3 / 6 / 10

template <int G>
__device__ __forceinline__ void group_sync() {
    if constexpr (G >= 128) {
        __syncthreads();
    } else if constexpr (G == 32) {
        __syncwarp();
    } else {
        const unsigned lane = threadIdx.x & 31u;
        __syncwarp((G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane / G * G));
    }
}

// Stages the oriented, Piola-mapped basis of one cell in the group's shared memory: s_psi [n_q][DIM][n_up] and
// s_div [n_q][n_up] (rows n_u .. n_up are zero).  Ends with a group_sync: the arrays are ready for every lane.
template <int DIM, int G>
__device__ __forceinline__ void stage_hdiv_basis(bool live, int r, int n_u, int n_up, int n_q, int nfd,
                                                 const double *__restrict__ u_val, const double *__restrict__ u_div,
                                                 const double *__restrict__ Tcases, const int8_t *__restrict__ cases,
                                                 int64_t cell, const double (&J)[3][DIM], double idet, double *s_psi,
                                                 double *s_div, double *s_T) {
    constexpr int NF = DIM + 1;
    if (live) {
        for (int i = r; i < NF * nfd * nfd; i += G) {
            const int fc = i / (nfd * nfd);
            s_T[i] = Tcases[(int)cases[cell * NF + fc] * nfd * nfd + (i - fc * nfd * nfd)];
        }
        // pass 1: Piola map of the reference basis (no orientation yet), zero padding of the tile rows
            for (int idx = r; idx < n_q * n_up; idx += G) {
                const int q = idx / n_up, j = idx - q * n_up;
                double vh[DIM], dh = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) vh[k] = 0.0;
                if (j < n_u) {
#pragma unroll
                    for (int k = 0; k < DIM; ++k) vh[k] = __ldg(u_val + ((int64_t)q * n_u + j) * DIM + k);
                    dh = __ldg(u_div + q * n_u + j);
                }
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    double v = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) v += J[c][k] * vh[k];
                    s_psi[(q * DIM + c) * n_up + j] = v * idet;
                }
                s_div[q * n_up + j] = dh * idet;
            }
        }
        group_sync<G>();
        if (live && nfd > 1) {
            // pass 2: orientation of the facet dofs, psi_a <- sum_b T[a][b] psi_b inside each facet block, in place: one
            // lane reads the facet's nfd values of one (point, component or divergence) and writes them back transformed
            for (int idx = r; idx < n_q * NF * (DIM + 1); idx += G) {
                const int q = idx / (NF * (DIM + 1)), rest = idx - q * NF * (DIM + 1);
                const int fc = rest / (DIM + 1), c = rest - fc * (DIM + 1);
                double *v = (c < DIM ? s_psi + (q * DIM + c) * n_up : s_div + q * n_up) + fc * nfd;
                const double *Tm = s_T + fc * nfd * nfd;
                double in[HT_MAXFD];
#pragma unroll
                for (int b = 0; b < HT_MAXFD; ++b) in[b] = b < nfd ? v[b] : 0.0;
                for (int a = 0; a < nfd; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int b = 0; b < HT_MAXFD; ++b)
                        if (b < nfd) acc += Tm[a * nfd + b] * in[b];
                    v[a] = acc;
                }
            }
        } else if (live && nfd == 1) {
            for (int idx = r; idx < n_q * NF * (DIM + 1); idx += G) {
                const int q = idx / (NF * (DIM + 1)), rest = idx - q * NF * (DIM + 1);
                const int fc = rest / (DIM + 1), c = rest - fc * (DIM + 1);
                double *v = (c < DIM ? s_psi + (q * DIM + c) * n_up : s_div + q * n_up) + fc;
                v[0] *= s_T[fc];
            }
        }
    group_sync<G>();
}

// MMA (G >= 32): the symmetric flux mass block M = Psi^T diag(w det / kappa) Psi, Psi [(n_q DIM) x n_u], is a small dense
// SYRK; a warp computes 8 x 8 tiles of it on the FP64 tensor cores (mma.sync m8n8k4, DMMA): per step of four (point,
// component) rows a lane loads one element of each operand from the staged basis -- 3 LDS + 1 DMUL + 1 DMMA for 256
// multiply-adds instead of 6 LDS + 9 DFMA for 9 per lane.  n_up is then padded to a multiple of 8 and the staged rows to a
// multiple of 4 (zero rows).
template <int DIM, int G, int T, bool MMA>
__global__ void __launch_bounds__(128)
    hdiv_l2_tiled_kernel(int n_u, int n_p, int n_q, int nfd, const double *__restrict__ u_val,
                         const double *__restrict__ u_div, const double *__restrict__ p_val,
                         const double *__restrict__ weights, const double *__restrict__ Tcases,
                         const int8_t *__restrict__ cases, const double *__restrict__ coords,
                         const int32_t *__restrict__ cell_nodes, int64_t n_cells, CoefH kappa,
                         const double *__restrict__ f, const double *__restrict__ f_u, double *__restrict__ Ke,
                         double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    constexpr int GROUPS = 128 / G;
    extern __shared__ double smem[];
    const int nt = (n_u + T - 1) / T;
    const int n_up = MMA ? ((n_u + 2 + 7) & ~7) : nt * T;
    const int n_rows = MMA ? ((n_q * DIM + 3) & ~3) : n_q * DIM;  // staged (point, component) rows of psi
    const int n_loc = n_u + n_p, n_fdofs = NF * nfd;
    const int group_doubles = n_rows * n_up + n_q * n_up + 2 * (n_q + 2) + NF * nfd * nfd;
    const int g = threadIdx.x / G, r = threadIdx.x - g * G;
    double *s_psi = smem + (size_t)g * group_doubles;  // [n_q][DIM][n_up] (+ zero rows up to n_rows)
    double *s_div = s_psi + n_rows * n_up;              // [n_q][n_up]
    double *s_wk = s_div + n_q * n_up;                  // [n_q + 2] w det / kappa (zero beyond n_q)
    double *s_wd = s_wk + n_q + 2;                      // [n_q + 2] w det
    double *s_T = s_wd + n_q + 2;                       // [NF][nfd][nfd]
    if (MMA) {  // the padding rows are never written again
        for (int i = r; i < (n_rows - n_q * DIM) * n_up; i += G) s_psi[n_q * DIM * n_up + i] = 0.0;
        if (r < 2) s_wk[n_q + r] = 0.0;
    }
    const int n_mass = nt * (nt + 1) / 2, n_cpl = nt * n_p, n_zero = n_p * n_p;
    const int n_work = n_mass + n_cpl + n_zero + n_loc;
    const int64_t n_trips = (n_cells + (int64_t)gridDim.x * GROUPS - 1) / ((int64_t)gridDim.x * GROUPS);
    // the vertex coordinates of the NEXT cell are loaded while the current one is computed (cell_nodes -> coords is a
    // chain of two dependent global loads that would otherwise sit at the head of every trip)
    double Xn[DIM + 1][3];
    {
        const int64_t c0 = (int64_t)blockIdx.x * GROUPS + g;
        if (c0 < n_cells) load_cell_coords<DIM>(coords, cell_nodes, c0, Xn);
    }
    for (int64_t trip = 0; trip < n_trips; ++trip) {
        // every group of a warp makes the same number of trips (group_sync names the group's lanes only, but all lanes of
        // a warp must reach the loop's convergence points); a group without a cell skips the work
        const int64_t cell = (trip * gridDim.x + blockIdx.x) * GROUPS + g;
        const bool live = cell < n_cells;
        double J[3][DIM], idet = 0.0, det = 0.0;
        if (live) {
            double X[DIM + 1][3];
#pragma unroll
            for (int a = 0; a <= DIM; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) X[a][b] = Xn[a][b];
            Geom<DIM> gm;
            compute_geom<DIM>(X, gm);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < DIM; ++b) J[a][b] = gm.J[a][b];
            det = gm.det;
            idet = 1.0 / gm.det;
        }
        group_sync<G>();  // the previous cell's readers are done with the group's shared memory
        if (live)
            for (int q = r; q < n_q; q += G) {
                const double wd = weights[q] * det;
                s_wd[q] = wd;
                s_wk[q] = wd * (1.0 / kappa.p[cell * kappa.sc + q * kappa.sq]);
            }
        stage_hdiv_basis<DIM, G>(live, r, n_u, n_up, n_q, nfd, u_val, u_div, Tcases, cases, cell, J, idet, s_psi, s_div, s_T);
        {
            const int64_t cn = ((trip + 1) * gridDim.x + blockIdx.x) * GROUPS + g;
            if (cn < n_cells) {
                load_cell_coords<DIM>(coords, cell_nodes, cn, Xn);
                if (r == 0) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(cases + cn * NF));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(kappa.p + cn * kappa.sc));
                }
            }
        }
        if constexpr (MMA) {
            constexpr int WARPS = G / 32;
            const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3, wig = (threadIdx.x >> 5) % WARPS;
            const int nt8 = (n_u + 7) / 8, n_tiles8 = nt8 * (nt8 + 1) / 2, k_steps = n_rows / 4;
            if (live) {
                double *K = Ke + cell * (int64_t)n_loc * n_loc;
                // up to four tiles per trip: their accumulation chains are independent, which hides the latency of the
                // dependent DMMA sequence of a single tile
                constexpr int NT = 4;
                for (int w0 = wig * NT; w0 < n_tiles8; w0 += WARPS * NT) {
                    int ti[NT], tj[NT];
                    const double *pa[NT], *pb[NT];
                    double c0[NT], c1[NT];
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
                        int t = 0, rem = min(w0 + u, n_tiles8 - 1);  // a surplus slot repeats the last tile (not stored)
                        while (rem >= nt8 - t) {
                            rem -= nt8 - t;
                            ++t;
                        }
                        ti[u] = t;
                        tj[u] = t + rem;
                        pa[u] = s_psi + tig * n_up + 8 * ti[u] + gid;
                        pb[u] = s_psi + tig * n_up + 8 * tj[u] + gid;
                        c0[u] = c1[u] = 0.0;
                    }
#pragma unroll 2
                    for (int st = 0; st < k_steps; ++st) {
                        const double wq = s_wk[(4 * st + tig) / DIM];
                        const size_t off = (size_t)4 * st * n_up;
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            const double a = pa[u][off] * wq;  // row k of diag(w) Psi, column 8 ti + gid
                            const double b = pb[u][off];       // row k of Psi, column 8 tj + gid
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                         : "+d"(c0[u]), "+d"(c1[u])
                                         : "d"(a), "d"(b));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
                        if (w0 + u >= n_tiles8) continue;
                        // accumulator layout: (row gid, columns 2 tig and 2 tig + 1) of the 8 x 8 tile
                        const int ra = 8 * ti[u] + gid, rb = 8 * tj[u] + 2 * tig;
                        if (ra < n_u) {
                            if (rb < n_u) {
                                K[ra * n_loc + rb] = c0[u];
                                if (ti[u] != tj[u]) K[rb * n_loc + ra] = c0[u];
                            }
                            if (rb + 1 < n_u) {
                                K[ra * n_loc + rb + 1] = c1[u];
                                if (ti[u] != tj[u]) K[(rb + 1) * n_loc + ra] = c1[u];
                            }
                        }
                    }
                }
            }
        }
        // few mass tiles (at most G / 2): lanes r and r + G / 2 share a tile, each sums one half of the quadrature points
        // and the halves are combined with a shuffle, so that all lanes of the group work in the dominant phase
        const bool split = !MMA && G <= 32 && n_mass <= G / 2;
        if (G <= 32 && split) {
            constexpr int H = G / 2;
            const int half = r / H, w = r - half * H;
            const bool mine = live && w < n_mass;
            int ti = 0, rem = mine ? w : 0;
            while (rem >= nt - ti) {
                rem -= nt - ti;
                ++ti;
            }
            const int tj = ti + rem;
            double acc[T][T];
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) acc[i][j] = 0.0;
            const int q_mid = (n_q + 1) / 2, q0 = half ? q_mid : 0, q1 = half ? n_q : q_mid;
            if (mine) {
#pragma unroll 2
                for (int q = q0; q < q1; ++q) {
                    double sq[T][T];
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) sq[i][j] = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        const double *row = s_psi + (q * DIM + c) * n_up;
                        double a[T], b[T];
#pragma unroll
                        for (int i = 0; i < T; ++i) {
                            a[i] = row[ti * T + i];
                            b[i] = row[tj * T + i];
                        }
#pragma unroll
                        for (int i = 0; i < T; ++i)
#pragma unroll
                            for (int j = 0; j < T; ++j) sq[i][j] = fma(a[i], b[j], sq[i][j]);
                    }
                    const double wk = s_wk[q];
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) acc[i][j] = fma(wk, sq[i][j], acc[i][j]);
                }
            }
            // lanes without a cell or a tile take part in the shuffle with zeros
            const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31u) / G * G);
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) acc[i][j] += __shfl_xor_sync(gmask, acc[i][j], H);
            if (mine) {
                double *K = Ke + cell * (int64_t)n_loc * n_loc;
#pragma unroll
                for (int i = 0; i < T; ++i)
#pragma unroll
                    for (int j = 0; j < T; ++j) {
                        const int ra = ti * T + i, rb = tj * T + j;
                        if (ra < n_u && rb < n_u) {
                            if (half == 0) K[ra * n_loc + rb] = acc[i][j];
                            if (half == 1 && ti != tj) K[rb * n_loc + ra] = acc[i][j];
                        }
                    }
            }
        }
        if (live) {
            double *K = Ke + cell * (int64_t)n_loc * n_loc;
            for (int w = ((split || MMA) ? n_mass : 0) + r; w < n_work; w += G) {
                if (w < n_mass) {
                    // tile (ti, tj), ti <= tj, from the linear index over the upper triangle
                    int ti = 0, rem = w;
                    while (rem >= nt - ti) {
                        rem -= nt - ti;
                        ++ti;
                    }
                    const int tj = ti + rem;
                    double acc[T][T];
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) acc[i][j] = 0.0;
#pragma unroll 2
                    for (int q = 0; q < n_q; ++q) {
                        double sq[T][T];
#pragma unroll
                        for (int i = 0; i < T; ++i)
#pragma unroll
                            for (int j = 0; j < T; ++j) sq[i][j] = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            const double *row = s_psi + (q * DIM + c) * n_up;
                            double a[T], b[T];
#pragma unroll
                            for (int i = 0; i < T; ++i) {
                                a[i] = row[ti * T + i];
                                b[i] = row[tj * T + i];
                            }
#pragma unroll
                            for (int i = 0; i < T; ++i)
#pragma unroll
                                for (int j = 0; j < T; ++j) sq[i][j] = fma(a[i], b[j], sq[i][j]);
                        }
                        const double wk = s_wk[q];
#pragma unroll
                        for (int i = 0; i < T; ++i)
#pragma unroll
                            for (int j = 0; j < T; ++j) acc[i][j] = fma(wk, sq[i][j], acc[i][j]);
                    }
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) {
                            const int ra = ti * T + i, rb = tj * T + j;
                            if (ra < n_u && rb < n_u) {
                                K[ra * n_loc + rb] = acc[i][j];
                                if (ti != tj) K[rb * n_loc + ra] = acc[i][j];
                            }
                        }
                } else if (w < n_mass + n_cpl) {
                    const int wc = w - n_mass, ti = wc / n_p, l = wc - ti * n_p;
                    double acc[T];
#pragma unroll
                    for (int i = 0; i < T; ++i) acc[i] = 0.0;
                    for (int q = 0; q < n_q; ++q) {
                        const double pw = s_wd[q] * __ldg(p_val + q * n_p + l);
#pragma unroll
                        for (int i = 0; i < T; ++i) acc[i] = fma(pw, s_div[q * n_up + ti * T + i], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const int ra = ti * T + i;
                        if (ra < n_u) {
                            K[ra * n_loc + n_u + l] = -acc[i];
                            K[(n_u + l) * n_loc + ra] = acc[i];
                        }
                    }
                } else if (w < n_mass + n_cpl + n_zero) {
                    const int wz = w - n_mass - n_cpl, k = wz / n_p, l = wz - k * n_p;
                    K[(n_u + k) * n_loc + n_u + l] = 0.0;
                } else {
                    const int ra = w - n_mass - n_cpl - n_zero;
                    double acc = 0.0;
                    if (ra >= n_u && f)
                        for (int q = 0; q < n_q; ++q) acc -= s_wd[q] * __ldg(p_val + q * n_p + (ra - n_u)) * f[cell * n_q + q];
                    if (ra < n_u && f_u)  // vector load -int f . psi (L2ProjectorWeakForm on an H(div) field)
                        for (int q = 0; q < n_q; ++q) {
                            double dot = 0.0;
#pragma unroll
                            for (int c = 0; c < DIM; ++c)
                                dot += s_psi[(q * DIM + c) * n_up + ra] * f_u[(cell * n_q + q) * DIM + c];
                            acc -= s_wd[q] * dot;
                        }
                    re[cell * n_loc + ra] = acc;
                }
            }
        }
    }
}

// Generalised H(div) x L2 block form covering the Arbogast-Taicher family of the reference
// (R/src/weak_forms/degenerate_elliptic_weak_form.py:41-144, unscaled_elliptic_weak_form.py:41-153,
//  scaled_to_physical_l2_projection.py:27-105).  With D_i = c_d div psi_i + g . psi_i (per point):
//   J = [[ int c_uu psi_j.psi_i , -int w_l D_i ], [ int D_j w_k , int c_pp w_l w_k ]],
//   r = J alpha - [ int s_u (beta_u.psi) . psi_i ; int s_p (beta_p w) w_k ] - [0 ; int load w_k].
// All coefficients are host-evaluated per quadrature point (the reference branches on isclose(porosity, 0) per
// point; those branches are applied when the arrays are built).
template <int DIM>
__global__ void __launch_bounds__(128)
    local_hdiv_general_kernel(int n_u, int n_p, int n_q, int nfd, const double *__restrict__ u_val,
                              const double *__restrict__ u_div, const double *__restrict__ p_val,
                              const double *__restrict__ weights, const double *__restrict__ Tcases,
                              const int8_t *__restrict__ cases, const double *__restrict__ coords,
                              const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                              const double *__restrict__ c_uu, const double *__restrict__ c_d,
                              const double *__restrict__ gvec, const double *__restrict__ c_pp,
                              const double *__restrict__ load_p, const double *__restrict__ alpha,
                              const double *__restrict__ beta, const int32_t *__restrict__ dest,
                              const double *__restrict__ s_u, const double *__restrict__ s_p,
                              double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    extern __shared__ double smem[];
    const int n_loc = n_u + n_p;
    double *s_psi = smem;                         // [n_q][n_u][DIM]
    double *s_D = s_psi + n_q * n_u * DIM;        // [n_q][n_u]   c_d div psi + g . psi
    double *s_wd = s_D + n_q * n_u;               // [n_q] w det
    double *s_T = s_wd + n_q;                     // [NF][nfd][nfd]
    double *s_a = s_T + NF * nfd * nfd;           // [n_loc] alpha (local)
    double *s_b = s_a + n_loc;                    // [n_loc] beta (local)
    double *s_bq = s_b + n_loc;                   // [n_q][DIM + 1] scaled beta field: s_u (beta_u.psi), s_p (beta_p w)
    const int n_fdofs = NF * nfd;
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const double idet = 1.0 / g.det;
        __syncthreads();
        for (int i = threadIdx.x; i < NF * nfd * nfd; i += blockDim.x) {
            int fc = i / (nfd * nfd);
            s_T[i] = Tcases[(int)cases[cell * NF + fc] * nfd * nfd + (i - fc * nfd * nfd)];
        }
        for (int q = threadIdx.x; q < n_q; q += blockDim.x) s_wd[q] = weights[q] * g.det;
        for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
            int d = dest ? dest[cell * n_loc + i] : 0;
            s_a[i] = alpha ? alpha[d] : 0.0;
            s_b[i] = beta ? beta[d] : 0.0;
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
            double D = (c_d ? c_d[cell * n_q + q] : 1.0) * (dh * idet);
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) v += g.J[c][k] * vh[k];
                v *= idet;
                s_psi[(q * n_u + j) * DIM + c] = v;
                if (gvec) D += gvec[(cell * n_q + q) * DIM + c] * v;
            }
            s_D[q * n_u + j] = D;
        }
        __syncthreads();
        if (beta) {
            for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
                double bu[DIM], bp = 0.0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) bu[c] = 0.0;
                for (int j = 0; j < n_u; ++j)
#pragma unroll
                    for (int c = 0; c < DIM; ++c) bu[c] += s_b[j] * s_psi[(q * n_u + j) * DIM + c];
                for (int k = 0; k < n_p; ++k) bp += s_b[n_u + k] * p_val[q * n_p + k];
                double su = s_u ? s_u[cell * n_q + q] : 1.0, sp = s_p ? s_p[cell * n_q + q] : 1.0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) s_bq[q * (DIM + 1) + c] = su * bu[c];
                s_bq[q * (DIM + 1) + DIM] = sp * bp;
            }
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
                    acc += s_wd[q] * c_uu[cell * n_q + q] * dot;
                }
            } else if (ra < n_u) {
                for (int q = 0; q < n_q; ++q) acc -= s_wd[q] * p_val[q * n_p + (rb - n_u)] * s_D[q * n_u + ra];
            } else if (rb < n_u) {
                for (int q = 0; q < n_q; ++q) acc += s_wd[q] * s_D[q * n_u + rb] * p_val[q * n_p + (ra - n_u)];
            } else if (c_pp) {
                for (int q = 0; q < n_q; ++q)
                    acc += s_wd[q] * c_pp[cell * n_q + q] * p_val[q * n_p + (rb - n_u)] * p_val[q * n_p + (ra - n_u)];
            }
            K[e] = acc;
        }
        __syncthreads();  // K of this cell is complete in global memory for the row products below
        for (int ra = threadIdx.x; ra < n_loc; ra += blockDim.x) {
            double acc = 0.0;
            if (alpha)
                for (int rb = 0; rb < n_loc; ++rb) acc += K[ra * n_loc + rb] * s_a[rb];
            for (int q = 0; q < n_q; ++q) {
                if (ra < n_u) {
                    if (beta) {
                        const double *ps = s_psi + (q * n_u + ra) * DIM;
                        double dot = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) dot += s_bq[q * (DIM + 1) + c] * ps[c];
                        acc -= s_wd[q] * dot;
                    }
                } else {
                    double w = p_val[q * n_p + (ra - n_u)];
                    if (beta) acc -= s_wd[q] * s_bq[q * (DIM + 1) + DIM] * w;
                    if (load_p) acc -= s_wd[q] * load_p[cell * n_q + q] * w;
                }
            }
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
                            double *__restrict__ re, int n_comp = 1, double sign = 1.0, int re_off = 0, int re_stride = 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (re_stride == 0) re_stride = nfd;
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
        for (int kc = lane; kc < nfd * n_comp; kc += 32) {
            const int k = kc / n_comp, c = kc - k * n_comp;
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
                acc += det_f * weights[q] * pd[(fc * n_q + q) * n_comp + c] * flux;
            }
            re[fc * re_stride + re_off + kc] = sign * acc;
        }
    }
}


// Mixed-dimensional Robin coupling across a fracture cell (CouplingWeakForm, R/examples/md_elliptic_scalar/
// CouplingWeakForm.py:11-151): the fracture cell c1 faces one boundary facet of the higher-dimensional domain on each
// side (+ / -).  With phi^s_k = psi^s_k . n_s the normal traces of the neighbours' facet functions (s = +, -), P_m the
// fracture's potential basis and 1/alpha = (delta / 2) / kappa_n, local dofs [u_+ (nfd) | u_- (nfd) | p (n_p)]:
//   J[u_s, u_s] = sum_q det_F w (1/alpha) phi^s phi^s^T,   J[u_s, p] = sum_q det_F w phi^s P^T,   J[p, u_s] = its transpose,
//   J[u_+, u_-] = J[p, p] = 0 (stored zeros: the reference scatters the whole local matrix),   r = J alpha_l.
// One warp per interface cell; the traces of both sides are staged in shared memory.
template <int DIM>
__global__ void __launch_bounds__(128)
    local_hdiv_coupling_kernel(int nfd, int n_p, int n_q, const double *__restrict__ trace_tab, const double *__restrict__ weights,
                               const double *__restrict__ Tcases, const double *__restrict__ coords,
                               const int32_t *__restrict__ tidx_a, const int8_t *__restrict__ cases_a,
                               const int32_t *__restrict__ facet_a, const int32_t *__restrict__ neigh_a,
                               const double *__restrict__ normals_a, const int32_t *__restrict__ tidx_b,
                               const int8_t *__restrict__ cases_b, const int32_t *__restrict__ facet_b,
                               const int32_t *__restrict__ neigh_b, const double *__restrict__ normals_b,
                               const double *__restrict__ p_val, const double *__restrict__ ainv, int64_t n_items,
                               const double *__restrict__ alpha, const int32_t *__restrict__ dest, double *__restrict__ Ke,
                               double *__restrict__ re) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_loc = 2 * nfd + n_p;
    double *s_phi = smem + (size_t)warp * ((2 * nfd + 1) * n_q + n_loc);  // [2][n_q][nfd] traces, [n_q] det_F w, [n_loc] alpha
    double *s_w = s_phi + 2 * nfd * n_q;
    double *s_al = s_w + n_q;
    for (int64_t it = (int64_t)blockIdx.x * 4 + warp; it < n_items; it += (int64_t)gridDim.x * 4) {
        double det_f = 1.0;
        for (int side = 0; side < 2; ++side) {
            const int32_t *tidx = side ? tidx_b : tidx_a;
            const int8_t *cases = side ? cases_b : cases_a;
            const int32_t *neigh = side ? neigh_b : neigh_a;
            const double *normals = side ? normals_b : normals_a;
            double X[DIM + 1][3];
            load_cell_coords<DIM>(coords, neigh, it, X);
            Geom<DIM> g;
            compute_geom<DIM>(X, g);
            if (side == 0) {
                double XF[DIM][3];
                load_cell_coords<DIM - 1>(coords, facet_a, it, XF);
                Geom<DIM - 1> gf;
                compute_geom<DIM - 1>(XF, gf);
                det_f = gf.det;
            }
            double jn[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) v += g.J[c][k] * normals[it * 3 + c];
                jn[k] = v / g.det;
            }
            const double *T = Tcases + (int)cases[it] * nfd * nfd;
            const double *tab = trace_tab + (int64_t)tidx[it] * n_q * nfd * DIM;
            for (int e = lane; e < n_q * nfd; e += 32) {
                const int q = e / nfd, k = e - q * nfd;
                double flux = 0.0;
                for (int b = 0; b < nfd; ++b) {
                    const double *src = tab + ((int64_t)q * nfd + b) * DIM;
                    double t = 0.0;
#pragma unroll
                    for (int kk = 0; kk < DIM; ++kk) t += jn[kk] * src[kk];
                    flux += T[k * nfd + b] * t;
                }
                s_phi[(side * n_q + q) * nfd + k] = flux;
            }
        }
        for (int q = lane; q < n_q; q += 32) s_w[q] = det_f * weights[q];
        for (int j = lane; j < n_loc; j += 32) s_al[j] = alpha ? alpha[dest[it * n_loc + j]] : 0.0;
        __syncwarp();
        for (int e = lane; e < n_loc * n_loc; e += 32) {
            const int i = e / n_loc, j = e - i * n_loc;
            const int bi = i < nfd ? 0 : (i < 2 * nfd ? 1 : 2), bj = j < nfd ? 0 : (j < 2 * nfd ? 1 : 2);
            double acc = 0.0;
            if (bi < 2 && bj == bi) {
                const double *pa = s_phi + bi * n_q * nfd + (i - bi * nfd), *pb = s_phi + bi * n_q * nfd + (j - bi * nfd);
                for (int q = 0; q < n_q; ++q) acc += s_w[q] * ainv[it * n_q + q] * pa[q * nfd] * pb[q * nfd];
            } else if (bi < 2 && bj == 2) {
                const double *pa = s_phi + bi * n_q * nfd + (i - bi * nfd);
                for (int q = 0; q < n_q; ++q) acc += s_w[q] * pa[q * nfd] * p_val[q * n_p + (j - 2 * nfd)];
            } else if (bi == 2 && bj < 2) {
                const double *pb = s_phi + bj * n_q * nfd + (j - bj * nfd);
                for (int q = 0; q < n_q; ++q) acc += s_w[q] * p_val[q * n_p + (i - 2 * nfd)] * pb[q * nfd];
            }
            if (Ke) Ke[it * n_loc * n_loc + e] = acc;
            // r_i = sum_j J_ij alpha_j: lanes of one row are combined below (n_loc is small: done per lane over all j)
        }
        for (int i = lane; i < n_loc; i += 32) {
            double r = 0.0;
            if (alpha) {
                const int bi = i < nfd ? 0 : (i < 2 * nfd ? 1 : 2);
                for (int j = 0; j < n_loc; ++j) {
                    const int bj = j < nfd ? 0 : (j < 2 * nfd ? 1 : 2);
                    double acc = 0.0;
                    if (bi < 2 && bj == bi) {
                        const double *pa = s_phi + bi * n_q * nfd + (i - bi * nfd), *pb = s_phi + bi * n_q * nfd + (j - bi * nfd);
                        for (int q = 0; q < n_q; ++q) acc += s_w[q] * ainv[it * n_q + q] * pa[q * nfd] * pb[q * nfd];
                    } else if (bi < 2 && bj == 2) {
                        const double *pa = s_phi + bi * n_q * nfd + (i - bi * nfd);
                        for (int q = 0; q < n_q; ++q) acc += s_w[q] * pa[q * nfd] * p_val[q * n_p + (j - 2 * nfd)];
                    } else if (bi == 2 && bj < 2) {
                        const double *pb = s_phi + bj * n_q * nfd + (j - bj * nfd);
                        for (int q = 0; q < n_q; ++q) acc += s_w[q] * p_val[q * n_p + (i - 2 * nfd)] * pb[q * nfd];
                    }
                    r += acc * s_al[j];
                }
            }
            re[it * n_loc + i] = r;
        }
        __syncwarp();
    }
}

// Penalised normal trace ("Neumann" condition of the mixed elasticity forms: LCEDualWeakFormBCNeumann,
// R/src/weak_forms/lce_dual_weak_form.py:759-871, le_dual_weak_form.py:581-644, lce_scaled_dual_weak_form.py:814-917):
// with phi_k = psi_k . n the normal trace of the neighbour's facet functions, per component c of the row-wise field
//   J[(k, c), (l, c)] = beta sum_q det_F w phi_k phi_l,     r[(k, c)] = -beta sum_q det_F w g_c(x_q) phi_k.
// One warp per facet; the block of this field sits at [off, off + nfd * n_comp) of the facet's local numbering (the
// caller zeroes Ke: blocks coupling different fields / components are stored zeros).
template <int DIM>
__global__ void __launch_bounds__(128)
    local_hdiv_facet_penalty_kernel(int nfd, int n_q, int n_comp, const double *__restrict__ trace_tab,
                                    const int32_t *__restrict__ tidx, const double *__restrict__ weights,
                                    const double *__restrict__ Tcases, const int8_t *__restrict__ cases,
                                    const double *__restrict__ coords, const int32_t *__restrict__ facet_nodes,
                                    const int32_t *__restrict__ neigh_nodes, const double *__restrict__ normals,
                                    const double *__restrict__ g, int64_t n_facets, double beta, int off, int n_loc,
                                    double *__restrict__ Ke, double *__restrict__ re) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s_phi = smem + (size_t)warp * (nfd * n_q + n_q);  // [n_q][nfd] normal traces, then [n_q] det_F w
    double *s_w = s_phi + nfd * n_q;
    for (int64_t fc = (int64_t)blockIdx.x * 4 + warp; fc < n_facets; fc += (int64_t)gridDim.x * 4) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, neigh_nodes, fc, X);
        Geom<DIM> gm;
        compute_geom<DIM>(X, gm);
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
        double jn[DIM];  // (J^T n) / det: psi . n = (1 / det) n^T J psi_hat
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) v += gm.J[c][k] * normals[fc * 3 + c];
            jn[k] = v / gm.det;
        }
        __syncwarp();
        for (int i = lane; i < n_q * nfd; i += 32) {
            const int q = i / nfd, k = i - q * nfd;
            double flux = 0.0;
            for (int b = 0; b < nfd; ++b) {
                const double *src = tab + ((int64_t)q * nfd + b) * DIM;
                double t = 0.0;
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk) t += jn[kk] * src[kk];
                flux += T[k * nfd + b] * t;
            }
            s_phi[i] = flux;
        }
        for (int q = lane; q < n_q; q += 32) s_w[q] = det_f * weights[q];
        __syncwarp();
        double *K = Ke + fc * (int64_t)n_loc * n_loc;
        for (int i = lane; i < nfd * nfd; i += 32) {
            const int k = i / nfd, l = i - k * nfd;
            double acc = 0.0;
            for (int q = 0; q < n_q; ++q) acc += beta * s_w[q] * s_phi[q * nfd + k] * s_phi[q * nfd + l];
            for (int c = 0; c < n_comp; ++c) K[(off + k * n_comp + c) * (int64_t)n_loc + off + l * n_comp + c] = acc;
        }
        for (int kc = lane; kc < nfd * n_comp; kc += 32) {
            const int k = kc / n_comp, c = kc - k * n_comp;
            double acc = 0.0;
            for (int q = 0; q < n_q; ++q) acc -= beta * s_w[q] * g[(fc * n_q + q) * n_comp + c] * s_phi[q * nfd + k];
            re[fc * n_loc + off + kc] = acc;
        }
    }
}

// Two-flux mixed transport (SundusDualWeakForm, R/examples/parabolic_diffusive_problem/SundusWeakForm.py:11-177):
// fields [mp | mc | p | c], mp/mc in the same H(div) element, p/c in the same L2 element.  Nonlinear in c through the
// reaction R = 1 + eta c^2, so the kernel evaluates the state at the current iterate and writes residual and tangent.
template <int DIM>
__global__ void __launch_bounds__(128)
    local_sundus_kernel(int n_u, int n_p, int n_q, int nfd, const double *__restrict__ u_val,
                        const double *__restrict__ u_div, const double *__restrict__ p_val,
                        const double *__restrict__ weights, const double *__restrict__ Tcases,
                        const int8_t *__restrict__ cases, const double *__restrict__ coords,
                        const int32_t *__restrict__ cell_nodes, int64_t n_cells, CoefH kappa, CoefH delta, CoefH eta,
                        double inv_dt, const double *__restrict__ f_f, const double *__restrict__ f_r,
                        const double *__restrict__ alpha, const double *__restrict__ alpha_prev,
                        const int32_t *__restrict__ dest, double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    extern __shared__ double smem[];
    const int n_loc = 2 * n_u + 2 * n_p;
    double *s_psi = smem;                         // [n_q][n_u][DIM]
    double *s_div = s_psi + n_q * n_u * DIM;      // [n_q][n_u]
    double *s_T = s_div + n_q * n_u;              // [NF][nfd][nfd]
    double *s_a = s_T + NF * nfd * nfd;           // [n_loc] current iterate (local)
    double *s_an = s_a + n_loc;                   // [n_loc] previous time level (local)
    double *s_q = s_an + n_loc;                   // [n_q][QS] per-point state
    constexpr int QS = 8 + 2 * DIM;               // wd, wd/kappa, wd/delta, eta, p_h, c_h, dpdt, dcdt | mp_h[DIM] | mc_h[DIM]
    double *s_dv = s_q + n_q * QS;                // [n_q][2] div mp_h, div mc_h
    const int n_fdofs = NF * nfd;
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const double idet = 1.0 / g.det;
        __syncthreads();
        for (int i = threadIdx.x; i < NF * nfd * nfd; i += blockDim.x) {
            int fc = i / (nfd * nfd);
            s_T[i] = Tcases[(int)cases[cell * NF + fc] * nfd * nfd + (i - fc * nfd * nfd)];
        }
        for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
            int d = dest[cell * n_loc + i];
            s_a[i] = alpha ? alpha[d] : 0.0;
            s_an[i] = alpha_prev ? alpha_prev[d] : 0.0;
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
        // per-point state at the current iterate
        for (int q = threadIdx.x; q < n_q; q += blockDim.x) {
            double *st = s_q + q * QS;
            double wd = weights[q] * g.det;
            st[0] = wd;
            st[1] = wd * (1.0 / kappa.p[cell * kappa.sc + q * kappa.sq]);
            st[2] = wd * (1.0 / delta.p[cell * delta.sc + q * delta.sq]);
            st[3] = eta.p[cell * eta.sc + q * eta.sq];
            double ph = 0, ch = 0, pn = 0, cn = 0;
            for (int k = 0; k < n_p; ++k) {
                double w = p_val[q * n_p + k];
                ph += s_a[2 * n_u + k] * w;
                ch += s_a[2 * n_u + n_p + k] * w;
                pn += s_an[2 * n_u + k] * w;
                cn += s_an[2 * n_u + n_p + k] * w;
            }
            st[4] = ph;
            st[5] = ch;
            st[6] = (ph - pn) * inv_dt;
            st[7] = (ch - cn) * inv_dt;
            double dmp = 0, dmc = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) st[8 + c] = st[8 + DIM + c] = 0.0;
            for (int j = 0; j < n_u; ++j) {
                const double *ps = s_psi + (q * n_u + j) * DIM;
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    st[8 + c] += s_a[j] * ps[c];
                    st[8 + DIM + c] += s_a[n_u + j] * ps[c];
                }
                dmp += s_a[j] * s_div[q * n_u + j];
                dmc += s_a[n_u + j] * s_div[q * n_u + j];
            }
            s_dv[2 * q] = dmp;
            s_dv[2 * q + 1] = dmc;
        }
        __syncthreads();
        double *K = Ke + cell * (int64_t)n_loc * n_loc;
        for (int e = threadIdx.x; e < n_loc * n_loc; e += blockDim.x) {
            int ra = e / n_loc, rb = e - ra * n_loc;
            // field of row / column: 0 mp, 1 mc, 2 p, 3 c; index inside the field
            int fa = ra < n_u ? 0 : (ra < 2 * n_u ? 1 : (ra < 2 * n_u + n_p ? 2 : 3));
            int fb = rb < n_u ? 0 : (rb < 2 * n_u ? 1 : (rb < 2 * n_u + n_p ? 2 : 3));
            int ia = ra - (fa == 0 ? 0 : fa == 1 ? n_u : fa == 2 ? 2 * n_u : 2 * n_u + n_p);
            int ib = rb - (fb == 0 ? 0 : fb == 1 ? n_u : fb == 2 ? 2 * n_u : 2 * n_u + n_p);
            double acc = 0.0;
            if (fa <= 1 && fb == fa) {  // flux mass with kappa^-1 / delta^-1
                for (int q = 0; q < n_q; ++q) {
                    const double *a = s_psi + (q * n_u + ia) * DIM, *b = s_psi + (q * n_u + ib) * DIM;
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) dot += a[c] * b[c];
                    acc += s_q[q * QS + 1 + fa] * dot;
                }
            } else if (fa <= 1 && fb == fa + 2) {  // -(w_l, div psi_i)
                for (int q = 0; q < n_q; ++q) acc -= s_q[q * QS] * p_val[q * n_p + ib] * s_div[q * n_u + ia];
            } else if (fa >= 2 && fb == fa - 2) {  // +(div psi_j, w_k)
                for (int q = 0; q < n_q; ++q) acc += s_q[q * QS] * s_div[q * n_u + ib] * p_val[q * n_p + ia];
            } else if (fa >= 2 && fb == fa) {  // time derivative (+ reaction tangent 2 eta c_h on the c block)
                for (int q = 0; q < n_q; ++q) {
                    const double *st = s_q + q * QS;
                    double coef = inv_dt + (fa == 3 ? 2.0 * st[3] * st[5] : 0.0);
                    acc += st[0] * coef * p_val[q * n_p + ib] * p_val[q * n_p + ia];
                }
            }
            K[e] = acc;
        }
        for (int ra = threadIdx.x; ra < n_loc; ra += blockDim.x) {
            int fa = ra < n_u ? 0 : (ra < 2 * n_u ? 1 : (ra < 2 * n_u + n_p ? 2 : 3));
            int ia = ra - (fa == 0 ? 0 : fa == 1 ? n_u : fa == 2 ? 2 * n_u : 2 * n_u + n_p);
            double acc = 0.0;
            for (int q = 0; q < n_q; ++q) {
                const double *st = s_q + q * QS;
                if (fa <= 1) {
                    const double *ps = s_psi + (q * n_u + ia) * DIM;
                    double dot = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) dot += st[8 + fa * DIM + c] * ps[c];
                    acc += st[1 + fa] * dot - st[0] * st[4 + fa] * s_div[q * n_u + ia];
                } else {
                    double w = p_val[q * n_p + ia];
                    double src = (fa == 2) ? (f_f ? f_f[cell * n_q + q] : 0.0) : (f_r ? f_r[cell * n_q + q] : 0.0);
                    double body = st[6 + (fa - 2)] + s_dv[2 * q + (fa - 2)];
                    if (fa == 3) body += 1.0 + st[3] * st[5] * st[5];
                    acc += st[0] * (body - src) * w;
                }
            }
            re[cell * n_loc + ra] = acc;
        }
    }
}

// Register-tiled, several-cells-per-CTA version of local_sundus_kernel (same staging and tile scheme as
// hdiv_l2_tiled_kernel).  The two flux mass blocks share the basis products (one tile pass, two weights), the two
// divergence couplings are the same numbers, every other block is zero or an L2 mass.
template <int DIM, int G, int T>
__global__ void __launch_bounds__(128)
    sundus_tiled_kernel(int n_u, int n_p, int n_q, int nfd, const double *__restrict__ u_val,
                        const double *__restrict__ u_div, const double *__restrict__ p_val,
                        const double *__restrict__ weights, const double *__restrict__ Tcases,
                        const int8_t *__restrict__ cases, const double *__restrict__ coords,
                        const int32_t *__restrict__ cell_nodes, int64_t n_cells, CoefH kappa, CoefH delta, CoefH eta,
                        double inv_dt, const double *__restrict__ f_f, const double *__restrict__ f_r,
                        const double *__restrict__ alpha, const double *__restrict__ alpha_prev,
                        const int32_t *__restrict__ dest, double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NF = DIM + 1;
    constexpr int GROUPS = 128 / G;
    constexpr int QS = 8 + 2 * DIM;  // wd, wd/kappa, wd/delta, eta, p_h, c_h, dpdt, dcdt | mp_h[DIM] | mc_h[DIM]
    extern __shared__ double smem[];
    const int nt = (n_u + T - 1) / T, n_up = nt * T;
    const int n_loc = 2 * n_u + 2 * n_p;
    const int group_doubles = n_q * DIM * n_up + n_q * n_up + NF * nfd * nfd + 2 * n_loc + n_q * QS + 2 * n_q;
    const int g = threadIdx.x / G, r = threadIdx.x - g * G;
    double *s_psi = smem + (size_t)g * group_doubles;  // [n_q][DIM][n_up]
    double *s_div = s_psi + n_q * DIM * n_up;           // [n_q][n_up]
    double *s_T = s_div + n_q * n_up;                   // [NF][nfd][nfd]
    double *s_a = s_T + NF * nfd * nfd;                 // [n_loc] current iterate (local)
    double *s_an = s_a + n_loc;                         // [n_loc] previous time level (local)
    double *s_q = s_an + n_loc;                         // [n_q][QS] per-point state
    double *s_dv = s_q + n_q * QS;                      // [n_q][2] div mp_h, div mc_h
    const int n_mass = nt * (nt + 1) / 2, n_cpl = nt * n_p, n_pp = n_p * n_p;
    const int n_work = n_mass + n_cpl + n_pp + n_loc;
    const int oc = n_u, op = 2 * n_u, occ = 2 * n_u + n_p;  // offsets of mc, p, c in the local numbering
    const int64_t n_trips = (n_cells + (int64_t)gridDim.x * GROUPS - 1) / ((int64_t)gridDim.x * GROUPS);
    double Xn[DIM + 1][3];
    {
        const int64_t c0 = (int64_t)blockIdx.x * GROUPS + g;
        if (c0 < n_cells) load_cell_coords<DIM>(coords, cell_nodes, c0, Xn);
    }
    for (int64_t trip = 0; trip < n_trips; ++trip) {
        const int64_t cell = (trip * gridDim.x + blockIdx.x) * GROUPS + g;
        const bool live = cell < n_cells;
        double J[3][DIM], idet = 0.0, det = 0.0;
        if (live) {
            double X[DIM + 1][3];
#pragma unroll
            for (int a = 0; a <= DIM; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) X[a][b] = Xn[a][b];
            Geom<DIM> gm;
            compute_geom<DIM>(X, gm);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < DIM; ++b) J[a][b] = gm.J[a][b];
            det = gm.det;
            idet = 1.0 / gm.det;
        }
        group_sync<G>();  // the previous cell's readers are done with the group's shared memory
        if (live) {
            for (int i = r; i < n_loc; i += G) {
                const int d = dest[cell * n_loc + i];
                s_a[i] = alpha ? alpha[d] : 0.0;
                s_an[i] = alpha_prev ? alpha_prev[d] : 0.0;
            }
            double *K = Ke + cell * (int64_t)n_loc * n_loc;
            for (int e = r; e < n_loc * n_loc; e += G) K[e] = 0.0;  // the structurally zero blocks; the rest is overwritten
        }
        stage_hdiv_basis<DIM, G>(live, r, n_u, n_up, n_q, nfd, u_val, u_div, Tcases, cases, cell, J, idet, s_psi, s_div, s_T);
        {
            const int64_t cn = ((trip + 1) * gridDim.x + blockIdx.x) * GROUPS + g;
            if (cn < n_cells) load_cell_coords<DIM>(coords, cell_nodes, cn, Xn);
        }
        if (live) {
            // per-point state at the current iterate
            for (int q = r; q < n_q; q += G) {
                double *st = s_q + q * QS;
                const double wd = weights[q] * det;
                st[0] = wd;
                st[1] = wd * (1.0 / kappa.p[cell * kappa.sc + q * kappa.sq]);
                st[2] = wd * (1.0 / delta.p[cell * delta.sc + q * delta.sq]);
                st[3] = eta.p[cell * eta.sc + q * eta.sq];
                double ph = 0, ch = 0, pn = 0, cn = 0;
                for (int k = 0; k < n_p; ++k) {
                    const double w = __ldg(p_val + q * n_p + k);
                    ph += s_a[op + k] * w;
                    ch += s_a[occ + k] * w;
                    pn += s_an[op + k] * w;
                    cn += s_an[occ + k] * w;
                }
                st[4] = ph;
                st[5] = ch;
                st[6] = (ph - pn) * inv_dt;
                st[7] = (ch - cn) * inv_dt;
                double mp[DIM], mc[DIM], dmp = 0, dmc = 0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) mp[c] = mc[c] = 0.0;
                for (int j = 0; j < n_u; ++j) {
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        const double ps = s_psi[(q * DIM + c) * n_up + j];
                        mp[c] += s_a[j] * ps;
                        mc[c] += s_a[oc + j] * ps;
                    }
                    dmp += s_a[j] * s_div[q * n_up + j];
                    dmc += s_a[oc + j] * s_div[q * n_up + j];
                }
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    st[8 + c] = mp[c];
                    st[8 + DIM + c] = mc[c];
                }
                s_dv[2 * q] = dmp;
                s_dv[2 * q + 1] = dmc;
            }
        }
        group_sync<G>();  // state ready; the zero fill of K is ordered before the block stores below
        if (live) {
            double *K = Ke + cell * (int64_t)n_loc * n_loc;
            for (int w = r; w < n_work; w += G) {
                if (w < n_mass) {
                    int ti = 0, rem = w;
                    while (rem >= nt - ti) {
                        rem -= nt - ti;
                        ++ti;
                    }
                    const int tj = ti + rem;
                    double ak[T][T], ad[T][T];
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) ak[i][j] = ad[i][j] = 0.0;
                    for (int q = 0; q < n_q; ++q) {
                        double sq[T][T];
#pragma unroll
                        for (int i = 0; i < T; ++i)
#pragma unroll
                            for (int j = 0; j < T; ++j) sq[i][j] = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            const double *row = s_psi + (q * DIM + c) * n_up;
                            double a[T], b[T];
#pragma unroll
                            for (int i = 0; i < T; ++i) {
                                a[i] = row[ti * T + i];
                                b[i] = row[tj * T + i];
                            }
#pragma unroll
                            for (int i = 0; i < T; ++i)
#pragma unroll
                                for (int j = 0; j < T; ++j) sq[i][j] = fma(a[i], b[j], sq[i][j]);
                        }
                        const double wk = s_q[q * QS + 1], wl = s_q[q * QS + 2];
#pragma unroll
                        for (int i = 0; i < T; ++i)
#pragma unroll
                            for (int j = 0; j < T; ++j) {
                                ak[i][j] = fma(wk, sq[i][j], ak[i][j]);
                                ad[i][j] = fma(wl, sq[i][j], ad[i][j]);
                            }
                    }
#pragma unroll
                    for (int i = 0; i < T; ++i)
#pragma unroll
                        for (int j = 0; j < T; ++j) {
                            const int ra = ti * T + i, rb = tj * T + j;
                            if (ra < n_u && rb < n_u) {
                                K[ra * n_loc + rb] = ak[i][j];
                                K[(oc + ra) * n_loc + oc + rb] = ad[i][j];
                                if (ti != tj) {
                                    K[rb * n_loc + ra] = ak[i][j];
                                    K[(oc + rb) * n_loc + oc + ra] = ad[i][j];
                                }
                            }
                        }
                } else if (w < n_mass + n_cpl) {
                    const int wc = w - n_mass, ti = wc / n_p, l = wc - ti * n_p;
                    double acc[T];
#pragma unroll
                    for (int i = 0; i < T; ++i) acc[i] = 0.0;
                    for (int q = 0; q < n_q; ++q) {
                        const double pw = s_q[q * QS] * __ldg(p_val + q * n_p + l);
#pragma unroll
                        for (int i = 0; i < T; ++i) acc[i] = fma(pw, s_div[q * n_up + ti * T + i], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const int ra = ti * T + i;
                        if (ra < n_u) {
                            K[ra * n_loc + op + l] = -acc[i];         // -(w_l, div psi_i), flux mp / pressure p
                            K[(op + l) * n_loc + ra] = acc[i];        // +(div psi_j, w_k)
                            K[(oc + ra) * n_loc + occ + l] = -acc[i];  // the same for mc / c
                            K[(occ + l) * n_loc + oc + ra] = acc[i];
                        }
                    }
                } else if (w < n_mass + n_cpl + n_pp) {
                    const int wz = w - n_mass - n_cpl, k = wz / n_p, l = wz - k * n_p;
                    double app = 0.0, acc = 0.0;  // time derivative (+ reaction tangent 2 eta c_h on the c block)
                    for (int q = 0; q < n_q; ++q) {
                        const double *st = s_q + q * QS;
                        const double ww = st[0] * __ldg(p_val + q * n_p + k) * __ldg(p_val + q * n_p + l);
                        app = fma(inv_dt, ww, app);
                        acc = fma(inv_dt + 2.0 * st[3] * st[5], ww, acc);
                    }
                    K[(op + k) * n_loc + op + l] = app;
                    K[(occ + k) * n_loc + occ + l] = acc;
                } else {
                    const int ra = w - n_mass - n_cpl - n_pp;
                    const int fa = ra < n_u ? 0 : (ra < 2 * n_u ? 1 : (ra < 2 * n_u + n_p ? 2 : 3));
                    const int ia = ra - (fa == 0 ? 0 : fa == 1 ? n_u : fa == 2 ? 2 * n_u : 2 * n_u + n_p);
                    double acc = 0.0;
                    for (int q = 0; q < n_q; ++q) {
                        const double *st = s_q + q * QS;
                        if (fa <= 1) {
                            double dot = 0.0;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) dot += st[8 + fa * DIM + c] * s_psi[(q * DIM + c) * n_up + ia];
                            acc += st[1 + fa] * dot - st[0] * st[4 + fa] * s_div[q * n_up + ia];
                        } else {
                            const double wv = __ldg(p_val + q * n_p + ia);
                            const double src = (fa == 2) ? (f_f ? f_f[cell * n_q + q] : 0.0) : (f_r ? f_r[cell * n_q + q] : 0.0);
                            double body = st[6 + (fa - 2)] + s_dv[2 * q + (fa - 2)];
                            if (fa == 3) body += 1.0 + st[3] * st[5] * st[5];
                            acc += st[0] * (body - src) * wv;
                        }
                    }
                    re[cell * n_loc + ra] = acc;
                }
            }
        }
    }
}

}  // namespace

extern "C" int nfa_local_hdiv_l2(int dim, int n_u, int n_p, int n_q, int nfd, const double *u_val,
                                 const double *u_div, const double *p_val, const double *weights,
                                 const double *Tcases, const int8_t *cases, const double *coords,
                                 const int32_t *cell_nodes, int64_t n_cells, const double *kappa, int64_t ksc,
                                 int64_t ksq, const double *f, const double *f_u, double *Ke, double *re,
                                 nfa_stream stream) {
    if (n_cells == 0) return 0;
    if ((dim + 1) * nfd > n_u) return fail("nfa_local_hdiv_l2: facet dofs exceed the basis size");
    cudaStream_t s = (cudaStream_t)stream;
    CoefH kp{kappa, ksc, ksq};
    size_t smem = sizeof(double) * ((size_t)n_q * n_u * (dim + 1) + 2 * n_q + (size_t)(dim + 1) * nfd * nfd);
    if (smem > 200 * 1024) return fail("nfa_local_hdiv_l2: element too large for the shared-memory staging");
    int grid = (int)(n_cells < 148 * 8 ? n_cells : 148 * 8);
    const char *e_tiled = getenv("NFA_HDIV_TILED");
    if ((!e_tiled || atoi(e_tiled) != 0) && nfd <= HT_MAXFD) {
        // register-tiled kernel: lanes per cell from the number of heavy work items (mass tiles + coupling columns)
        constexpr int T = 3;
        const int nt = (n_u + T - 1) / T, heavy = nt * (nt + 1) / 2 + nt * n_p;
        int G = heavy <= 8 ? 8 : (heavy <= 16 ? 16 : (heavy <= 48 ? 32 : 128));
        if (const char *e = getenv("NFA_HDIV_LANES")) G = atoi(e);
        // FP64 tensor cores (DMMA) for the mass block when whole warps share a cell (RT_2 and up, BDM_2 and up): opt-in
        // (NFA_HDIV_MMA=1).  Measured on B200 it is SLOWER than the register-tiled DFMA tiles: RT_2 36.6 vs 24.5 ms,
        // RT_3 38.8 vs 33.2 ms, and the same with four independent accumulation chains per warp (37.3 / 38.6 ms): the
        // DMMA rate itself is the limit, not the latency of the dependent sequence.
        const char *e_mma = getenv("NFA_HDIV_MMA");
        const bool mma = G >= 32 && e_mma && atoi(e_mma) != 0;
        const int n_up = mma ? ((n_u + 2 + 7) & ~7) : nt * T;
        const int n_rows = mma ? ((n_q * dim + 3) & ~3) : n_q * dim;
        const size_t group_doubles = (size_t)n_rows * n_up + (size_t)n_q * n_up + 2 * (n_q + 2) + (size_t)(dim + 1) * nfd * nfd;
        const size_t tsmem = sizeof(double) * group_doubles * (128 / G);
        if (tsmem <= 200 * 1024 && (G == 8 || G == 16 || G == 32 || G == 128)) {
            int device = 0, n_sm = 148, per_sm = 1;
            NFA_CUDA(cudaGetDevice(&device));
            NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
#define NFA_HT_LAUNCH(D, GG)                                                                                         \
    {                                                                                                                \
        auto kern = (mma && GG >= 32) ? hdiv_l2_tiled_kernel<D, GG, T, (GG >= 32)> : hdiv_l2_tiled_kernel<D, GG, T, false>; \
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));               \
        NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, tsmem));                          \
        const int64_t want = (n_cells + (128 / GG) - 1) / (128 / GG);                                                \
        const int tgrid = (int)std::min<int64_t>(want, (int64_t)n_sm * std::max(per_sm, 1));                         \
        kern<<<tgrid, 128, tsmem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,      \
                                       cell_nodes, n_cells, kp, f, f_u, Ke, re);                                     \
    }
            if (dim == 3) {
                if (G == 8) NFA_HT_LAUNCH(3, 8) else if (G == 16) NFA_HT_LAUNCH(3, 16) else if (G == 32) NFA_HT_LAUNCH(3, 32) else NFA_HT_LAUNCH(3, 128)
            } else if (dim == 2) {
                if (G == 8) NFA_HT_LAUNCH(2, 8) else if (G == 16) NFA_HT_LAUNCH(2, 16) else if (G == 32) NFA_HT_LAUNCH(2, 32) else NFA_HT_LAUNCH(2, 128)
            } else
                return fail("nfa_local_hdiv_l2: dim must be 2 or 3");
#undef NFA_HT_LAUNCH
            NFA_LAUNCH_CHECK();
            return 0;
        }
    }
    if (dim == 3) {
        auto kern = local_hdiv_l2_kernel<3>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,
                                     cell_nodes, n_cells, kp, f, f_u, Ke, re);
    } else if (dim == 2) {
        auto kern = local_hdiv_l2_kernel<2>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,
                                     cell_nodes, n_cells, kp, f, f_u, Ke, re);
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

extern "C" int nfa_local_sundus(int dim, int n_u, int n_p, int n_q, int nfd, const double *u_val, const double *u_div,
                                const double *p_val, const double *weights, const double *Tcases, const int8_t *cases,
                                const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *kappa,
                                int64_t ksc, int64_t ksq, const double *delta, int64_t dsc, int64_t dsq,
                                const double *eta, int64_t esc, int64_t esq, double delta_t, const double *f_f,
                                const double *f_r, const double *alpha, const double *alpha_prev, const int32_t *dest,
                                double *Ke, double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!dest) return fail("nfa_local_sundus: dest is required");
    if (delta_t == 0.0) return fail("nfa_local_sundus: delta_t is zero");
    cudaStream_t s = (cudaStream_t)stream;
    CoefH kp{kappa, ksc, ksq}, dl{delta, dsc, dsq}, et{eta, esc, esq};
    int n_loc = 2 * n_u + 2 * n_p;
    size_t smem = sizeof(double) * ((size_t)n_q * n_u * (dim + 1) + (size_t)(dim + 1) * nfd * nfd + 2 * n_loc +
                                    (size_t)n_q * (8 + 2 * dim) + 2 * n_q);
    if (smem > 200 * 1024) return fail("nfa_local_sundus: element too large for the shared-memory staging");
    int grid = (int)(n_cells < 148 * 8 ? n_cells : 148 * 8);
    const char *e_tiled = getenv("NFA_HDIV_TILED");
    if ((!e_tiled || atoi(e_tiled) != 0) && nfd <= HT_MAXFD) {
        constexpr int T = 3;
        const int nt = (n_u + T - 1) / T, n_up = nt * T, heavy = nt * (nt + 1) / 2 + nt * n_p + n_p * n_p;
        int G = heavy <= 8 ? 8 : (heavy <= 16 ? 16 : (heavy <= 48 ? 32 : 128));
        if (const char *e = getenv("NFA_HDIV_LANES")) G = atoi(e);
        const size_t group_doubles = (size_t)n_q * dim * n_up + (size_t)n_q * n_up + (size_t)(dim + 1) * nfd * nfd +
                                     2 * n_loc + (size_t)n_q * (8 + 2 * dim) + 2 * n_q;
        const size_t tsmem = sizeof(double) * group_doubles * (128 / G);
        if (tsmem <= 200 * 1024 && (G == 8 || G == 16 || G == 32 || G == 128)) {
            int device = 0, n_sm = 148, per_sm = 1;
            NFA_CUDA(cudaGetDevice(&device));
            NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
#define NFA_ST_LAUNCH(D, GG)                                                                                          \
    {                                                                                                                 \
        auto kern = sundus_tiled_kernel<D, GG, T>;                                                                    \
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));                \
        NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, tsmem));                           \
        const int64_t want = (n_cells + (128 / GG) - 1) / (128 / GG);                                                 \
        const int tgrid = (int)std::min<int64_t>(want, (int64_t)n_sm * std::max(per_sm, 1));                          \
        kern<<<tgrid, 128, tsmem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords,       \
                                       cell_nodes, n_cells, kp, dl, et, 1.0 / delta_t, f_f, f_r, alpha, alpha_prev,   \
                                       dest, Ke, re);                                                                 \
    }
            if (dim == 3) {
                if (G == 8) NFA_ST_LAUNCH(3, 8) else if (G == 16) NFA_ST_LAUNCH(3, 16) else if (G == 32) NFA_ST_LAUNCH(3, 32) else NFA_ST_LAUNCH(3, 128)
            } else if (dim == 2) {
                if (G == 8) NFA_ST_LAUNCH(2, 8) else if (G == 16) NFA_ST_LAUNCH(2, 16) else if (G == 32) NFA_ST_LAUNCH(2, 32) else NFA_ST_LAUNCH(2, 128)
            } else
                return fail("nfa_local_sundus: dim must be 2 or 3");
#undef NFA_ST_LAUNCH
            NFA_LAUNCH_CHECK();
            return 0;
        }
    }
    if (dim == 3) {
        auto kern = local_sundus_kernel<3>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords, cell_nodes,
                                     n_cells, kp, dl, et, 1.0 / delta_t, f_f, f_r, alpha, alpha_prev, dest, Ke, re);
    } else if (dim == 2) {
        auto kern = local_sundus_kernel<2>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords, cell_nodes,
                                     n_cells, kp, dl, et, 1.0 / delta_t, f_f, f_r, alpha, alpha_prev, dest, Ke, re);
    } else
        return fail("nfa_local_sundus: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_hdiv_l2_general(int dim, int n_u, int n_p, int n_q, int nfd, const double *u_val,
                                         const double *u_div, const double *p_val, const double *weights,
                                         const double *Tcases, const int8_t *cases, const double *coords,
                                         const int32_t *cell_nodes, int64_t n_cells, const double *c_uu,
                                         const double *c_d, const double *gvec, const double *c_pp,
                                         const double *load_p, const double *alpha, const double *beta,
                                         const int32_t *dest, const double *s_u, const double *s_p, double *Ke,
                                         double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!c_uu) return fail("nfa_local_hdiv_l2_general: c_uu is required");
    if ((alpha || beta) && !dest) return fail("nfa_local_hdiv_l2_general: alpha/beta given without dest");
    cudaStream_t s = (cudaStream_t)stream;
    int n_loc = n_u + n_p;
    size_t smem = sizeof(double) * ((size_t)n_q * n_u * (dim + 1) + n_q + (size_t)(dim + 1) * nfd * nfd + 2 * n_loc +
                                    (size_t)n_q * (dim + 1));
    if (smem > 200 * 1024) return fail("nfa_local_hdiv_l2_general: element too large for the shared-memory staging");
    int grid = (int)(n_cells < 148 * 8 ? n_cells : 148 * 8);
    if (dim == 3) {
        auto kern = local_hdiv_general_kernel<3>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords, cell_nodes,
                                     n_cells, c_uu, c_d, gvec, c_pp, load_p, alpha, beta, dest, s_u, s_p, Ke, re);
    } else if (dim == 2) {
        auto kern = local_hdiv_general_kernel<2>;
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(n_u, n_p, n_q, nfd, u_val, u_div, p_val, weights, Tcases, cases, coords, cell_nodes,
                                     n_cells, c_uu, c_d, gvec, c_pp, load_p, alpha, beta, dest, s_u, s_p, Ke, re);
    } else
        return fail("nfa_local_hdiv_l2_general: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}


extern "C" int nfa_local_hdiv_facet_vec(int dim, int nfd, int n_q, int n_comp, const double *trace_tab,
                                        const int32_t *tidx, const double *weights, const double *Tcases,
                                        const int8_t *cases, const double *coords, const int32_t *facet_nodes,
                                        const int32_t *neigh_nodes, const double *normals, const double *g,
                                        int64_t n_facets, double sign, int re_off, int re_stride, double *re,
                                        nfa_stream stream) {
    if (n_facets == 0) return 0;
    if (n_comp < 1 || re_stride < re_off + nfd * n_comp) return fail("nfa_local_hdiv_facet_vec: bad output layout");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t ctas = (n_facets + 3) / 4;
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (dim == 3)
        local_hdiv_facet_kernel<3><<<grid, 128, 0, s>>>(nfd, n_q, trace_tab, tidx, weights, Tcases, cases, coords,
                                                        facet_nodes, neigh_nodes, normals, g, n_facets, re, n_comp, sign,
                                                        re_off, re_stride);
    else if (dim == 2)
        local_hdiv_facet_kernel<2><<<grid, 128, 0, s>>>(nfd, n_q, trace_tab, tidx, weights, Tcases, cases, coords,
                                                        facet_nodes, neigh_nodes, normals, g, n_facets, re, n_comp, sign,
                                                        re_off, re_stride);
    else
        return fail("nfa_local_hdiv_facet_vec: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_hdiv_facet_penalty(int dim, int nfd, int n_q, int n_comp, const double *trace_tab,
                                            const int32_t *tidx, const double *weights, const double *Tcases,
                                            const int8_t *cases, const double *coords, const int32_t *facet_nodes,
                                            const int32_t *neigh_nodes, const double *normals, const double *g,
                                            int64_t n_facets, double beta, int off, int n_loc, double *Ke, double *re,
                                            nfa_stream stream) {
    if (n_facets == 0) return 0;
    if (n_comp < 1 || n_loc < off + nfd * n_comp) return fail("nfa_local_hdiv_facet_penalty: bad output layout");
    if (!Ke || !re || !g) return fail("nfa_local_hdiv_facet_penalty: NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ctas = (n_facets + 3) / 4;
    const int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    const size_t smem = sizeof(double) * 4 * ((size_t)nfd * n_q + n_q);
    if (smem > 48 * 1024) return fail("nfa_local_hdiv_facet_penalty: facet family too large");
    if (dim == 3)
        local_hdiv_facet_penalty_kernel<3><<<grid, 128, smem, s>>>(nfd, n_q, n_comp, trace_tab, tidx, weights, Tcases, cases,
                                                                   coords, facet_nodes, neigh_nodes, normals, g, n_facets,
                                                                   beta, off, n_loc, Ke, re);
    else if (dim == 2)
        local_hdiv_facet_penalty_kernel<2><<<grid, 128, smem, s>>>(nfd, n_q, n_comp, trace_tab, tidx, weights, Tcases, cases,
                                                                   coords, facet_nodes, neigh_nodes, normals, g, n_facets,
                                                                   beta, off, n_loc, Ke, re);
    else
        return fail("nfa_local_hdiv_facet_penalty: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

// ---- error norms of a discrete field (R/src/postprocess/l2_error_post_processor.py:7-63 l2_error, :185-237 div_error)
// One warp per cell, one lane per quadrature point: the discrete field (or its divergence) is evaluated at the point
// from the global vector alpha through the cell's destination indices and compared with host-evaluated exact values.
//   kind 0  Lagrange, n_comp components:      err = sum_q w det sum_c (e_c - sum_a phi_a alpha_{a,c})^2
//   kind 1  H(div) values, rows c of a tensor: err = sum_q w det sum_{c,d} (e_{c,d} - sum_a psi_a[d] alpha_{a,c})^2
//   kind 2  H(div) divergence:                 err = sum_q w det sum_c (e_c - sum_a div psi_a alpha_{a,c})^2
// out[cell] = the cell's contribution (summed by the caller in cell order: deterministic).
template <int DIM>
__global__ void __launch_bounds__(128)
    field_error_kernel(int kind, int n_phi, int n_comp, int n_q, int nfd, const double *__restrict__ val,
                       const double *__restrict__ dv, const double *__restrict__ weights,
                       const double *__restrict__ Tcases, const int8_t *__restrict__ cases,
                       const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                       const double *__restrict__ alpha, const int32_t *__restrict__ dest, int dest_stride,
                       int dest_off, const double *__restrict__ exact, const double *__restrict__ scale,
                       const double *__restrict__ gscale, double *__restrict__ out) {
    constexpr int NF = DIM + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t cell = (int64_t)blockIdx.x * 4 + warp; cell < n_cells; cell += (int64_t)gridDim.x * 4) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        const double idet = 1.0 / g.det;
        const int32_t *dd = dest + cell * dest_stride + dest_off;
        const int n_e = kind == 1 ? n_comp * DIM : n_comp;  // exact values per point
        double err = 0.0;
        for (int q = lane; q < n_q; q += 32) {
            double loc = 0.0;
            for (int c = 0; c < n_comp; ++c) {
                double uh[DIM];  // kind 1: the row c of the tensor; else only uh[0]
#pragma unroll
                for (int d = 0; d < DIM; ++d) uh[d] = 0.0;
                for (int a = 0; a < n_phi; ++a) {
                    const double al = alpha[dd[a * n_comp + c]];
                    if (kind == 0) {
                        uh[0] += val[q * n_phi + a] * al;
                        continue;
                    }
                    double vh[DIM], dh = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) vh[k] = 0.0;
                    if (a < NF * nfd) {
                        const int fc = a / nfd, ai = a - fc * nfd;
                        const double *T = Tcases + (int)cases[cell * NF + fc] * nfd * nfd + ai * nfd;
                        for (int b = 0; b < nfd; ++b) {
                            const double *src = val + ((int64_t)q * n_phi + fc * nfd + b) * DIM;
#pragma unroll
                            for (int k = 0; k < DIM; ++k) vh[k] += T[b] * src[k];
                            dh += T[b] * dv[q * n_phi + fc * nfd + b];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < DIM; ++k) vh[k] = val[((int64_t)q * n_phi + a) * DIM + k];
                        dh = dv[q * n_phi + a];
                    }
                    if (kind == 2) {
                        uh[0] += dh * idet * al;
                    } else if (kind == 3) {  // scaled divergence: l div psi + grad l . psi
                        double D = scale[cell * n_q + q] * dh * idet;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) {
                            double v = 0.0;
#pragma unroll
                            for (int k = 0; k < DIM; ++k) v += g.J[d][k] * vh[k];
                            D += gscale[(cell * n_q + q) * DIM + d] * v * idet;
                        }
                        uh[0] += D * al;
                    } else {
#pragma unroll
                        for (int d = 0; d < DIM; ++d) {
                            double v = 0.0;
#pragma unroll
                            for (int k = 0; k < DIM; ++k) v += g.J[d][k] * vh[k];
                            uh[d] += v * idet * al;
                        }
                    }
                }
                const double *ex = exact + ((int64_t)cell * n_q + q) * n_e;
                if (kind == 1) {
#pragma unroll
                    for (int d = 0; d < DIM; ++d) {
                        const double df = ex[c * DIM + d] - uh[d];
                        loc += df * df;
                    }
                } else {
                    const double df = ex[c] - uh[0];
                    loc += df * df;
                }
            }
            err += weights[q] * g.det * loc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if (lane == 0) out[cell] = err;
    }
}

extern "C" int nfa_field_error(int dim, int kind, int n_phi, int n_comp, int n_q, int nfd, const double *val,
                               const double *dv, const double *weights, const double *Tcases, const int8_t *cases,
                               const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *alpha,
                               const int32_t *dest, int dest_stride, int dest_off, const double *exact, double *out,
                               nfa_stream stream) {
    return nfa_field_error_scaled(dim, kind, n_phi, n_comp, n_q, nfd, val, dv, weights, Tcases, cases, coords, cell_nodes,
                                  n_cells, alpha, dest, dest_stride, dest_off, exact, nullptr, nullptr, out, stream);
}

extern "C" int nfa_field_error_scaled(int dim, int kind, int n_phi, int n_comp, int n_q, int nfd, const double *val,
                                      const double *dv, const double *weights, const double *Tcases, const int8_t *cases,
                                      const double *coords, const int32_t *cell_nodes, int64_t n_cells,
                                      const double *alpha, const int32_t *dest, int dest_stride, int dest_off,
                                      const double *exact, const double *scale, const double *gscale, double *out,
                                      nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (kind < 0 || kind > 3)
        return fail("nfa_field_error: kind must be 0 (Lagrange), 1 (H(div) value), 2 (H(div) divergence) or 3 (scaled divergence)");
    if (kind == 3 && (!scale || !gscale)) return fail("nfa_field_error_scaled: scale / grad scale arrays missing");
    if (kind != 0 && (!dv || !Tcases || !cases)) return fail("nfa_field_error: H(div) tables missing");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t ctas = (n_cells + 3) / 4;
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
#define NFA_FE(D)                                                                                              \
    field_error_kernel<D><<<grid, 128, 0, s>>>(kind, n_phi, n_comp, n_q, nfd, val, dv, weights, Tcases, cases, \
                                               coords, cell_nodes, n_cells, alpha, dest, dest_stride, dest_off, exact, \
                                               scale, gscale, out)
    if (dim == 3)
        NFA_FE(3);
    else if (dim == 2)
        NFA_FE(2);
    else if (dim == 1 && kind == 0)
        NFA_FE(1);
    else
        return fail("nfa_field_error: unsupported dimension");
#undef NFA_FE
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_hdiv_coupling(int dim, int nfd, int n_p, int n_q, const double *trace_tab, const double *weights,
                                       const double *Tcases, const double *coords, const int32_t *tidx_a,
                                       const int8_t *cases_a, const int32_t *facet_a, const int32_t *neigh_a,
                                       const double *normals_a, const int32_t *tidx_b, const int8_t *cases_b,
                                       const int32_t *facet_b, const int32_t *neigh_b, const double *normals_b,
                                       const double *p_val, const double *ainv, int64_t n_items, const double *alpha,
                                       const int32_t *dest, double *Ke, double *re, nfa_stream stream) {
    if (n_items == 0) return 0;
    if (!trace_tab || !weights || !Tcases || !p_val || !ainv || !re) return fail("nfa_local_hdiv_coupling: NULL argument");
    if (alpha && !dest) return fail("nfa_local_hdiv_coupling: a state vector needs the destination table");
    cudaStream_t s = (cudaStream_t)stream;
    const int n_loc = 2 * nfd + n_p;
    const size_t smem = sizeof(double) * 4 * ((size_t)(2 * nfd + 1) * n_q + n_loc);
    if (smem > 48 * 1024) return fail("nfa_local_hdiv_coupling: element too large for the shared-memory staging");
    const int64_t ctas = (n_items + 3) / 4;
    const int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (dim == 3)
        local_hdiv_coupling_kernel<3><<<grid, 128, smem, s>>>(nfd, n_p, n_q, trace_tab, weights, Tcases, coords, tidx_a, cases_a,
                                                             facet_a, neigh_a, normals_a, tidx_b, cases_b, facet_b, neigh_b,
                                                             normals_b, p_val, ainv, n_items, alpha, dest, Ke, re);
    else if (dim == 2)
        local_hdiv_coupling_kernel<2><<<grid, 128, smem, s>>>(nfd, n_p, n_q, trace_tab, weights, Tcases, coords, tidx_a, cases_a,
                                                             facet_a, neigh_a, normals_a, tidx_b, cases_b, facet_b, neigh_b,
                                                             normals_b, p_val, ainv, n_items, alpha, dest, Ke, re);
    else
        return fail("nfa_local_hdiv_coupling: dim must be 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}
