// libnfa: lowest-order H(div) x L2 forms assembled by ROW OWNER, without local matrices in HBM (sm_100a).
//
// Forms: LaplaceDualWeakForm on RT_1 x P_0 (BASELINE config 3 at k = 1, R/src/weak_forms/laplace_dual_weak_form.py:10-87)
// and SundusDualWeakForm on RT_1 x RT_1 x P_0 x P_0 (config 5, R/examples/parabolic_diffusive_problem/SundusWeakForm.py:
// 11-177).  The two-stage path writes a dense n_loc x n_loc block per cell (800 B for the 10 x 10 Sundus block, most of it
// structural zeros) and reads it back through a sorted 32-bit id per entry; its local kernel gives a cell to 8 lanes that
// diverge over four kinds of work.  Here
//   1. rt1_records_kernel: ONE THREAD PER CELL, every accumulator in registers, no divergence: the cell's distinct numbers
//      -- per flux row a: M[a][0..n_u), C[a], r[a]; then the L2 diagonal entries and residuals -- form a record of
//      (n_flux n_u (n_u + 2) + 4) doubles (416 B instead of 800 + 80), transposed through shared memory so that the
//      records leave the SM as full lines;
//   2. rows_expand_kernel: ONE THREAD PER MATRIX ROW.  In these spaces a dof belongs to at most two cells; the row's thread
//      reads the record runs of its one or two (cell, local row) sources (16-byte loads), places the entries at their CSR
//      slots (uint8 per (row, source, run entry), found once by nfa_rows_locate in the plan's CSR, whatever other forms
//      added to the pattern) inside a shared-memory image of 256 consecutive CSR rows, lower cell first -- the order of the reference's
//      Mat.setValue(..., addv=True) calls (R/src/assembly/ScatterFormData.py:5-11) -- and the block writes the image to
//      HBM with coalesced stores.  Every CSR value of the rows is written exactly once (explicit zeros included), the
//      right-hand side row is the sum of its one or two record entries: no atomics, no zero fill, bit-reproducible.
// Which record entries feed local row a is a table supplied by the host form (roff[a] = start of the row's run of up to 8
// record doubles, codes[a][k] = what entry k of the run is: a local column with its sign, the residual, or nothing), so
// rows_expand_kernel is form-independent.
#include <algorithm>

#include "nfa_common.cuh"

using namespace nfa;

namespace {

struct CoefR {
    const double *p;
    int64_t sc, sq;
};

constexpr int REC_CELLS = 128;  // cells per CTA pass of the records kernel
constexpr int ROWS_PER_CTA = 256;

// FORM 0: mixed Poisson [u | p]; FORM 1: Sundus [mp | mc | p | c]; FORM 2: generalised mixed elliptic form [u | p]
// (HdivL2GeneralForm: kappa = c_uu multiplies the flux mass, delta = c_d scales the divergence, eta = c_pp is the weight of
// the (p, p) mass, f_f = load_p, f_r = g [cell][q][DIM] adds g . psi to the coupling; absent arrays: 1, 0, 0, 0)
template <int DIM, int FORM>
__global__ void __launch_bounds__(REC_CELLS, 3)
    rt1_records_kernel(int n_q, const double *__restrict__ u_val, const double *__restrict__ u_div,
                       const double *__restrict__ p_val, const double *__restrict__ weights,
                       const double *__restrict__ Tcases, const int8_t *__restrict__ cases,
                       const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                       CoefR kappa, CoefR delta, CoefR eta, double inv_dt, const double *__restrict__ f_f,
                       const double *__restrict__ f_r, const double *__restrict__ alpha,
                       const double *__restrict__ alpha_prev, const int32_t *__restrict__ dest, int stride,
                       double *__restrict__ rec) {
    constexpr int NU = DIM + 1, NS = NU * (NU + 1) / 2, NFLUX = FORM == 1 ? 2 : 1;
    constexpr int NLOC = NFLUX * NU + NFLUX;
    constexpr int ROWW = (NU + 3) & ~1;      // M[a][0..NU), C[a], r[a] (+ one pad in 2-D: rows start 16-byte aligned)
    constexpr int SOFF = NFLUX * NU * ROWW;  // scalar rows' block: FORM 0: r_p, C[.]; 1: K_pp, K_cc, r_p, r_c, C[.]; 2: K_pp, r_p, C[.]
    constexpr int NSC = FORM == 1 ? 4 : (FORM == 2 ? 2 : 1);
    constexpr int USED = SOFF + NSC + NU;
    extern __shared__ double sm[];
    double *s_val = sm;                       // [n_q][NU][DIM]
    double *s_div = s_val + n_q * NU * DIM;   // [n_q][NU]
    double *s_pw = s_div + n_q * NU;          // [n_q]
    double *s_w = s_pw + n_q;                 // [n_q]
    double *s_rec = s_w + n_q;                // [REC_CELLS][pitch]
    const int pitch = stride | 1;             // odd pitch: the threads' record rows start in distinct banks
    const int tid = threadIdx.x;
    for (int i = tid; i < n_q * NU * DIM; i += REC_CELLS) s_val[i] = u_val[i];
    for (int i = tid; i < n_q * NU; i += REC_CELLS) s_div[i] = u_div[i];
    for (int i = tid; i < n_q; i += REC_CELLS) {
        s_pw[i] = p_val[i];
        s_w[i] = weights[i];
    }
    __syncthreads();
    for (int64_t c0 = (int64_t)blockIdx.x * REC_CELLS; c0 < n_cells; c0 += (int64_t)gridDim.x * REC_CELLS) {
        const int64_t cell = c0 + tid;
        double *R = s_rec + tid * pitch;
        if (cell < n_cells) {
            double X[DIM + 1][3];
            load_cell_coords<DIM>(coords, cell_nodes, cell, X);
            Geom<DIM> g;
            compute_geom<DIM>(X, g);
            const double det = g.det, idet = 1.0 / g.det;
            double sgn[NU];
#pragma unroll
            for (int a = 0; a < NU; ++a) sgn[a] = __ldg(Tcases + (int)cases[cell * NU + a]);  // 1 x 1 orientation matrices
            // state of the cell (Sundus): the scalar fields' local coefficients now (the reaction term needs c_h inside the
            // point loop), the flux coefficients after the loop (registers)
            double a_s[NFLUX], an_s[NFLUX];
#pragma unroll
            for (int F = 0; F < NFLUX; ++F) a_s[F] = an_s[F] = 0.0;
            if (FORM >= 1 && (alpha || alpha_prev)) {
                const int32_t *d = dest + cell * NLOC;
#pragma unroll
                for (int F = 0; F < NFLUX; ++F) {
                    if (alpha) a_s[F] = __ldg(alpha + d[NFLUX * NU + F]);
                    if (alpha_prev) an_s[F] = __ldg(alpha_prev + d[NFLUX * NU + F]);
                }
            }
            double M[NFLUX][NS], C[NU];
#pragma unroll
            for (int F = 0; F < NFLUX; ++F)
#pragma unroll
                for (int t = 0; t < NS; ++t) M[F][t] = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) C[a] = 0.0;
            double W2 = 0.0, Wc = 0.0, Ff = 0.0, Rc = 0.0;
            // coefficient values of point q + 1 are requested before the arithmetic of point q
            double kq = kappa.p[cell * kappa.sc], dq = 0.0, eq = 0.0, ffq = 0.0, frq = 0.0;
            if (FORM == 1) {
                dq = delta.p[cell * delta.sc];
                eq = eta.p[cell * eta.sc];
                frq = f_r ? f_r[cell * n_q] : 0.0;
            } else if (FORM == 2) {
                dq = delta.p ? delta.p[cell * delta.sc] : 1.0;
                eq = eta.p ? eta.p[cell * eta.sc] : 0.0;
            }
            ffq = f_f ? f_f[cell * n_q] : 0.0;
            for (int q = 0; q < n_q; ++q) {
                const double wd = s_w[q] * det;
                const double k_now = kq, d_now = dq, e_now = eq, ff_now = ffq, fr_now = frq;
                if (q + 1 < n_q) {
                    kq = kappa.p[cell * kappa.sc + (q + 1) * kappa.sq];
                    if (FORM == 1) {
                        dq = delta.p[cell * delta.sc + (q + 1) * delta.sq];
                        eq = eta.p[cell * eta.sc + (q + 1) * eta.sq];
                        if (f_r) frq = f_r[cell * n_q + q + 1];
                    } else if (FORM == 2) {
                        if (delta.p) dq = delta.p[cell * delta.sc + (q + 1) * delta.sq];
                        if (eta.p) eq = eta.p[cell * eta.sc + (q + 1) * eta.sq];
                    }
                    if (f_f) ffq = f_f[cell * n_q + q + 1];
                }
                double wt[NFLUX];
                wt[0] = FORM == 2 ? wd * k_now : wd * (1.0 / k_now);
                if (FORM == 1) wt[NFLUX - 1] = wd * (1.0 / d_now);
                double psi[NU][DIM];
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        double v = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) v += g.J[c][k] * s_val[(q * NU + a) * DIM + k];
                        psi[a][c] = (v * idet) * sgn[a];
                    }
                int t = 0;
#pragma unroll
                for (int a = 0; a < NU; ++a)
#pragma unroll
                    for (int b = a; b < NU; ++b) {
                        double dot = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) dot += psi[a][c] * psi[b][c];
#pragma unroll
                        for (int F = 0; F < NFLUX; ++F) M[F][t] = fma(wt[F], dot, M[F][t]);
                        ++t;
                    }
                const double pw = s_pw[q], wp = wd * pw;
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double D = (s_div[q * NU + a] * idet) * sgn[a];
                    if (FORM == 2) {
                        D *= d_now;
                        if (f_r) {
#pragma unroll
                            for (int c = 0; c < DIM; ++c) D = fma(f_r[(cell * n_q + q) * DIM + c], psi[a][c], D);
                        }
                    }
                    C[a] = fma(wp, D, C[a]);
                }
                W2 = fma(wp, FORM == 2 ? pw * e_now : pw, W2);
                if (FORM == 1) {
                    const double et = e_now, ch = a_s[NFLUX - 1] * pw;
                    Wc = fma(wp * pw, et * ch, Wc);
                    Ff = fma(wp, ff_now, Ff);
                    Rc = fma(wp, 1.0 + et * ch * ch - fr_now, Rc);
                } else {
                    Ff = fma(wp, ff_now, Ff);
                }
            }
            double a_f[NFLUX][NU];
#pragma unroll
            for (int F = 0; F < NFLUX; ++F)
#pragma unroll
                for (int a = 0; a < NU; ++a) a_f[F][a] = (FORM >= 1 && alpha) ? __ldg(alpha + dest[cell * NLOC + F * NU + a]) : 0.0;
            // record rows of the flux dofs: M[a][.], C[a], r[a]
#pragma unroll
            for (int F = 0; F < NFLUX; ++F)
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double r = 0.0;
#pragma unroll
                    for (int b = 0; b < NU; ++b) {
                        const int lo = a < b ? a : b, hi = a < b ? b : a;
                        const double m = M[F][lo * NU - lo * (lo - 1) / 2 + (hi - lo)];
                        R[(F * NU + a) * ROWW + b] = m;
                        r = fma(m, a_f[F][b], r);
                    }
                    R[(F * NU + a) * ROWW + NU] = C[a];
                    R[(F * NU + a) * ROWW + NU + 1] = FORM >= 1 ? r - a_s[F] * C[a] : 0.0;
                    if (ROWW > NU + 2) R[(F * NU + a) * ROWW + NU + 2] = 0.0;
                }
            if (FORM == 1) {
                const double app = inv_dt * W2;
                double dmp = 0.0, dmc = 0.0;
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    dmp = fma(a_f[0][a], C[a], dmp);
                    dmc = fma(a_f[NFLUX - 1][a], C[a], dmc);
                }
                R[SOFF + 0] = app;
                R[SOFF + 1] = app + 2.0 * Wc;
                R[SOFF + 2] = (a_s[0] - an_s[0]) * app + dmp - Ff;
                R[SOFF + 3] = (a_s[NFLUX - 1] - an_s[NFLUX - 1]) * app + dmc + Rc;
            } else if (FORM == 2) {
                double da = 0.0;
#pragma unroll
                for (int a = 0; a < NU; ++a) da = fma(a_f[0][a], C[a], da);
                R[SOFF + 0] = W2;  // int c_pp w w
                R[SOFF + 1] = da + W2 * a_s[0] - Ff;
            } else {
                R[SOFF] = -Ff;
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) R[SOFF + NSC + a] = C[a];  // the scalar rows read their couplings here
            for (int k = USED; k < stride; ++k) R[k] = 0.0;
        }
        __syncthreads();
        const int64_t left = n_cells - c0, live = left < REC_CELLS ? left : REC_CELLS;
        double *out = rec + c0 * stride;
        for (int i = tid; i < (int)live * stride; i += REC_CELLS) {
            const int c = i / stride, k = i - c * stride;
            out[i] = s_rec[c * pitch + k];
        }
        __syncthreads();
    }
}

// Runs: local row a of a source reads the record doubles [roff[a], roff[a] + 8) (16-byte aligned; as many as the row's
// codes use); codes[a][k]: 0 = not used by this row, +-(b + 1) = local column b with that sign, 127 = the row's residual.
constexpr int RUN = 8;
constexpr int CODE_RHS = 127;

// CSR slot of every run entry of a row's one or two sources inside the row's segment (binary search; one-off)
__global__ void rows_locate_kernel(int64_t n_rows, const int32_t *__restrict__ rows, const int32_t *__restrict__ indptr,
                                   const int32_t *__restrict__ indices, const int32_t *__restrict__ row_src, int n_loc,
                                   const int8_t *__restrict__ codes, const int32_t *__restrict__ dest,
                                   uint8_t *__restrict__ slots, int *__restrict__ flag) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * n_rows) return;
    const int32_t row = rows[id >> 1];
    const int32_t src = row_src[id];
    uint8_t *out = slots + id * RUN;
    for (int k = 0; k < RUN; ++k) out[k] = 0;
    if (src < 0) return;
    const int64_t c = src >> 5;
    const int a = src & 31;
    const int32_t b0 = indptr[row], b1 = indptr[row + 1];
    if (a >= n_loc || dest[c * n_loc + a] != row) {
        atomicOr(flag, 4);
        return;
    }
    for (int k = 0; k < RUN; ++k) {
        const int code = codes[a * RUN + k];
        if (code == 0 || code == CODE_RHS) continue;
        const int32_t col = dest[c * n_loc + (code > 0 ? code : -code) - 1];
        int32_t lo = b0, hi = b1;
        while (lo < hi) {
            const int32_t mid = (lo + hi) >> 1;
            if (indices[mid] < col)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo >= b1 || indices[lo] != col)
            atomicOr(flag, 1);
        else if (lo - b0 > 255)
            atomicOr(flag, 2);
        out[k] = (uint8_t)(lo - b0);
    }
}

__global__ void __launch_bounds__(ROWS_PER_CTA)
    rows_expand_kernel(int64_t n_blocks, const int4 *__restrict__ blocks, const int32_t *__restrict__ indptr,
                       const int32_t *__restrict__ row_src, const uint8_t *__restrict__ slots, int n_loc,
                       const int32_t *__restrict__ roff, const int8_t *__restrict__ codes, const double *__restrict__ rec,
                       int stride, double *__restrict__ vals, double *__restrict__ rhs, int rhs_add) {
    extern __shared__ double tile[];
    __shared__ int s_off[32];
    __shared__ int s_len[32];
    __shared__ int8_t s_code[32 * RUN];
    const int tid = threadIdx.x;
    for (int i = tid; i < n_loc * RUN; i += ROWS_PER_CTA) s_code[i] = codes[i];
    if (tid < n_loc) {
        s_off[tid] = roff[tid];
        int len = 0;
        for (int k = 0; k < RUN; ++k)
            if (codes[tid * RUN + k] != 0) len = k + 1;
        s_len[tid] = len;
    }
    __syncthreads();  // the run tables are read before the first barrier of the loop
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int4 bd = __ldg(blocks + blk);  // first row, rows (<= ROWS_PER_CTA, consecutive), first table entry
        const int64_t r0 = bd.x, r1 = r0 + bd.y;
        const int32_t base = indptr[r0];
        const int len = indptr[r1] - base;
        const int64_t r = r0 + tid;
        // the row's lists and record runs are requested before the image is cleared (single-use lists: streaming loads)
        int2 srcs = make_int2(-1, -1);
        uint4 sl = make_uint4(0, 0, 0, 0);
        int row_at = 0;
        double v[2][RUN];
        int a2[2] = {0, 0};
        if (r < r1) {
            const int64_t rl = (int64_t)bd.z + tid;
            srcs = __ldcs(reinterpret_cast<const int2 *>(row_src) + rl);
            sl = __ldcs(reinterpret_cast<const uint4 *>(slots) + rl);
            row_at = indptr[r] - base;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int32_t src = side == 0 ? srcs.x : srcs.y;
#pragma unroll
                for (int k = 0; k < RUN; ++k) v[side][k] = 0.0;
                if (src < 0) continue;
                const int a = src & 31;
                a2[side] = a;
                const double2 *R = reinterpret_cast<const double2 *>(rec + (int64_t)(src >> 5) * stride + s_off[a]);
                const int n2 = (s_len[a] + 1) >> 1;
#pragma unroll
                for (int k2 = 0; k2 < RUN / 2; ++k2)
                    if (k2 < n2) {
                        const double2 t = __ldg(R + k2);
                        v[side][2 * k2] = t.x;
                        v[side][2 * k2 + 1] = t.y;
                    }
            }
        }
        __syncthreads();  // the previous block's copy-out is done
        for (int i = tid; i < len; i += ROWS_PER_CTA) tile[i] = 0.0;
        __syncthreads();
        if (r < r1) {
            double *row = tile + row_at;
            double rv = 0.0;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int32_t src = side == 0 ? srcs.x : srcs.y;
                if (src < 0) continue;
                const int8_t *cd = s_code + a2[side] * RUN;
                const uint32_t w0 = side == 0 ? sl.x : sl.z, w1 = side == 0 ? sl.y : sl.w;
#pragma unroll
                for (int k = 0; k < RUN; ++k) {
                    const int code = cd[k];
                    if (code == 0) continue;
                    const double val = v[side][k];
                    if (code == CODE_RHS) {
                        rv += val;
                    } else {
                        const uint32_t slot = ((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 255u;
                        if (side == 0)
                            row[slot] = code > 0 ? val : -val;  // a slot is fed at most once per source
                        else
                            row[slot] += code > 0 ? val : -val;
                    }
                }
            }
            if (rhs) {
                if (rhs_add)
                    rhs[r] += rv;
                else
                    __stcs(rhs + r, rv);
            }
        }
        __syncthreads();
        double *out = vals + base;
        for (int i = tid; i < len; i += ROWS_PER_CTA) __stcs(out + i, tile[i]);  // written once, never read here
    }
}

}  // namespace

extern "C" int nfa_hdiv_rt1_records(int dim, int form, int n_q, const double *u_val, const double *u_div,
                                    const double *p_val, const double *weights, const double *Tcases, const int8_t *cases,
                                    const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *kappa,
                                    int64_t ksc, int64_t ksq, const double *delta, int64_t dsc, int64_t dsq,
                                    const double *eta, int64_t esc, int64_t esq, double delta_t, const double *f_f,
                                    const double *f_r, const double *alpha, const double *alpha_prev, const int32_t *dest,
                                    int stride, double *rec, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (dim != 2 && dim != 3) return fail("nfa_hdiv_rt1_records: dim must be 2 or 3");
    if (form < 0 || form > 2) return fail("nfa_hdiv_rt1_records: form must be 0 (mixed Poisson), 1 (Sundus) or 2 (general)");
    if (n_q < 1 || n_q > 64) return fail("nfa_hdiv_rt1_records: 1 to 64 quadrature points");
    if (!u_val || !u_div || !p_val || !weights || !Tcases || !cases || !coords || !cell_nodes || !kappa || !rec)
        return fail("nfa_hdiv_rt1_records: NULL argument");
    const int nu = dim + 1, nflux = form == 1 ? 2 : 1;
    const int used = nflux * nu * ((nu + 3) & ~1) + (form == 1 ? 4 : (form == 2 ? 2 : 1)) + nu;
    if (stride < used || (stride & 3)) return fail("nfa_hdiv_rt1_records: record stride too small or not a multiple of 4");
    if (form == 1) {
        if (!delta || !eta || !dest) return fail("nfa_hdiv_rt1_records: the Sundus form needs delta, eta and dest");
        if (delta_t == 0.0) return fail("nfa_hdiv_rt1_records: delta_t is zero");
    }
    cudaStream_t s = (cudaStream_t)stream;
    CoefR kp{kappa, ksc, ksq}, dl{delta, dsc, dsq}, et{eta, esc, esq};
    const size_t smem = sizeof(double) * ((size_t)n_q * nu * dim + (size_t)n_q * nu + 2 * n_q + (size_t)REC_CELLS * (stride | 1));
    int device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const int64_t want = (n_cells + REC_CELLS - 1) / REC_CELLS;
    const int grid = (int)std::min<int64_t>(want, (int64_t)n_sm * 8);
    const double inv_dt = form == 1 ? 1.0 / delta_t : 0.0;
    // form 2 (HdivL2GeneralForm): kappa = c_uu, delta = c_d, eta = c_pp (NULL allowed for the last two), f_f = load_p,
    // f_r = g [n_cells][n_q][dim]
#define NFA_REC_LAUNCH(D, F)                                                                                          \
    {                                                                                                                 \
        auto kern = rt1_records_kernel<D, F>;                                                                         \
        NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        kern<<<grid, REC_CELLS, smem, s>>>(n_q, u_val, u_div, p_val, weights, Tcases, cases, coords, cell_nodes, n_cells, \
                                           kp, dl, et, inv_dt, f_f, f_r, alpha, alpha_prev, dest, stride, rec);       \
    }
    if (form == 2 && alpha && !dest) return fail("nfa_hdiv_rt1_records: alpha given without dest");
    if (dim == 3) {
        if (form == 1) NFA_REC_LAUNCH(3, 1) else if (form == 2) NFA_REC_LAUNCH(3, 2) else NFA_REC_LAUNCH(3, 0)
    } else {
        if (form == 1) NFA_REC_LAUNCH(2, 1) else if (form == 2) NFA_REC_LAUNCH(2, 2) else NFA_REC_LAUNCH(2, 0)
    }
#undef NFA_REC_LAUNCH
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_rows_locate(int64_t n_rows, const int32_t *rows, const int32_t *indptr, const int32_t *indices,
                               const int32_t *row_src, int n_loc, const int8_t *codes, const int32_t *dest, uint8_t *slots,
                               nfa_stream stream) {
    if (n_rows <= 0) return 0;
    if (!rows || !indptr || !indices || !row_src || !codes || !dest || !slots) return fail("nfa_rows_locate: NULL argument");
    if (n_loc < 1 || n_loc > 32) return fail("nfa_rows_locate: 1 to 32 local dofs");
    cudaStream_t s = (cudaStream_t)stream;
    int *d_flag = nullptr, h_flag = 0;
    NFA_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    cudaMemsetAsync(d_flag, 0, sizeof(int), s);
    rows_locate_kernel<<<grid_for(2 * n_rows, 256), 256, 0, s>>>(n_rows, rows, indptr, indices, row_src, n_loc, codes, dest,
                                                                 slots, d_flag);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_flag);
    if (e != cudaSuccess) return fail(std::string("nfa_rows_locate: ") + cudaGetErrorString(e));
    if (h_flag & 4) return fail("nfa_rows_locate: a row's source does not list the row's dof");
    if (h_flag & 1) return fail("nfa_rows_locate: a local column is missing from the row's CSR segment");
    if (h_flag & 2) return fail("nfa_rows_locate: a row holds more than 256 entries");
    return 0;
}

extern "C" int nfa_rows_expand(int64_t n_blocks, const int32_t *blocks, int max_row_len, const int32_t *indptr,
                               const int32_t *row_src, const uint8_t *slots, int n_loc, const int32_t *roff,
                               const int8_t *codes, const double *rec, int stride, double *vals, double *rhs, int rhs_add,
                               nfa_stream stream) {
    if (n_blocks <= 0) return 0;
    if (!blocks || !indptr || !row_src || !slots || !roff || !codes || !rec || !vals)
        return fail("nfa_rows_expand: NULL argument");
    if (n_loc < 1 || n_loc > 32) return fail("nfa_rows_expand: 1 to 32 local dofs");
    if (max_row_len < 1 || max_row_len > 256) return fail("nfa_rows_expand: rows of 1 to 256 entries");
    if (stride & 1) return fail("nfa_rows_expand: the record stride must be even (16-byte aligned runs)");
    if (((uintptr_t)rec & 15) || ((uintptr_t)slots & 15) || ((uintptr_t)row_src & 7) || ((uintptr_t)blocks & 15))
        return fail("nfa_rows_expand: records / slots / blocks must be 16-byte aligned, sources 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * (size_t)ROWS_PER_CTA * max_row_len;
    if (smem > 200 * 1024) return fail("nfa_rows_expand: rows too long for the shared-memory image");
    int device = 0, n_sm = 148, per_sm = 1;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    NFA_CUDA(cudaFuncSetAttribute(rows_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rows_expand_kernel, ROWS_PER_CTA, smem));
    const int grid = (int)std::min<int64_t>(n_blocks, (int64_t)n_sm * std::max(per_sm, 1));
    rows_expand_kernel<<<grid, ROWS_PER_CTA, smem, s>>>(n_blocks, reinterpret_cast<const int4 *>(blocks), indptr, row_src,
                                                        slots, n_loc, roff, codes, rec, stride, vals, rhs, rhs_add);
    NFA_LAUNCH_CHECK();
    return 0;
}
