// libnfa: row-item fused assembly of vector-P2 elasticity on affine tetrahedra (sm_100a) -- the config-2 headline path.
//
// The local matrix of LEPrimalWeakForm (R/src/weak_forms/le_primal_weak_form.py:91-126, integrated exactly for
// cell-wise constant lambda, mu) is written with the physical gradients g_m of the barycentric coordinates:
//   grad phi_vertex(i) = (4 lam_i - 1) g_i,   grad phi_edge(ij) = 4 (lam_j g_i + lam_i g_j),
// whose scalar factors integrate to the universal constants 1/10, -1/30, 4/15, 2/15.  A block of the local matrix is
// then Kf(sum c u (x) w) with u, w sums of the g_m of the vertices involved and
//   Kf(T) = det (lambda T + mu T^T + mu tr(T) I).
// With g' = sqrt(det mu) g and rho = lambda / mu:  Kf(u (x) w) = rho u' w'^T + w' u'^T + (u'.w') I.
//
// Work is organised by OUTPUT ROW (no atomics, no staging of contributions): the block row of a node is cut into
// ITEMS of at most three CSR blocks that receive contributions from exactly the same cells (nfempy_b200/assembly/
// p2_items.py lists the six item types).  One thread owns one item, walks its cells in ascending cell order -- the
// order of the reference's Mat.setValue(..., addv=True) calls, R/src/assembly/ScatterFormData.py:5-11 -- sums in
// registers and writes every CSR value exactly once into the CTA's shared-memory tile (the CSR image of a run of
// whole block rows), which leaves for HBM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, TMA engine).
// Inputs per (item, cell): one 32-bit packed entry (coalesced, ELL per warp) and one 32-byte record per vertex used,
// fetched with 256-bit loads (LDG.E.256) from the per-cell geometry records (L1/L2 resident: 128 B per cell).
#include <algorithm>
#include <cstdlib>

#include "nfa_plan.cuh"

using namespace nfa;

namespace {

#ifndef NFA_P2_THREADS
#define NFA_P2_THREADS 512
#endif
#ifndef P2_UNROLL_NG
#define P2_UNROLL_NG 2
#endif
constexpr int P2_THREADS = NFA_P2_THREADS;

// Experiment switches (tools/build_variant.sh builds a library per setting; the defaults are what bench.py measures).
//   NFA_P2_PF_ELL  1: the ELL lines of the NEXT task beyond its first two entries are prefetched into L2 (2: into L1) one
//                  task ahead -- a lane prefetches line `lane`; without it every round of a long item waits for a DRAM miss
//   NFA_P2_PF_IH   1: the item headers of the next task are prefetched into L1 one task ahead
//   NFA_P2_L2HINT  bit 0: geometry records are loaded L2::evict_last, bit 1: the tile leaves with L2::evict_first,
//                  bit 2: plan lists (ELL, headers) are loaded L2::evict_first
//   NFA_P2_E2      1: face items (type E: at most two cells) load both cells' records at once and finish block by block
#ifndef NFA_P2_PF_ELL
#define NFA_P2_PF_ELL 1
#endif
#ifndef NFA_P2_PF_IH
#define NFA_P2_PF_IH 1
#endif
#ifndef NFA_P2_L2HINT
#define NFA_P2_L2HINT 0
#endif
#ifndef NFA_P2_E2
#define NFA_P2_E2 1
#endif
//   NFA_P2_DEEP    1: task headers run three tasks ahead and the ELL lines + item headers of task i+2 are prefetched into
//                  L2, so that the first entries of task i+1 (loaded while task i computes) are L2 hits, not DRAM misses
//   NFA_P2_RELAXED 1: the per-task decrement of the tile's task counter is a relaxed shared-memory atomic; the warp that
//                  takes the counter to zero issues the acquire fence.  (The acq_rel form made every task wait for all of
//                  its outstanding global loads -- the look-ahead loads of the next tasks -- before it could retire.)
//   NFA_P2_GRAB2   1: the ticket of the task counter is taken one iteration before it is used (the shared-memory atomic's
//                  round trip hides behind a task instead of stalling the top of the loop)
#ifndef NFA_P2_DEEP
#define NFA_P2_DEEP 0  // measured: 0.872 ms against 0.845 -- three claimed tasks per warp overrun the four-tile window
#endif
#ifndef NFA_P2_GRAB2
#define NFA_P2_GRAB2 0  // measured: 0.870 ms against 0.845 (same reason: one more claimed task per warp)
#endif
#ifndef NFA_P2_RELAXED
#define NFA_P2_RELAXED 1
#endif
#ifndef NFA_P2_UA
#define NFA_P2_UA 8  // cells in flight, vertex-diagonal items (3 or more: the general loop)
#endif
#ifndef NFA_P2_UB
#define NFA_P2_UB 4  // cells in flight, vertex-edge items
#endif

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct G4 {
    double x, y, z, rho;
};

__device__ __forceinline__ G4 ldg256(const double *p) {
    G4 g;
#if (NFA_P2_L2HINT & 1)
    const uint64_t pol = l2_policy_evict_last();
    asm("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
        : "=d"(g.x), "=d"(g.y), "=d"(g.z), "=d"(g.rho)
        : "l"(p), "l"(pol));
    return g;
#endif
#ifdef NFA_P2_PLAIN_LD
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(g.x), "=d"(g.y), "=d"(g.z), "=d"(g.rho) : "l"(p));
#else
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(g.x), "=d"(g.y), "=d"(g.z), "=d"(g.rho) : "l"(p));
#endif
    return g;
}

// plan lists are read once per launch: optionally marked L2::evict_first so that they do not push the records out of L2
__device__ __forceinline__ uint32_t ld_list(const uint32_t *p) {
#if (NFA_P2_L2HINT & 4)
    uint32_t v;
    const uint64_t pol = l2_policy_evict_first();
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ int4 ld_list(const int4 *p) {
#if (NFA_P2_L2HINT & 4)
    int4 v;
    const uint64_t pol = l2_policy_evict_first();
    asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint2 ld_list(const uint2 *p) {
#if (NFA_P2_L2HINT & 4)
    uint2 v;
    const uint64_t pol = l2_policy_evict_first();
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}

// geometry records [n_cells][4][4]: (sqrt(det mu) g_m, lambda / mu), m = 0..3
__global__ void p2tet_geo_kernel(const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes,
                                 int64_t n_cells, const double *__restrict__ lam, int64_t ls,
                                 const double *__restrict__ mu, int64_t ms, const uint8_t *__restrict__ mask,
                                 double *__restrict__ rec) {
    int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    double X[4][3];
    load_cell_coords<3>(coords, cell_nodes, cell, X);
    Geom<3> g;
    compute_geom<3>(X, g);
    // mu > 0 is the caller's contract; cells with lambda = mu = 0 (the halo cells of a rank's local plan, whose
    // coefficients the sharded assembler zeroes) contribute exact zeros
    // cells switched off by the plan's mask (the halo cells of a rank's local plan when lambda, mu are constants: the
    // coefficients then stay single numbers and the constant-rho instantiation runs) contribute exact zeros as well
    // (zero gradient vectors; rho keeps its value, which the constant-rho instantiation reads from any record)
    const bool off = mask && !mask[cell];
    const double m = mu[cell * ms], l = lam[cell * ls];
    const double sc = (m > 0.0 && !off) ? sqrt(g.det * m) : 0.0, rho = m > 0.0 ? l / m : 0.0;
    double *r = rec + cell * 16;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double a = sc * g.Jinv[k][0], b = sc * g.Jinv[k][1], c = sc * g.Jinv[k][2];
        s0 -= a;
        s1 -= b;
        s2 -= c;
        reinterpret_cast<double2 *>(r + 4 * (k + 1))[0] = make_double2(a, b);
        reinterpret_cast<double2 *>(r + 4 * (k + 1))[1] = make_double2(c, rho);
    }
    reinterpret_cast<double2 *>(r)[0] = make_double2(s0, s1);
    reinterpret_cast<double2 *>(r)[1] = make_double2(s2, rho);
}

// acc += Kf(u (x) w) without the trace part; returns u.w.  CR (constant rho = lambda / mu over the mesh): acc += u (x) w
// only -- Kf is applied once per block when the item is finished (fix_general), which halves the FP64 work.
template <bool CR>
__device__ __forceinline__ double add_general(double (&acc)[9], double rho, const double (&u)[3], const double (&w)[3]) {
    if constexpr (CR) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) acc[3 * r + s] = fma(u[r], w[s], acc[3 * r + s]);
        return 0.0;
    }
    const double ru[3] = {rho * u[0], rho * u[1], rho * u[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) acc[3 * r + s] = fma(ru[r], w[s], fma(w[r], u[s], acc[3 * r + s]));
    return fma(u[0], w[0], fma(u[1], w[1], u[2] * w[2]));
}

// acc (xx, xy, xz, yy, yz, zz) += (rho + 1) u u^T; returns |u|^2   (CR: acc += u u^T, see fix_sym)
template <bool CR>
__device__ __forceinline__ double add_sym(double (&acc)[6], double rho, const double (&u)[3]) {
    const double r1 = CR ? 1.0 : rho + 1.0;
    const double a = r1 * u[0], b = r1 * u[1], c = r1 * u[2];
    acc[0] = fma(a, u[0], acc[0]);
    acc[1] = fma(a, u[1], acc[1]);
    acc[2] = fma(a, u[2], acc[2]);
    acc[3] = fma(b, u[1], acc[3]);
    acc[4] = fma(b, u[2], acc[4]);
    acc[5] = fma(c, u[2], acc[5]);
    return fma(u[0], u[0], fma(u[1], u[1], u[2] * u[2]));
}

// CR: turn the plain sums T = sum u (x) w into (Kf(T) without the trace part, tr T)
template <bool CR>
__device__ __forceinline__ void fix_general(double (&p)[9], double &tr, double rho) {
    if constexpr (CR) {
        tr = p[0] + p[4] + p[8];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            p[4 * r] *= rho + 1.0;
#pragma unroll
            for (int s = r + 1; s < 3; ++s) {
                const double a = p[3 * r + s], b = p[3 * s + r];
                p[3 * r + s] = fma(rho, a, b);
                p[3 * s + r] = fma(rho, b, a);
            }
        }
    }
}

template <bool CR>
__device__ __forceinline__ void fix_sym(double (&q)[6], double &tr, double rho) {
    if constexpr (CR) {
        tr = q[0] + q[3] + q[5];
#pragma unroll
        for (int i = 0; i < 6; ++i) q[i] *= rho + 1.0;
    }
}

__device__ __forceinline__ void store_block(double *tile, unsigned where, unsigned pitch, const double (&k)[9]) {
    if (where == 0xFFFFu) return;
    double *p = tile + where;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) p[r * pitch + s] = k[3 * r + s];
}

__device__ __forceinline__ void sym_to_full(const double (&a)[6], double c, double tr, double (&k)[9]) {
    k[0] = c * (a[0] + tr);
    k[1] = k[3] = c * a[1];
    k[2] = k[6] = c * a[2];
    k[4] = c * (a[3] + tr);
    k[5] = k[7] = c * a[4];
    k[8] = c * (a[5] + tr);
}

// Where a finished item goes: the chunk's tile (once the tile has been handed over to this chunk) at the item's offsets.
struct TileCtx {
    double *tile;
    volatile int *free_at;
    int need;
    unsigned ihx, ihy;  // packed 16-bit tile offsets of the three blocks and the row pitch (decoded when storing)
    __device__ __forceinline__ void wait() const {
        if ((threadIdx.x & 31) == 0) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(const_cast<int *>(free_at));
            for (;;) {  // acquire load instead of a plain load + MEMBAR
                int v;
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
                if (v >= need) break;
                __nanosleep(32);
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ void put(int k, const double (&v)[9]) const {
        store_block(tile, k == 0 ? (ihx & 0xFFFFu) : (k == 1 ? (ihx >> 16) : (ihy & 0xFFFFu)), ihy >> 16, v);
    }
};

// ---- per item type: which vertices' records an entry needs, how a cell adds to the sums, how the sums become blocks
template <int NG>
struct Rec {
    G4 g[NG];
    bool valid;
};

template <int NG>
__device__ __forceinline__ Rec<NG> load_rec(uint32_t code, const double *__restrict__ rec) {
    Rec<NG> r;
    r.valid = code != 0xFFFFFFFFu;
#pragma unroll
    for (int m = 0; m < NG; ++m) r.g[m] = G4{0.0, 0.0, 0.0, 0.0};
    if (r.valid) {
        const double *rc = rec + (size_t)(code >> 6) * 16;
        r.g[0] = ldg256(rc + 4 * ((code >> 4) & 3u));
        if constexpr (NG > 1) r.g[1] = ldg256(rc + 4 * ((code >> 2) & 3u));
        if constexpr (NG > 2) r.g[2] = ldg256(rc + 4 * (code & 3u));
    }
    return r;
}

constexpr double C10 = 0.1, C30 = 1.0 / 30.0, C15 = 2.0 / 15.0;

template <int TYPE, bool CR>
struct Item;

template <bool CR>
struct Item<0, CR> {
    double rho = 0.0;  // A: (v, v) = 0.1 Kf(g_i g_i^T)
    static constexpr int NG = 1;
    double q[6] = {0, 0, 0, 0, 0, 0}, tr = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z};
        tr += add_sym<CR>(q, r.g[0].rho, u);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double o[9];
        fix_sym<CR>(q, tr, rho);
        sym_to_full(q, C10, tr, o);
        ctx.put(0, o);
    }
};

template <bool CR>
struct Item<1, CR> {
    double rho = 0.0;  // B: (v, v') = -1/30 P, (v, E) = 0.1 P - Q / 30 with P = Kf(g_i (x) g_k), Q = Kf(g_i g_i^T)
    static constexpr int NG = 2;
    double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, q[6] = {0, 0, 0, 0, 0, 0}, trp = 0.0, trq = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z}, w[3] = {r.g[1].x, r.g[1].y, r.g[1].z};
        trp += add_general<CR>(p, r.g[0].rho, u, w);
        trq += add_sym<CR>(q, r.g[0].rho, u);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double o[9];
        fix_general<CR>(p, trp, rho);
        fix_sym<CR>(q, trq, rho);
#pragma unroll
        for (int i = 0; i < 9; i += 4) p[i] += trp;
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = -C30 * p[i];
        ctx.put(0, o);
        sym_to_full(q, -C30, trq, o);
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = fma(C10, p[i], o[i]);
        ctx.put(1, o);
    }
};

template <bool CR>
struct Item<2, CR> {
    double rho = 0.0;  // C: (v, e_xy) = Kf(g_i (x) (g_i + g_z)) / 30, since g_x + g_y = -(g_i + g_z)
    static constexpr int NG = 2;
    double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, trp = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z};
        const double w[3] = {u[0] + r.g[1].x, u[1] + r.g[1].y, u[2] + r.g[1].z};
        trp += add_general<CR>(p, r.g[0].rho, u, w);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double o[9];
        fix_general<CR>(p, trp, rho);
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = C30 * (p[i] + ((i % 4 == 0) ? trp : 0.0));
        ctx.put(0, o);
    }
};

template <bool CR>
struct Item<3, CR> {
    double rho = 0.0;  // D: (E, v_i) = 0.1 P^T - Qi / 30, (E, v_j) = 0.1 P - Qj / 30, (E, E) = 2/15 (2 Qi + 2 Qj + P + P^T)
    static constexpr int NG = 2;
    double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, qi[6] = {0, 0, 0, 0, 0, 0}, qj[6] = {0, 0, 0, 0, 0, 0};
    double trp = 0.0, tri = 0.0, trj = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z}, w[3] = {r.g[1].x, r.g[1].y, r.g[1].z};
        trp += add_general<CR>(p, r.g[0].rho, u, w);
        tri += add_sym<CR>(qi, r.g[0].rho, u);
        trj += add_sym<CR>(qj, r.g[0].rho, w);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double fi[9], fj[9], o[9];
        fix_general<CR>(p, trp, rho);
        fix_sym<CR>(qi, tri, rho);
        fix_sym<CR>(qj, trj, rho);
        sym_to_full(qi, 1.0, tri, fi);
        sym_to_full(qj, 1.0, trj, fj);
#pragma unroll
        for (int i = 0; i < 9; i += 4) p[i] += trp;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) o[3 * r + s] = C10 * p[3 * s + r] - C30 * fi[3 * r + s];
        ctx.put(0, o);
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = C10 * p[i] - C30 * fj[i];
        ctx.put(1, o);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
                o[3 * r + s] = C15 * (2.0 * (fi[3 * r + s] + fj[3 * r + s]) + p[3 * r + s] + p[3 * s + r]);
        ctx.put(2, o);
    }
};

template <bool CR>
struct Item<4, CR> {
    double rho = 0.0;  // E: (E, v_x), (E, e_ix), (E, e_jx) with s = g_i + g_j
    static constexpr int NG = 3;
    double k0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, k1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, k2[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double rho = r.g[0].rho;
        const double a[3] = {r.g[0].x, r.g[0].y, r.g[0].z}, b[3] = {r.g[1].x, r.g[1].y, r.g[1].z};
        const double x[3] = {r.g[2].x, r.g[2].y, r.g[2].z};
        const double s[3] = {a[0] + b[0], a[1] + b[1], a[2] + b[2]};
        const double ax[3] = {a[0] + x[0], a[1] + x[1], a[2] + x[2]};
        const double bx[3] = {b[0] + x[0], b[1] + x[1], b[2] + x[2]};
        t0 += add_general<CR>(k0, rho, s, x);
        t1 += add_general<CR>(k1, rho, s, ax);
        t1 += add_general<CR>(k1, rho, b, x);
        t2 += add_general<CR>(k2, rho, s, bx);
        t2 += add_general<CR>(k2, rho, a, x);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        fix_general<CR>(k0, t0, rho);
        fix_general<CR>(k1, t1, rho);
        fix_general<CR>(k2, t2, rho);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const bool d = (i % 4 == 0);
            k0[i] = -C30 * (k0[i] + (d ? t0 : 0.0));
            k1[i] = C15 * (k1[i] + (d ? t1 : 0.0));
            k2[i] = C15 * (k2[i] + (d ? t2 : 0.0));
        }
        ctx.put(0, k0);
        ctx.put(1, k1);
        ctx.put(2, k2);
    }
};

template <bool CR>
struct Item<5, CR> {
    double rho = 0.0;  // F: (E, E_opp) = -2/15 Kf(s s^T), since g_x + g_y = -(g_i + g_j)
    static constexpr int NG = 2;
    double q[6] = {0, 0, 0, 0, 0, 0}, tr = 0.0;
    __device__ __forceinline__ void add(const Rec<NG> &r) {
        rho = r.g[0].rho;
        const double s[3] = {r.g[0].x + r.g[1].x, r.g[0].y + r.g[1].y, r.g[0].z + r.g[1].z};
        tr += add_sym<CR>(q, r.g[0].rho, s);
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double o[9];
        fix_sym<CR>(q, tr, rho);
        sym_to_full(q, -C15, tr, o);
        ctx.put(0, o);
    }
};

// One task = up to 32 items of one type; lane = item.  e points at the lane's first ELL entry (stride 32); c0, c1 are
// the lane's first two entries (fetched while the previous task was computed).  Items that read at most two records per
// cell keep two cells in flight; type E (three records, 27 sums) walks one cell at a time to stay inside the register
// budget.  pf0, pf1: first entries of the NEXT task -- their cells' records (one 128-byte line each) are pulled into L1
// once this task's sums are complete.
template <int TYPE, bool CR, int UNROLL_NG>
__device__ __forceinline__ void run_task(const uint32_t *__restrict__ e, int kmax, uint32_t c0, uint32_t c1,
                                         const double *__restrict__ rec, uint32_t pf0, uint32_t pf1,
                                         const TileCtx &ctx) {
    constexpr int NG = Item<TYPE, CR>::NG;
    // cells in flight per iteration: long items with few sums walk many cells at once (their task would otherwise
    // hold its chunk's tile for tens of dependent load latencies); type E stays at one (register budget)
    constexpr int U = (TYPE == 0) ? NFA_P2_UA : ((TYPE == 1) ? NFA_P2_UB : ((NG <= UNROLL_NG) ? 2 : 1));
#if NFA_P2_E2
    if constexpr (TYPE == 4) {
        // a face has at most two cells: both cells' records are requested at once (one memory round trip instead of two)
        // and the three blocks are formed and stored one after the other, so that no 27 sums are live at any time
        const Rec<3> r0 = load_rec<3>(c0, rec), r1 = load_rec<3>(c1, rec);  // an absent cell reads as zeros
        if (pf0 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf0 >> 6) * 16));
        if (pf1 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf1 >> 6) * 16));
        const double rho = r0.g[0].rho;
        const double a0[3] = {r0.g[0].x, r0.g[0].y, r0.g[0].z}, b0[3] = {r0.g[1].x, r0.g[1].y, r0.g[1].z};
        const double x0[3] = {r0.g[2].x, r0.g[2].y, r0.g[2].z};
        const double a1[3] = {r1.g[0].x, r1.g[0].y, r1.g[0].z}, b1[3] = {r1.g[1].x, r1.g[1].y, r1.g[1].z};
        const double x1[3] = {r1.g[2].x, r1.g[2].y, r1.g[2].z};
        const double s0[3] = {a0[0] + b0[0], a0[1] + b0[1], a0[2] + b0[2]};
        const double s1[3] = {a1[0] + b1[0], a1[1] + b1[1], a1[2] + b1[2]};
        ctx.wait();
        {
            double k[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0.0;
            t += add_general<CR>(k, r0.g[0].rho, s0, x0);
            t += add_general<CR>(k, r1.g[0].rho, s1, x1);
            fix_general<CR>(k, t, rho);
#pragma unroll
            for (int i = 0; i < 9; ++i) k[i] = -C30 * (k[i] + ((i % 4 == 0) ? t : 0.0));
            ctx.put(0, k);
        }
        {
            const double ax0[3] = {a0[0] + x0[0], a0[1] + x0[1], a0[2] + x0[2]};
            const double ax1[3] = {a1[0] + x1[0], a1[1] + x1[1], a1[2] + x1[2]};
            double k[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0.0;
            t += add_general<CR>(k, r0.g[0].rho, s0, ax0);
            t += add_general<CR>(k, r0.g[0].rho, b0, x0);
            t += add_general<CR>(k, r1.g[0].rho, s1, ax1);
            t += add_general<CR>(k, r1.g[0].rho, b1, x1);
            fix_general<CR>(k, t, rho);
#pragma unroll
            for (int i = 0; i < 9; ++i) k[i] = C15 * (k[i] + ((i % 4 == 0) ? t : 0.0));
            ctx.put(1, k);
        }
        {
            const double bx0[3] = {b0[0] + x0[0], b0[1] + x0[1], b0[2] + x0[2]};
            const double bx1[3] = {b1[0] + x1[0], b1[1] + x1[1], b1[2] + x1[2]};
            double k[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0.0;
            t += add_general<CR>(k, r0.g[0].rho, s0, bx0);
            t += add_general<CR>(k, r0.g[0].rho, a0, x0);
            t += add_general<CR>(k, r1.g[0].rho, s1, bx1);
            t += add_general<CR>(k, r1.g[0].rho, a1, x1);
            fix_general<CR>(k, t, rho);
#pragma unroll
            for (int i = 0; i < 9; ++i) k[i] = C15 * (k[i] + ((i % 4 == 0) ? t : 0.0));
            ctx.put(2, k);
        }
        return;
    }
    if constexpr (TYPE == 2) {  // vertex row, opposite edge of a face: at most two cells, one block
        const Rec<2> r0 = load_rec<2>(c0, rec), r1 = load_rec<2>(c1, rec);
        if (pf0 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf0 >> 6) * 16));
        if (pf1 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf1 >> 6) * 16));
        const double u0[3] = {r0.g[0].x, r0.g[0].y, r0.g[0].z}, u1[3] = {r1.g[0].x, r1.g[0].y, r1.g[0].z};
        const double w0[3] = {u0[0] + r0.g[1].x, u0[1] + r0.g[1].y, u0[2] + r0.g[1].z};
        const double w1[3] = {u1[0] + r1.g[1].x, u1[1] + r1.g[1].y, u1[2] + r1.g[1].z};
        double k[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, t = 0.0;
        t += add_general<CR>(k, r0.g[0].rho, u0, w0);
        t += add_general<CR>(k, r1.g[0].rho, u1, w1);
        fix_general<CR>(k, t, r0.g[0].rho);
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = C30 * (k[i] + ((i % 4 == 0) ? t : 0.0));
        ctx.wait();
        ctx.put(0, k);
        return;
    }
#endif
    Item<TYPE, CR> item;
    if constexpr (U > 2) {
        uint32_t c[U];
        c[0] = c0;
        c[1] = c1;
#pragma unroll
        for (int u = 2; u < U; ++u) c[u] = (u < kmax) ? ld_list(e + 32 * u) : 0xFFFFFFFFu;
#pragma unroll 1
        for (int k = 0; k < kmax; k += U) {
            Rec<NG> r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) r[u] = load_rec<NG>(c[u], rec);
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = (k + U + u < kmax) ? ld_list(e + 32 * (k + U + u)) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r[u].valid) item.add(r[u]);
        }
    } else if constexpr (U == 2) {
#pragma unroll 1
        for (int k = 0; k < kmax; k += 2) {
            const Rec<NG> r0 = load_rec<NG>(c0, rec), r1 = load_rec<NG>(c1, rec);
            c0 = (k + 2 < kmax) ? ld_list(e + 32 * (k + 2)) : 0xFFFFFFFFu;
            c1 = (k + 3 < kmax) ? ld_list(e + 32 * (k + 3)) : 0xFFFFFFFFu;
            if (r0.valid) item.add(r0);
            if (r1.valid) item.add(r1);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < kmax; ++k) {
            const Rec<NG> r0 = load_rec<NG>(c0, rec);
            c0 = c1;
            c1 = (k + 2 < kmax) ? ld_list(e + 32 * (k + 2)) : 0xFFFFFFFFu;
            if (r0.valid) item.add(r0);
        }
    }
    if (pf0 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf0 >> 6) * 16));
    if (pf1 != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)(pf1 >> 6) * 16));
    ctx.wait();
    item.finish(ctx);
}

// Warps of a CTA pull tasks from the CTA's contiguous chunk range through one shared counter; chunk c lives in tile
// (c - c_begin) % NBUF.  The warp that finishes the last task of a chunk sends the tile to HBM with one bulk copy,
// waits until the copy engine has read it and hands the tile to chunk c + NBUF.  No CTA-wide barrier in the loop.
template <bool CR, int THREADS, int NB>
__global__ void __launch_bounds__(THREADS, 1)
    p2tet_elasticity_kernel(int nbuf_rt, int n_chunks, const int4 *__restrict__ chunk_hdr, const int4 *__restrict__ task_hdr,
                            const uint2 *__restrict__ item_hdr, const uint32_t *__restrict__ ell,
                            const double *__restrict__ rec, int tile_stride, const int32_t *__restrict__ cells_ptr,
                            const int32_t *__restrict__ cells, double *__restrict__ vals) {
    // NB > 0: the number of tiles is a compile-time power of two, so that the tile of a chunk ((c - c_begin) % NBUF) and its
    // round ((c - c_begin) / NBUF) cost a mask and a shift instead of two integer divisions per task (the task loop's scalar
    // bookkeeping is 40 % of the instructions the kernel executes)
    const int NBUF = NB > 0 ? NB : nbuf_rt;
    extern __shared__ __align__(128) double smem[];  // NBUF tiles of tile_stride doubles, then the control words
    // ctrl: [0] next task, [1] c_begin, [2] c_end, [3] task_end, [4 .. 4+NBUF) tasks left per tile, then free_at per tile.
    // Everything the tail of an iteration needs is re-read from here / from the (L1-resident) headers instead of being
    // kept in registers across the accumulation loops, which need the whole register file.
    volatile int *ctrl = reinterpret_cast<volatile int *>(smem + (size_t)NBUF * tile_stride);
    if (threadIdx.x == 0) {
        const int c_begin = (int)((int64_t)n_chunks * blockIdx.x / gridDim.x);
        const int c_end = (int)((int64_t)n_chunks * (blockIdx.x + 1) / gridDim.x);
        ctrl[1] = c_begin;
        ctrl[2] = c_end;
        ctrl[0] = ctrl[3] = 0;
        if (c_begin < c_end) {
            ctrl[0] = __ldg(chunk_hdr + c_begin).z;
            ctrl[3] = __ldg(chunk_hdr + c_end - 1).z + __ldg(chunk_hdr + c_end - 1).w;
        }
        for (int b = 0; b < NBUF; ++b) {
            ctrl[4 + b] = (c_begin + b < c_end) ? __ldg(chunk_hdr + c_begin + b).w : 0;
            ctrl[4 + NBUF + b] = 0;
        }
    }
    __syncthreads();
    // software pipeline over tasks, per warp: while task i is computed, the header of task i+2 and the first two ELL
    // entries of task i+1 are in flight, and the records task i+1 reads first are prefetched into L1
    const int lane = threadIdx.x & 31;
    const int task_end = ctrl[3], c_begin_r = ctrl[1];
    const int4 none = make_int4(0, 0, 0, -1);
    const uint32_t ctrl_s = (uint32_t)__cvta_generic_to_shared(const_cast<int *>(ctrl));
    auto grab = [&]() {
        int t = 0;
        if (lane == 0)  // plain shared-memory atomic (the compiler wraps atomicAdd in its warp-aggregation sequence)
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(ctrl_s) : "memory");
        return __shfl_sync(0xffffffffu, t, 0);
    };
    int4 th = none, th1 = none;
#if NFA_P2_DEEP
    int4 th2 = none;
#endif
    {
        const int t0 = grab();
        if (t0 < task_end) th = ld_list(task_hdr + t0);
        const int t1 = grab();
        if (t1 < task_end) th1 = ld_list(task_hdr + t1);
#if NFA_P2_DEEP
        const int t2 = grab();
        if (t2 < task_end) th2 = ld_list(task_hdr + t2);
#endif
    }
#if NFA_P2_GRAB2
    int t_pending = 0;  // lane 0: the ticket taken for the NEXT iteration
    if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t_pending) : "r"(ctrl_s) : "memory");
    auto grab_ahead = [&]() {
        const int t = __shfl_sync(0xffffffffu, t_pending, 0);
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t_pending) : "r"(ctrl_s) : "memory");
        return t;
    };
#else
    auto grab_ahead = grab;
#endif
    uint32_t c0 = 0xFFFFFFFFu, c1 = 0xFFFFFFFFu;
    if (th.w != -1) {
        const uint32_t *e = ell + (uint32_t)th.y + lane;
        c0 = ld_list(e);
        if (((th.x >> 8) & 0xFF) > 1) c1 = ld_list(e + 32);
    }
    while (th.w != -1) {
#if NFA_P2_DEEP
        int4 th3 = none;
        {
            const int t3 = grab_ahead();
            if (t3 < task_end) th3 = ld_list(task_hdr + t3);
        }
        if (th2.w != -1) {  // task i+2: its ELL lines (a lane per line) and item headers go to L2
            const int kmax2 = (th2.x >> 8) & 0xFF;
            for (int k = lane; k < kmax2; k += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(ell + (uint32_t)th2.y + 32 * k));
            if ((lane & 15) == 0 && lane < (th2.w & 63)) asm volatile("prefetch.global.L2 [%0];" ::"l"(item_hdr + th2.z + lane));
        }
#else
        int4 th2 = none;
        {
            const int t2 = grab_ahead();
            if (t2 < task_end) th2 = ld_list(task_hdr + t2);
        }
#endif
        uint32_t n0 = 0xFFFFFFFFu, n1 = 0xFFFFFFFFu;  // first entries of task i+1
        if (th1.w != -1) {
            const uint32_t *e1 = ell + (uint32_t)th1.y + lane;
            n0 = ld_list(e1);
            const int kmax1 = (th1.x >> 8) & 0xFF;
            if (kmax1 > 1) n1 = ld_list(e1 + 32);
#if NFA_P2_PF_ELL && !NFA_P2_DEEP
            // the rest of task i+1's ELL lines (one 128-byte line per entry index): a lane prefetches line `lane`, so that
            // the later rounds of a long item find their entries in cache instead of waiting for a DRAM miss each
            for (int k = 2 + lane; k < kmax1; k += 32) {
                const uint32_t *pl = ell + (uint32_t)th1.y + 32 * k;
#if NFA_P2_PF_ELL == 2
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pl));
#else
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
#endif
            }
#endif
#if NFA_P2_PF_IH
            // task i+1's item headers (32 x 8 bytes = two lines): lanes 0 and 16 pull them into L1
            if ((lane & 15) == 0 && lane < (th1.w & 63)) asm volatile("prefetch.global.L1 [%0];" ::"l"(item_hdr + th1.z + lane));
#endif
        }
        {
            const int type = th.x & 0xFF, kmax = (th.x >> 8) & 0xFF, n_items = th.w & 63, c = (int)((unsigned)th.w >> 6);
            const uint32_t *e = ell + (uint32_t)th.y + lane;
            uint2 ih = make_uint2(0xFFFFFFFFu, 0xFFFFu);
            if (lane < n_items) ih = ld_list(item_hdr + th.z + lane);
            const int it = c - c_begin_r, b = it % NBUF;
            // the tile sits on the same 16-byte phase as its image in vals (parity of the chunk's first block)
            double *tile = smem + (size_t)b * tile_stride + ((th.x >> 16) & 1);
            const TileCtx ctx = {tile, ctrl + 4 + NBUF + b, it / NBUF, ih.x, ih.y};
            switch (type) {
                case 0: run_task<0, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
                case 1: run_task<1, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
                case 2: run_task<2, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
                case 3: run_task<3, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
                case 4: run_task<4, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
                default: run_task<5, CR, (THREADS <= 384 ? 3 : P2_UNROLL_NG)>(e, kmax, c0, c1, rec, n0, n1, ctx); break;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            const int c = (int)((unsigned)th.w >> 6);
            const int it = c - c_begin_r, b = it % NBUF;
            // release of this task's tile stores / acquire of the other tasks' in one acq_rel atomic (no separate MEMBAR)
            int left;
#if NFA_P2_RELAXED
            // the task's tile stores and this decrement pass through the SM's shared-memory pipe in program order; the warp
            // that sees the count reach zero fences before it hands the tile to the copy engine
            asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], -1;" : "=r"(left) : "r"(ctrl_s + 4u * (4u + (unsigned)b)) : "memory");
#else
            asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], -1;" : "=r"(left) : "r"(ctrl_s + 4u * (4u + (unsigned)b)) : "memory");
#endif
            if (left == 1) {  // the chunk's last task: the tile is complete
#if NFA_P2_RELAXED
                asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int4 ch = __ldg(chunk_hdr + c);
                const int b0 = ch.x, n_vals = 9 * ch.y, head = b0 & 1;
                const int n_bulk = (n_vals - head) & ~1;
                const double *tile = smem + (size_t)b * tile_stride + head;
                double *dst = vals + (int64_t)9 * b0;
                if (n_bulk > 0) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + head);
#if (NFA_P2_L2HINT & 2)
                    const uint64_t pol = l2_policy_evict_first();  // the values are not read again by this kernel
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + head),
                                 "r"(src), "r"(n_bulk * 8), "l"(pol)
                                 : "memory");
#else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head),
                                 "r"(src), "r"(n_bulk * 8)
                                 : "memory");
#endif
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (head) dst[0] = tile[0];
                if (head + n_bulk < n_vals) dst[n_vals - 1] = tile[n_vals - 1];
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                ctrl[4 + b] = (c + NBUF < ctrl[2]) ? __ldg(chunk_hdr + c + NBUF).w : 0;
                __threadfence_block();
                ctrl[4 + NBUF + b] = it / NBUF + 1;
            }
        }
        th = th1;
        th1 = th2;
#if NFA_P2_DEEP
        th2 = th3;
#endif
        c0 = n0;
        c1 = n1;
    }
    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// =====================================================================================================================
// Staged variant: no global load on the compute path.  Everything a chunk reads -- its ELL entries, item headers, task
// headers and the 128-byte geometry records of the cells it touches (entries name a cell by its slot in the chunk's
// cell list) -- is copied into the chunk's shared-memory buffer by bulk asynchronous copies (cp.async.bulk global ->
// shared, completion counted on the buffer's mbarrier) issued when the buffer is handed over, i.e. NBUF chunks ahead of
// use.  The warp that retires a chunk sends its tile to HBM, waits until the copy engine has read it and refills the
// buffer for chunk c + NBUF; the mbarrier phase a task waits for therefore means "tile free and inputs present".
struct StageLayout {
    int rec_off, ell_off, item_off, task_off, buf_bytes;  // byte offsets inside a buffer (the tile sits at offset 0)
};

__device__ __forceinline__ G4 lds_g4(const double *p) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    return G4{a.x, a.y, b.x, b.y};
}

template <int NG>
__device__ __forceinline__ Rec<NG> load_rec_s(uint32_t code, const double *rec_s) {
    Rec<NG> r;
    r.valid = code != 0xFFFFFFFFu;
#pragma unroll
    for (int m = 0; m < NG; ++m) r.g[m] = G4{0.0, 0.0, 0.0, 0.0};
    if (r.valid) {
        const double *rc = rec_s + (size_t)(code >> 6) * 16;
        r.g[0] = lds_g4(rc + 4 * ((code >> 4) & 3u));
        if constexpr (NG > 1) r.g[1] = lds_g4(rc + 4 * ((code >> 2) & 3u));
        if constexpr (NG > 2) r.g[2] = lds_g4(rc + 4 * (code & 3u));
    }
    return r;
}

#ifndef NFA_P2S_UA
#define NFA_P2S_UA 4
#endif
#ifndef NFA_P2S_U
#define NFA_P2S_U 2
#endif

template <int TYPE, bool CR>
__device__ __forceinline__ void run_task_s(const uint32_t *e, int kmax, const double *rec_s, const TileCtx &ctx) {
    constexpr int NG = Item<TYPE, CR>::NG;
    constexpr int U = (TYPE == 0) ? NFA_P2S_UA : ((TYPE == 4) ? 1 : NFA_P2S_U);
    Item<TYPE, CR> item;
#pragma unroll 1
    for (int k = 0; k < kmax; k += U) {
        uint32_t c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (k + u < kmax) ? e[32 * (k + u)] : 0xFFFFFFFFu;
        Rec<NG> r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = load_rec_s<NG>(c[u], rec_s);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r[u].valid) item.add(r[u]);
    }
    item.finish(ctx);
}

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <bool CR, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    p2tet_staged_kernel(int NBUF, int n_chunks, const int4 *__restrict__ chunk_hdr, const int4 *__restrict__ chunk_ext,
                        const int4 *__restrict__ task_hdr_rel, const uint2 *__restrict__ item_hdr,
                        const uint32_t *__restrict__ ell_local, const int32_t *__restrict__ cells_ptr,
                        const int32_t *__restrict__ cells, const double *__restrict__ rec, StageLayout L,
                        int rec_tma, double *__restrict__ vals) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ctrl_base = smem_raw + (size_t)NBUF * L.buf_bytes;
    // ctrl: [0] next task, [1] c_begin, [2] c_end, [3] task_end, [4 .. 4+NBUF) tasks left in the buffer's chunk,
    // [12 .. 12+NBUF) generation of the buffer = (chunk offset / NBUF) of the chunk it was last handed to.  A task first
    // waits for its generation (monotonic, so a warp that ran 2 NBUF chunks ahead cannot mistake an older phase of the
    // mbarrier for its own: the parity alone repeats every two phases), then for the mbarrier phase of that fill.
    volatile int *ctrl = reinterpret_cast<volatile int *>(ctrl_base);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(ctrl_base + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const int c_begin = (int)((int64_t)n_chunks * blockIdx.x / gridDim.x);
        const int c_end = (int)((int64_t)n_chunks * (blockIdx.x + 1) / gridDim.x);
        ctrl[1] = c_begin;
        ctrl[2] = c_end;
        ctrl[0] = ctrl[3] = 0;
        if (c_begin < c_end) {
            ctrl[0] = __ldg(chunk_hdr + c_begin).z;
            ctrl[3] = __ldg(chunk_hdr + c_end - 1).z + __ldg(chunk_hdr + c_end - 1).w;
        }
        for (int b = 0; b < NBUF; ++b) {
            ctrl[4 + b] = 0;
            ctrl[12 + b] = -1;
            // arrivals per phase: lane 0's arrive.expect_tx + one deferred cp.async arrival per lane of the filling warp
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 33;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar + b)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int c_begin = ctrl[1], c_end = ctrl[2], task_end = ctrl[3];

    // hands buffer b to chunk cn: the whole warp issues the copies (one 128-byte record per lane and trip)
    auto fill = [&](int b, int cn) {
        const int4 ch = __ldg(chunk_hdr + cn), ex = __ldg(chunk_ext + cn);
        const int cp0 = __ldg(cells_ptr + cn), n_cells = __ldg(cells_ptr + cn + 1) - cp0;
        unsigned char *buf = smem_raw + (size_t)b * L.buf_bytes;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar + b);
        if (lane == 0) {
            ctrl[4 + b] = ch.w;
            ctrl[12 + b] = (cn - c_begin) / NBUF;
            __threadfence_block();
            const uint32_t bytes = (rec_tma ? (uint32_t)n_cells * 128u : 0u) + (uint32_t)ex.y * 4u + (uint32_t)ex.w * 8u +
                                   (uint32_t)ch.w * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            bulk_g2s(buf + L.ell_off, ell_local + ex.x, (uint32_t)ex.y * 4u, bar);
            bulk_g2s(buf + L.item_off, item_hdr + ex.z, (uint32_t)ex.w * 8u, bar);
            bulk_g2s(buf + L.task_off, task_hdr_rel + ch.z, (uint32_t)ch.w * 16u, bar);
        }
        // records.  The cell list is read with coalesced loads (128 cells per trip) before any copy is issued, so the
        // copies do not wait on one dependent load each.  rec_tma: one bulk copy per 128-byte record line; else 16-byte
        // asynchronous copies (LDGSTS, L2 only), eight per line, a warp trip moves four lines.  Completion is counted on
        // the same mbarrier (bulk copies: bytes; LDGSTS: one deferred arrival per lane).
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(buf + L.rec_off);
        for (int m0 = 0; m0 < n_cells; m0 += 128) {
            int cid[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int q = m0 + 32 * m + lane;
                cid[m] = q < n_cells ? __ldg(cells + cp0 + q) : 0;
            }
            if (rec_tma) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int q = m0 + 32 * m + lane;
                    if (q < n_cells) bulk_g2s(buf + L.rec_off + (size_t)q * 128, rec + (size_t)cid[m] * 16, 128u, bar);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int q = m0 + 32 * m + 4 * jj + (lane >> 3);
                        const int c = __shfl_sync(0xffffffffu, cid[m], 4 * jj + (lane >> 3));
                        if (q < n_cells)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 128u * (uint32_t)q + 16u * (lane & 7)),
                                         "l"(rec + (size_t)c * 16 + 2 * (lane & 7))
                                         : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        // the lists the NEXT fill of this buffer starts with (chunk cn + NBUF; a CTA's chunks are contiguous in every
        // array, so its cell list starts about NBUF lists further on) are pulled into L2 now
        if (cn + NBUF < c_end) {
            const void *pf = lane == 0   ? (const void *)(chunk_hdr + cn + NBUF)
                             : lane == 1 ? (const void *)(chunk_ext + cn + NBUF)
                             : lane == 2 ? (const void *)(cells_ptr + cn + NBUF)
                                         : (const void *)(cells + cp0 + (size_t)(NBUF - 1) * n_cells + 32 * (lane - 3));
            if (lane < 3 + 2 * NBUF) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        }
    };
    if (warp < NBUF && c_begin + warp < c_end) fill(warp, c_begin + warp);

    int it = 0;  // the chunk (offset from c_begin) this warp's current task belongs to
    int4 chh = make_int4(0, 0, 0, 0);
    if (c_begin < c_end) chh = __ldg(chunk_hdr + c_begin);
    int t = 0;
    if (lane == 0) t = atomicAdd(const_cast<int *>(ctrl), 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    while (t < task_end) {
        while (t >= chh.z + chh.w) {
            ++it;
            chh = __ldg(chunk_hdr + c_begin + it);
        }
        int tn = 0;  // next task: the round trip of the shared-memory atomic hides behind this task
        if (lane == 0) tn = atomicAdd(const_cast<int *>(ctrl), 1);
        const int b = it % NBUF;
        {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar + b), parity = (uint32_t)(it / NBUF) & 1u;
            while (ctrl[12 + b] < it / NBUF) __nanosleep(64);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar), "r"(parity)
                    : "memory");
            }
        }
        unsigned char *buf = smem_raw + (size_t)b * L.buf_bytes;
        {
            const int4 th = reinterpret_cast<const int4 *>(buf + L.task_off)[t - chh.z];
            const int type = th.x & 0xFF, kmax = (th.x >> 8) & 0xFF, n_items = th.w & 63;
            const uint32_t *e = reinterpret_cast<const uint32_t *>(buf + L.ell_off) + th.y + lane;
            uint2 ih = make_uint2(0xFFFFFFFFu, 0xFFFFu);
            if (lane < n_items) ih = reinterpret_cast<const uint2 *>(buf + L.item_off)[th.z + lane];
            // the tile sits on the same 16-byte phase as its image in vals (parity of the chunk's first block)
            double *tile = reinterpret_cast<double *>(buf) + ((th.x >> 16) & 1);
            const double *rec_s = reinterpret_cast<const double *>(buf + L.rec_off);
            const TileCtx ctx = {tile, nullptr, 0, ih.x, ih.y};
            switch (type) {
                case 0: run_task_s<0, CR>(e, kmax, rec_s, ctx); break;
                case 1: run_task_s<1, CR>(e, kmax, rec_s, ctx); break;
                case 2: run_task_s<2, CR>(e, kmax, rec_s, ctx); break;
                case 3: run_task_s<3, CR>(e, kmax, rec_s, ctx); break;
                case 4: run_task_s<4, CR>(e, kmax, rec_s, ctx); break;
                default: run_task_s<5, CR>(e, kmax, rec_s, ctx); break;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = atomicSub(const_cast<int *>(ctrl) + 4 + b, 1) == 1;  // the chunk's last task: the tile is complete
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (lane == 0) {
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int b0 = chh.x, n_vals = 9 * chh.y, head = b0 & 1;
                const int n_bulk = (n_vals - head) & ~1;
                const double *tile = reinterpret_cast<const double *>(buf) + head;
                double *dst = vals + (int64_t)9 * b0;
                if (n_bulk > 0) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + head);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head),
                                 "r"(src), "r"(n_bulk * 8)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (head) dst[0] = tile[0];
                if (head + n_bulk < n_vals) dst[n_vals - 1] = tile[n_vals - 1];
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            const int cn = c_begin + it + NBUF;
            if (cn < c_end) fill(b, cn);
        }
        t = __shfl_sync(0xffffffffu, tn, 0);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace

extern "C" int nfa_bplan_block_pattern(const nfa_bplan *bp, int32_t *bptr, int32_t *bcols, nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_block_pattern: plan is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (bptr) NFA_CUDA(cudaMemcpyAsync(bptr, bp->bptr.p, 4 * (bp->n_nodes + 1), cudaMemcpyDeviceToDevice, s));
    if (bcols) NFA_CUDA(cudaMemcpyAsync(bcols, bp->bcols.p, 4 * bp->n_bnz, cudaMemcpyDeviceToDevice, s));
    return 0;
}

extern "C" int64_t nfa_bplan_n_blocks(const nfa_bplan *bp) { return bp ? bp->n_bnz : -1; }

extern "C" int nfa_bplan_set_p2_items(nfa_bplan *bp, int64_t n_chunks, const int32_t *chunk_hdr, int64_t n_tasks,
                                      const int32_t *task_hdr, int64_t n_items, const uint16_t *item_hdr,
                                      int64_t ell_size, const uint32_t *ell, int max_tile_blocks,
                                      const int32_t *chunk_cells_ptr, int64_t n_chunk_cells, const int32_t *chunk_cells,
                                      nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_set_p2_items: plan is NULL");
    if (bp->n_phi != 10 || bp->n_comp != 3) return fail("nfa_bplan_set_p2_items: plan is not vector-P2 on tetrahedra");
    if (n_chunks <= 0 || n_tasks <= 0 || n_items <= 0 || ell_size <= 0 || max_tile_blocks <= 0)
        return fail("nfa_bplan_set_p2_items: empty item plan");
    if ((int64_t)max_tile_blocks * 9 + 2 >= 65535) return fail("nfa_bplan_set_p2_items: tile exceeds 16-bit offsets");
    cudaStream_t s = (cudaStream_t)stream;
    nfa_bplan::P2Items &it = bp->p2;
    if (it.chunk_hdr.alloc(16 * n_chunks) || it.task_hdr.alloc(16 * n_tasks) || it.item_hdr.alloc(8 * (n_items + 2)) ||
        it.ell.alloc(4 * ell_size) || it.rec.alloc(sizeof(double) * 16 * bp->n_cells) ||
        it.cells_ptr.alloc(4 * (n_chunks + 1)) || it.cells.alloc(4 * std::max<int64_t>(n_chunk_cells, 1)))
        return 2;
    if (!chunk_cells_ptr || !chunk_cells) return fail("nfa_bplan_set_p2_items: prefetch lists missing");
    NFA_CUDA(cudaMemcpyAsync(it.cells_ptr.p, chunk_cells_ptr, 4 * (n_chunks + 1), cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.cells.p, chunk_cells, 4 * n_chunk_cells, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.chunk_hdr.p, chunk_hdr, 16 * n_chunks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.task_hdr.p, task_hdr, 16 * n_tasks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.item_hdr.p, item_hdr, 8 * n_items, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.ell.p, ell, 4 * ell_size, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    it.n_chunks = n_chunks;
    it.n_tasks = n_tasks;
    it.n_items = n_items;
    it.ell_size = ell_size;
    it.max_tile_blocks = max_tile_blocks;
    return 0;
}

extern "C" int nfa_bplan_has_p2_items(const nfa_bplan *bp) { return bp && bp->p2.n_chunks > 0 ? 1 : 0; }

extern "C" int nfa_bplan_set_cell_mask(nfa_bplan *bp, const uint8_t *mask, nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_set_cell_mask: plan is NULL");
    nfa_bplan::P2Items &it = bp->p2;
    if (!mask) {
        it.cell_mask.release();
        return 0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (it.cell_mask.alloc(bp->n_cells)) return 2;
    NFA_CUDA(cudaMemcpyAsync(it.cell_mask.p, mask, bp->n_cells, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int nfa_bplan_set_p2_stage(nfa_bplan *bp, const int32_t *chunk_ext, const int32_t *task_hdr_rel,
                                      const uint32_t *ell_local, int max_chunk_cells, int max_chunk_ell,
                                      int max_chunk_items, int max_chunk_tasks, nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_set_p2_stage: plan is NULL");
    nfa_bplan::P2Items &it = bp->p2;
    if (it.n_chunks <= 0) return fail("nfa_bplan_set_p2_stage: set the row items first (nfa_bplan_set_p2_items)");
    if (!chunk_ext || !task_hdr_rel || !ell_local) return fail("nfa_bplan_set_p2_stage: NULL list");
    if (max_chunk_cells <= 0 || max_chunk_ell <= 0 || max_chunk_items <= 0 || max_chunk_tasks <= 0)
        return fail("nfa_bplan_set_p2_stage: empty chunk bounds");
    cudaStream_t s = (cudaStream_t)stream;
    if (it.chunk_ext.alloc(16 * it.n_chunks) || it.task_hdr_rel.alloc(16 * it.n_tasks) || it.ell_local.alloc(4 * it.ell_size))
        return 2;
    NFA_CUDA(cudaMemcpyAsync(it.chunk_ext.p, chunk_ext, 16 * it.n_chunks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.task_hdr_rel.p, task_hdr_rel, 16 * it.n_tasks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.ell_local.p, ell_local, 4 * it.ell_size, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    it.max_chunk_cells = max_chunk_cells;
    it.max_chunk_ell = max_chunk_ell;
    it.max_chunk_items = max_chunk_items;
    it.max_chunk_tasks = max_chunk_tasks;
    return 0;
}

// The staged kernel when the plan carries its lists and at least two chunk buffers fit in shared memory.
static int launch_p2tet_staged(const nfa_bplan *bp, double *vals, bool cr, cudaStream_t s, bool *launched) {
    const nfa_bplan::P2Items &it = bp->p2;
    *launched = false;
    if (it.max_chunk_cells <= 0) return 0;
    // opt-in (NFA_P2TET_STAGE=1): measured 1.23 ms against 0.95 ms for the global-load kernel on the bench workload --
    // the random 128-bit shared-memory reads of the records cost 3x the LSU wavefronts of the 256-bit global loads and
    // the in-order hand-over of filled buffers leaves the warps spinning (profiles/r01_p2tet_staged_ncu_summary.txt)
    const char *e_stage = getenv("NFA_P2TET_STAGE");
    if (!e_stage || atoi(e_stage) == 0) return 0;
    auto r128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    StageLayout L;
    const size_t tile_bytes = r128(sizeof(double) * (9 * (size_t)it.max_tile_blocks + 2));
    L.rec_off = (int)tile_bytes;
    L.ell_off = L.rec_off + 128 * it.max_chunk_cells;
    L.item_off = L.ell_off + (int)r128(4 * (size_t)it.max_chunk_ell);
    L.task_off = L.item_off + (int)r128(8 * (size_t)it.max_chunk_items);
    L.buf_bytes = L.task_off + (int)r128(16 * (size_t)it.max_chunk_tasks);
    const size_t ctrl_bytes = 256, limit = 227 * 1024;
    int nbuf = (int)std::min<size_t>(8, (limit - ctrl_bytes) / (size_t)L.buf_bytes);
    if (const char *e = getenv("NFA_P2TET_NBUF")) nbuf = std::min(nbuf, std::max(1, atoi(e)));
    if (nbuf < 2) return 0;  // chunks too large for two buffers: the caller falls back to the global-load kernel
    const size_t smem = (size_t)nbuf * L.buf_bytes + ctrl_bytes;
    int threads = 512, rec_tma = 0;
    if (const char *e = getenv("NFA_P2TET_THREADS")) threads = atoi(e);
    if (const char *e = getenv("NFA_P2TET_REC_TMA")) rec_tma = atoi(e) != 0;
    using kern_t = void (*)(int, int, const int4 *, const int4 *, const int4 *, const uint2 *, const uint32_t *,
                            const int32_t *, const int32_t *, const double *, StageLayout, int, double *);
    kern_t kern = nullptr;
#define NFA_P2S_PICK(CRV, TH) \
    if (cr == CRV && threads == TH) kern = p2tet_staged_kernel<CRV, TH>;
    NFA_P2S_PICK(false, 512) NFA_P2S_PICK(false, 640) NFA_P2S_PICK(false, 768) NFA_P2S_PICK(false, 1024)
    NFA_P2S_PICK(true, 512) NFA_P2S_PICK(true, 640) NFA_P2S_PICK(true, 768) NFA_P2S_PICK(true, 1024)
#undef NFA_P2S_PICK
    if (!kern) return fail("nfa_assemble_p2tet_elasticity: NFA_P2TET_THREADS must be 512, 640, 768 or 1024 (staged kernel)");
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    int64_t grid = std::min<int64_t>(it.n_chunks, (int64_t)n_sm);
    if (const char *e = getenv("NFA_P2TET_GRID")) grid = std::max<int64_t>(1, std::min<int64_t>(grid, atoi(e)));  // tests
    if (timing_enabled()) timing_begin(s);
    kern<<<(unsigned)grid, threads, smem, s>>>(nbuf, (int)it.n_chunks, (const int4 *)it.chunk_hdr.p,
                                               (const int4 *)it.chunk_ext.p, (const int4 *)it.task_hdr_rel.p,
                                               (const uint2 *)it.item_hdr.p, (const uint32_t *)it.ell_local.p,
                                               (const int32_t *)it.cells_ptr.p, (const int32_t *)it.cells.p,
                                               (const double *)it.rec.p, L, rec_tma, vals);
    if (timing_enabled()) timing_end(s);
    NFA_LAUNCH_CHECK();
    *launched = true;
    return 0;
}

extern "C" int nfa_assemble_p2tet_elasticity(const nfa_bplan *bp, const double *coords, const int32_t *cell_nodes,
                                             const double *lam, int64_t ls, const double *mu, int64_t ms, double *vals,
                                             nfa_stream stream) {
    if (!bp) return fail("nfa_assemble_p2tet_elasticity: plan is NULL");
    const nfa_bplan::P2Items &it = bp->p2;
    if (it.n_chunks <= 0) return fail("nfa_assemble_p2tet_elasticity: the plan carries no row items");
    cudaStream_t s = (cudaStream_t)stream;
    p2tet_geo_kernel<<<grid_for(bp->n_cells, 128), 128, 0, s>>>(coords, cell_nodes, bp->n_cells, lam, ls, mu, ms,
                                                                (const uint8_t *)it.cell_mask.p, (double *)it.rec.p);
    NFA_LAUNCH_CHECK();
    {
        bool launched = false;
        if (int rc = launch_p2tet_staged(bp, vals, ls == 0 && ms == 0, s, &launched)) return rc;
        if (launched) return 0;
    }
    const int tile_stride = (9 * it.max_tile_blocks + 2 + 15) & ~15;  // doubles; keeps both tiles 128-byte aligned
    // as many tiles as fit in 220 KB of shared memory (at most 8): the more chunks are in flight, the less a long task
    // (a vertex-star item walks ~24 cells) delays the hand-over of its chunk's tile
    int nbuf = (int)std::min<size_t>(8, (220 * 1024) / (sizeof(double) * (size_t)tile_stride));
    if (const char *e = getenv("NFA_P2TET_NBUF")) nbuf = std::min(nbuf, std::max(1, atoi(e)));
    if (nbuf < 1) return fail("nfa_assemble_p2tet_elasticity: chunk tile exceeds the shared memory of an SM");
    const size_t smem = sizeof(double) * nbuf * (size_t)tile_stride + sizeof(int) * (4 + 2 * nbuf);
    // both coefficients constant over the mesh: rho = lambda / mu is one number and Kf is applied once per block
    const bool cr = (ls == 0 && ms == 0);
    int threads = P2_THREADS;  // 512 (128 registers): more warps spill and lose (1.5 ms), see DESIGN.md
    if (const char *e = getenv("NFA_P2TET_THREADS")) threads = atoi(e);
    if (threads > 640) threads = 512;  // larger counts are staged-kernel settings
    using kern_t = void (*)(int, int, const int4 *, const int4 *, const uint2 *, const uint32_t *, const double *, int,
                            const int32_t *, const int32_t *, double *);
    kern_t kern = nullptr;
    const bool nb4 = nbuf == 4 && !getenv("NFA_P2TET_RT_NBUF");
#define NFA_P2_PICK(CRV, TH)                                                      \
    if (cr == CRV && threads == TH)                                               \
        kern = nb4 ? p2tet_elasticity_kernel<CRV, TH, 4> : p2tet_elasticity_kernel<CRV, TH, 0>;
    NFA_P2_PICK(false, 384) NFA_P2_PICK(false, 448) NFA_P2_PICK(false, 512) NFA_P2_PICK(false, 576) NFA_P2_PICK(false, 640)
    NFA_P2_PICK(true, 384) NFA_P2_PICK(true, 448) NFA_P2_PICK(true, 512) NFA_P2_PICK(true, 576) NFA_P2_PICK(true, 640)
#undef NFA_P2_PICK
    if (!kern) return fail("nfa_assemble_p2tet_elasticity: NFA_P2TET_THREADS must be 384, 448, 512, 576 or 640");
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail("nfa_assemble_p2tet_elasticity: kernel does not fit on an SM");
    const int64_t grid = std::min<int64_t>(it.n_chunks, (int64_t)n_sm * per_sm);
    if (timing_enabled()) timing_begin(s);
    kern<<<(unsigned)grid, threads, smem, s>>>(nbuf, (int)it.n_chunks, (const int4 *)it.chunk_hdr.p,
                                                  (const int4 *)it.task_hdr.p, (const uint2 *)it.item_hdr.p,
                                                  (const uint32_t *)it.ell.p, (const double *)it.rec.p, tile_stride,
                                                  (const int32_t *)it.cells_ptr.p, (const int32_t *)it.cells.p, vals);
    if (timing_enabled()) timing_end(s);
    NFA_LAUNCH_CHECK();
    return 0;
}
