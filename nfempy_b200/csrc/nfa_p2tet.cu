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
// p2_items.py lists the five item types).  One thread owns one item, walks its cells in ascending cell order -- the
// order of the reference's Mat.setValue(..., addv=True) calls, R/src/assembly/ScatterFormData.py:5-11 -- sums in
// registers and writes every CSR value exactly once into the CTA's shared-memory tile (the CSR image of a run of
// whole block rows), which leaves for HBM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, TMA engine).
// Inputs per (item, cell): one 32-bit packed entry (coalesced, ELL per warp) and one 32-byte record per vertex used,
// fetched with 256-bit loads (LDG.E.256) from the per-cell geometry records (L1/L2 resident: 128 B per cell).
#include <algorithm>
#include <cstdlib>
#include <vector>

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
//   NFA_P2_STREAM_PF  1: the plan lists a CTA reads (task headers, ELL entries, item headers) are contiguous in task order;
//                  every task prefetches into L2 the part of each list that lies a fixed distance AHEAD of what it
//                  consumes (a software stream prefetcher: no further tasks are claimed, so the tile window is not
//                  overrun, and the look-ahead loads of the next task become L2 hits instead of DRAM misses)
#ifndef NFA_P2_STREAM_PF
#define NFA_P2_STREAM_PF 0  // measured: no gain (0.778 with, 0.772 without; 0.765 against 0.768 with L1NA = 1)
#endif
#ifndef NFA_P2_PFD_ELL
#define NFA_P2_PFD_ELL 4096  // look-ahead in ELL entries (16 KB, about 35 tasks)
#endif
#ifndef NFA_P2_PFD_IH
#define NFA_P2_PFD_IH 1280  // look-ahead in item-header slots (10 KB)
#endif
#ifndef NFA_P2_PFD_TH
#define NFA_P2_PFD_TH 64  // look-ahead in task headers
#endif
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
#ifndef NFA_P2_UA_LD
#define NFA_P2_UA_LD 5  // the same when the items also gather the load vectors (two records per cell)
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
//   NFA_P2_L1NA    bit 0: ELL entries, bit 1: task headers, bit 2: item headers are loaded L1::no_allocate -- every line of
//                  these lists is read exactly once, and the 31 KB of L1 left beside the tiles is better spent on records
#ifndef NFA_P2_L1NA
#define NFA_P2_L1NA 1  // measured: ELL no-allocate 0.768 ms against 0.778; headers as well: 0.772 / 0.782
#endif
__device__ __forceinline__ uint32_t ld_list(const uint32_t *p) {
#if (NFA_P2_L1NA & 1)
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#elif (NFA_P2_L2HINT & 4)
    uint32_t v;
    const uint64_t pol = l2_policy_evict_first();
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ int4 ld_list(const int4 *p) {
#if (NFA_P2_L1NA & 2)
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#elif (NFA_P2_L2HINT & 4)
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
#if (NFA_P2_L1NA & 4)
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#elif (NFA_P2_L2HINT & 4)
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
    unsigned ihz;       // type E: tile offsets of the opposite-edge blocks of the item's first / second cell; type D: row node
    double *rhs;        // right-hand side (items A and D own the three rows of their node), or NULL
    int rhs_overwrite;  // 1: the sums replace the rows (caller: rhs is zero or every node has cells), 0: they are added
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
    __device__ __forceinline__ void put_opposite(int cell, const double (&v)[9]) const {
        store_block(tile, cell == 0 ? (ihz & 0xFFFFu) : (ihz >> 16), ihy >> 16, v);
    }
    // the load-vector sums of row node `node` (ascending cell order, like the reference's res_g[dest] += r_el,
    // R/src/assembly/ScatterFormData.py:14-18): every row has exactly one owner thread, so plain stores / read-add-store
    __device__ __forceinline__ void put_rhs(unsigned node, const double (&f)[3]) const {
        double *r = rhs + 3 * (size_t)node;
        if (rhs_overwrite) {
            r[0] = f[0];
            r[1] = f[1];
            r[2] = f[2];
        } else {
            const double o0 = r[0], o1 = r[1], o2 = r[2];
            r[0] = o0 + f[0];
            r[1] = o1 + f[1];
            r[2] = o2 + f[2];
        }
    }
};

// ---- per item type: which vertices' records an entry needs, how a cell adds to the sums, how the sums become blocks
// LDA: 0 = no load vector, 1 = the row node is vertex i (items A), 2 = the row node is the edge (i, j) (items D): the
// cell's local load vector of that node, re4 [n_cells][10][4] (three components + one pad: one 32-byte sector), rides
// along as record NG
template <int NG, int LDA = 0>
struct Rec {
    G4 g[NG + (LDA ? 1 : 0)];
    bool valid;
};

template <int NG, int LDA = 0>
__device__ __forceinline__ Rec<NG, LDA> load_rec(uint32_t code, const double *__restrict__ rec,
                                                 const double *__restrict__ re4 = nullptr) {
    Rec<NG, LDA> r;
    r.valid = code != 0xFFFFFFFFu;
#pragma unroll
    for (int m = 0; m < NG + (LDA ? 1 : 0); ++m) r.g[m] = G4{0.0, 0.0, 0.0, 0.0};
    if (r.valid) {
        const double *rc = rec + (size_t)(code >> 6) * 16;
        const unsigned i = (code >> 4) & 3u, j = (code >> 2) & 3u;
        r.g[0] = ldg256(rc + 4 * i);
        if constexpr (NG > 1) r.g[1] = ldg256(rc + 4 * j);
        if constexpr (NG > 2) r.g[2] = ldg256(rc + 4 * (code & 3u));
        if constexpr (LDA != 0) {
            // local P2 node: vertex i, or the edge node of (i, j) in the order (2,3),(1,3),(1,2),(0,3),(0,2),(0,1) = 4..9
            const unsigned a = LDA == 1 ? i : 10u - (i + j) - ((i < j ? i : j) >= 1u ? 1u : 0u);
            r.g[NG] = ldg256(re4 + ((size_t)(code >> 6) * 10 + a) * 4);
        }
    }
    return r;
}

constexpr double C10 = 0.1, C30 = 1.0 / 30.0, C15 = 2.0 / 15.0;

template <int TYPE, bool CR, bool LD>
struct Item;

template <bool CR, bool LD>
struct Item<0, CR, LD> {
    double rho = 0.0;  // A: (v, v) = 0.1 Kf(g_i g_i^T); LD: + the three right-hand-side rows of the vertex
    static constexpr int NG = 1, LDA = LD ? 1 : 0;
    double q[6] = {0, 0, 0, 0, 0, 0}, tr = 0.0;
    double f[3] = {0, 0, 0};
    __device__ __forceinline__ void add(const Rec<NG, LDA> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z};
        tr += add_sym<CR>(q, r.g[0].rho, u);
        if constexpr (LD) {
            f[0] += r.g[NG].x;
            f[1] += r.g[NG].y;
            f[2] += r.g[NG].z;
        }
    }
    __device__ __forceinline__ void finish(const TileCtx &ctx) {
        double o[9];
        fix_sym<CR>(q, tr, rho);
        sym_to_full(q, C10, tr, o);
        ctx.put(0, o);
        // the vertex's node id rides in the two unused block slots of the item header
        if constexpr (LD)  // (lanes beyond the task's items carry an all-ones header)
            if ((ctx.ihx & 0xFFFFu) != 0xFFFFu) ctx.put_rhs((ctx.ihx >> 16) | ((ctx.ihy & 0xFFFFu) << 16), f);
    }
};

template <bool CR, bool LD>
struct Item<1, CR, LD> {
    double rho = 0.0;  // B: (v, v') = -1/30 P, (v, E) = 0.1 P - Q / 30 with P = Kf(g_i (x) g_k), Q = Kf(g_i g_i^T)
    static constexpr int NG = 2, LDA = 0;
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

template <bool CR, bool LD>
struct Item<2, CR, LD> {
    double rho = 0.0;  // C: (v, e_xy) = Kf(g_i (x) (g_i + g_z)) / 30, since g_x + g_y = -(g_i + g_z)
    static constexpr int NG = 2, LDA = 0;
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

template <bool CR, bool LD>
struct Item<3, CR, LD> {
    double rho = 0.0;  // D: (E, v_i) = 0.1 P^T - Qi / 30, (E, v_j) = 0.1 P - Qj / 30, (E, E) = 2/15 (2 Qi + 2 Qj + P + P^T)
    static constexpr int NG = 2, LDA = LD ? 2 : 0;  // LD: + the three right-hand-side rows of the edge node
    double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, qi[6] = {0, 0, 0, 0, 0, 0}, qj[6] = {0, 0, 0, 0, 0, 0};
    double trp = 0.0, tri = 0.0, trj = 0.0;
    double f[3] = {0, 0, 0};
    __device__ __forceinline__ void add(const Rec<NG, LDA> &r) {
        rho = r.g[0].rho;
        const double u[3] = {r.g[0].x, r.g[0].y, r.g[0].z}, w[3] = {r.g[1].x, r.g[1].y, r.g[1].z};
        trp += add_general<CR>(p, r.g[0].rho, u, w);
        tri += add_sym<CR>(qi, r.g[0].rho, u);
        trj += add_sym<CR>(qj, r.g[0].rho, w);
        if constexpr (LD) {
            f[0] += r.g[NG].x;
            f[1] += r.g[NG].y;
            f[2] += r.g[NG].z;
        }
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
        if constexpr (LD)  // second header slot of a type D item: the edge's node id
            if ((ctx.ihx & 0xFFFFu) != 0xFFFFu) ctx.put_rhs(ctx.ihz, f);
    }
};

template <bool CR, bool LD>
struct Item<4, CR, LD> {
    double rho = 0.0;  // E: (E, v_x), (E, e_ix), (E, e_jx) with s = g_i + g_j
    static constexpr int NG = 3, LDA = 0;
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

// One task = up to 32 items of one type; lane = item.  e points at the lane's first ELL entry (stride 32); c0, c1 are
// the lane's first two entries (fetched while the previous task was computed).  Items that read at most two records per
// cell keep two cells in flight; type E (three records, 27 sums) walks one cell at a time to stay inside the register
// budget.  pf0, pf1: first entries of the NEXT task -- their cells' records (one 128-byte line each) are pulled into L1
// once this task's sums are complete.
template <int TYPE, bool CR, bool LD, int UNROLL_NG>
__device__ __forceinline__ void run_task(const uint32_t *__restrict__ e, int kmax, uint32_t c0, uint32_t c1,
                                         const double *__restrict__ rec, const double *__restrict__ re4, uint32_t pf0,
                                         uint32_t pf1, const TileCtx &ctx) {
    constexpr int NG = Item<TYPE, CR, LD>::NG, LDA = Item<TYPE, CR, LD>::LDA;
    // cells in flight per iteration: long items with few sums walk many cells at once (their task would otherwise
    // hold its chunk's tile for tens of dependent load latencies); type E stays at one (register budget)
    constexpr int U = (TYPE == 0) ? (LD ? NFA_P2_UA_LD : NFA_P2_UA) : ((TYPE == 1) ? NFA_P2_UB : ((NG <= UNROLL_NG) ? 2 : 1));
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
        // (E, E_opp) of the cells that name this face as their designated one: a single contribution each,
        // -2/15 Kf(s s^T) with s = g_i + g_j (the opposite edge's g_x + g_y = -s)
        if ((ctx.ihz & 0xFFFFu) != 0xFFFFu) {
            double q[6] = {0, 0, 0, 0, 0, 0}, o[9];
            double t = add_sym<CR>(q, r0.g[0].rho, s0);
            fix_sym<CR>(q, t, r0.g[0].rho);
            sym_to_full(q, -C15, t, o);
            ctx.put_opposite(0, o);
        }
        if ((ctx.ihz >> 16) != 0xFFFFu) {
            double q[6] = {0, 0, 0, 0, 0, 0}, o[9];
            double t = add_sym<CR>(q, r1.g[0].rho, s1);
            fix_sym<CR>(q, t, r1.g[0].rho);
            sym_to_full(q, -C15, t, o);
            ctx.put_opposite(1, o);
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
    Item<TYPE, CR, LD> item;
    if constexpr (U > 2) {
        uint32_t c[U];
        c[0] = c0;
        c[1] = c1;
#pragma unroll
        for (int u = 2; u < U; ++u) c[u] = (u < kmax) ? ld_list(e + 32 * u) : 0xFFFFFFFFu;
#pragma unroll 1
        for (int k = 0; k < kmax; k += U) {
            Rec<NG, LDA> r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) r[u] = load_rec<NG, LDA>(c[u], rec, re4);
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = (k + U + u < kmax) ? ld_list(e + 32 * (k + U + u)) : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r[u].valid) item.add(r[u]);
        }
    } else if constexpr (U == 2) {
#pragma unroll 1
        for (int k = 0; k < kmax; k += 2) {
            const Rec<NG, LDA> r0 = load_rec<NG, LDA>(c0, rec, re4), r1 = load_rec<NG, LDA>(c1, rec, re4);
            c0 = (k + 2 < kmax) ? ld_list(e + 32 * (k + 2)) : 0xFFFFFFFFu;
            c1 = (k + 3 < kmax) ? ld_list(e + 32 * (k + 3)) : 0xFFFFFFFFu;
            if (r0.valid) item.add(r0);
            if (r1.valid) item.add(r1);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < kmax; ++k) {
            const Rec<NG, LDA> r0 = load_rec<NG, LDA>(c0, rec, re4);
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
// LD: items A and D also sum the padded local load vectors re4 [n_cells][10][4] of their node (written by
// nfa_local_h1_load_pitched) and write the node's three right-hand-side rows -- the row owner walks exactly the cells the
// rows need, so the separate gather / reduce pass over the local load vectors is gone.
template <bool CR, int THREADS, int NB, bool LD>
__global__ void __launch_bounds__(THREADS, 1)
    p2tet_elasticity_kernel(int nbuf_rt, int n_chunks, const int4 *__restrict__ chunk_hdr, const int4 *__restrict__ task_hdr,
                            const uint2 *__restrict__ item_hdr, const uint32_t *__restrict__ ell,
                            const double *__restrict__ rec, int tile_stride, double *__restrict__ vals,
                            unsigned ell_size, unsigned n_slots, int n_tasks, const double *__restrict__ re4,
                            double *__restrict__ rhs, int rhs_overwrite, const int *__restrict__ cta_ptr) {
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
        // chunk range of this CTA: equal chunk counts, or the plan's work-balanced split (nfa_bplan_set_p2_split)
        const int c_begin = cta_ptr ? __ldg(cta_ptr + blockIdx.x) : (int)((int64_t)n_chunks * blockIdx.x / gridDim.x);
        const int c_end = cta_ptr ? __ldg(cta_ptr + blockIdx.x + 1) : (int)((int64_t)n_chunks * (blockIdx.x + 1) / gridDim.x);
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
            const int per2 = ((th2.x & 0xFF) >= 3) ? 2 : 1;
            if ((lane & (15 >> (per2 - 1))) == 0 && lane < (th2.w & 63))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(item_hdr + th2.z + per2 * lane));
        }
#else
        int4 th2 = none;
        {
            const int t2 = grab_ahead();
            if (t2 < task_end) th2 = ld_list(task_hdr + t2);
#if NFA_P2_STREAM_PF
            if (lane == 0 && (t2 & 7) == 0 && t2 + NFA_P2_PFD_TH < n_tasks)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(task_hdr + t2 + NFA_P2_PFD_TH));
#endif
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
            {
                const int per1 = ((th1.x & 0xFF) >= 3) ? 2 : 1;  // type D and E items take two 8-byte slots
                if ((lane & (15 >> (per1 - 1))) == 0 && lane < (th1.w & 63))
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(item_hdr + th1.z + per1 * lane));
            }
#endif
        }
#if NFA_P2_STREAM_PF
        {
            // as much of each list as this task consumes, a fixed distance further on in the stream
            const int kmax0 = (th.x >> 8) & 0xFF;
            const unsigned pe = (unsigned)th.y + NFA_P2_PFD_ELL + 32u * (unsigned)lane;
            if (lane < kmax0 && pe < ell_size) asm volatile("prefetch.global.L2 [%0];" ::"l"(ell + pe));
            const int per0 = ((th.x & 0xFF) >= 3) ? 2 : 1;
            const unsigned ps = (unsigned)th.z + NFA_P2_PFD_IH + 16u * (unsigned)lane;  // 16 slots = one 128-byte line
            if (16 * lane < per0 * (th.w & 63) && ps < n_slots) asm volatile("prefetch.global.L2 [%0];" ::"l"(item_hdr + ps));
        }
#endif
        {
            const int type = th.x & 0xFF, kmax = (th.x >> 8) & 0xFF, n_items = th.w & 63, c = (int)((unsigned)th.w >> 6);
            const uint32_t *e = ell + (uint32_t)th.y + lane;
            uint2 ih = make_uint2(0xFFFFFFFFu, 0xFFFFu);
            unsigned ihz = 0xFFFFFFFFu;
            if (lane < n_items) {
                if (type >= 3) {  // two slots per item: blocks + pitch, then the opposite-edge blocks (E) / the row node (D)
                    const uint2 *hp = item_hdr + th.z + 2 * lane;
                    ih = ld_list(hp);
                    ihz = ld_list(hp + 1).x;
                } else {
                    ih = ld_list(item_hdr + th.z + lane);
                }
            }
            const int it = c - c_begin_r, b = it % NBUF;
            // the tile sits on the same 16-byte phase as its image in vals (parity of the chunk's first block)
            double *tile = smem + (size_t)b * tile_stride + ((th.x >> 16) & 1);
            const TileCtx ctx = {tile, ctrl + 4 + NBUF + b, it / NBUF, ih.x, ih.y, ihz, rhs, rhs_overwrite};
            constexpr int UNG = THREADS <= 384 ? 3 : P2_UNROLL_NG;
            switch (type) {
                case 0: run_task<0, CR, LD, UNG>(e, kmax, c0, c1, rec, re4, n0, n1, ctx); break;
                case 1: run_task<1, CR, LD, UNG>(e, kmax, c0, c1, rec, re4, n0, n1, ctx); break;
                case 2: run_task<2, CR, LD, UNG>(e, kmax, c0, c1, rec, re4, n0, n1, ctx); break;
                case 3: run_task<3, CR, LD, UNG>(e, kmax, c0, c1, rec, re4, n0, n1, ctx); break;
                default: run_task<4, CR, LD, UNG>(e, kmax, c0, c1, rec, re4, n0, n1, ctx); break;
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
                                      const int32_t *task_hdr, int64_t n_slots, const uint16_t *item_hdr,
                                      int64_t ell_size, const uint32_t *ell, int max_tile_blocks, nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_set_p2_items: plan is NULL");
    if (bp->n_phi != 10 || bp->n_comp != 3) return fail("nfa_bplan_set_p2_items: plan is not vector-P2 on tetrahedra");
    if (n_chunks <= 0 || n_tasks <= 0 || n_slots <= 0 || ell_size <= 0 || max_tile_blocks <= 0)
        return fail("nfa_bplan_set_p2_items: empty item plan");
    if ((int64_t)max_tile_blocks * 9 + 2 >= 65535) return fail("nfa_bplan_set_p2_items: tile exceeds 16-bit offsets");
    cudaStream_t s = (cudaStream_t)stream;
    nfa_bplan::P2Items &it = bp->p2;
    if (it.chunk_hdr.alloc(16 * n_chunks) || it.task_hdr.alloc(16 * n_tasks) || it.item_hdr.alloc(8 * (n_slots + 2)) ||
        it.ell.alloc(4 * ell_size) || it.rec.alloc(sizeof(double) * 16 * bp->n_cells))
        return 2;
    NFA_CUDA(cudaMemcpyAsync(it.chunk_hdr.p, chunk_hdr, 16 * n_chunks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.task_hdr.p, task_hdr, 16 * n_tasks, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.item_hdr.p, item_hdr, 8 * n_slots, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaMemcpyAsync(it.ell.p, ell, 4 * ell_size, cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    it.n_chunks = n_chunks;
    it.n_tasks = n_tasks;
    it.n_items = n_slots;
    it.ell_size = ell_size;
    it.max_tile_blocks = max_tile_blocks;
    return 0;
}

// ---- bank-aware lane order ------------------------------------------------------------------------------------------
// The items of a (chunk, type) group become the lanes of consecutive tasks, 16 to a half warp.  A half warp's tile stores
// (STS.64, nine per block) serialise when two lanes hit the same bank pair; with the items in first-cell order the
// positions are effectively random (2.6 wavefronts per store instead of 1).  One thread per group deals the items greedily:
// the next lane takes, among the next `window` unassigned items, the one that raises the fewest per-position maxima of the
// half warp's bank counts (ties: smaller sum of counts, then first in order).  nfempy_b200/assembly/p2_items.py
// greedy_lane_order is the same algorithm on the host (tests).
namespace {
__global__ void p2_lane_order_kernel(int64_t n_groups, const int64_t *__restrict__ group_ptr, const int32_t *__restrict__ hdr6,
                                     int window, int64_t *__restrict__ perm) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int64_t g0 = group_ptr[g], g1 = group_ptr[g + 1];
    for (int64_t k = g0; k < g1; ++k) perm[k] = k;  // perm[pos ..] = the unassigned items, in order
    unsigned char cnt[15][16], mx[15];
    const int cols[5] = {0, 1, 2, 4, 5};
    for (int64_t pos = g0; pos < g1; ++pos) {
        if (((pos - g0) & 15) == 0) {
            for (int p = 0; p < 15; ++p) {
                mx[p] = 0;
                for (int b = 0; b < 16; ++b) cnt[p][b] = 0;
            }
        }
        const int64_t c_end = (pos + window < g1) ? pos + window : g1;
        int64_t best = pos;
        int best_key = 0x7fffffff;
        for (int64_t c = pos; c < c_end; ++c) {
            const int32_t *h = hdr6 + 6 * perm[c];
            const int pitch = h[3];
            int inc = 0, tie = 0;
            for (int s = 0; s < 5; ++s) {
                const int w = h[cols[s]];
                if (w == 0xFFFF) continue;
                for (int r = 0; r < 3; ++r) {
                    const int cc = cnt[3 * s + r][(w + r * pitch) & 15];
                    inc += (cc + 1 > mx[3 * s + r]) ? 1 : 0;
                    tie += cc;
                }
            }
            const int key = inc * 65536 + tie;
            if (key < best_key) {
                best_key = key;
                best = c;
            }
        }
        const int64_t chosen = perm[best];
        for (int64_t c = best; c > pos; --c) perm[c] = perm[c - 1];
        perm[pos] = chosen;
        const int32_t *h = hdr6 + 6 * chosen;
        const int pitch = h[3];
        for (int s = 0; s < 5; ++s) {
            const int w = h[cols[s]];
            if (w == 0xFFFF) continue;
            for (int r = 0; r < 3; ++r) {
                unsigned char &cc = cnt[3 * s + r][(w + r * pitch) & 15];
                ++cc;
                if (cc > mx[3 * s + r]) mx[3 * s + r] = cc;
            }
        }
    }
}
}  // namespace

extern "C" int nfa_p2_lane_order(int64_t n_groups, const int64_t *group_ptr, const int32_t *hdr6, int window,
                                 int64_t *perm, nfa_stream stream) {
    if (n_groups <= 0) return 0;
    if (!group_ptr || !hdr6 || !perm) return fail("nfa_p2_lane_order: NULL argument");
    if (window < 1) return fail("nfa_p2_lane_order: window must be positive");
    p2_lane_order_kernel<<<grid_for(n_groups, 64), 64, 0, (cudaStream_t)stream>>>(n_groups, group_ptr, hdr6, window, perm);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_bplan_set_p2_split(nfa_bplan *bp, int n_cta, const int32_t *cta_ptr, nfa_stream stream) {
    if (!bp) return fail("nfa_bplan_set_p2_split: plan is NULL");
    nfa_bplan::P2Items &it = bp->p2;
    if (it.n_chunks <= 0) return fail("nfa_bplan_set_p2_split: the plan carries no row items");
    it.n_split = 0;
    if (n_cta <= 0 || !cta_ptr) return 0;  // back to equal chunk counts
    std::vector<int32_t> h(n_cta + 1);
    cudaStream_t s = (cudaStream_t)stream;
    NFA_CUDA(cudaMemcpyAsync(h.data(), cta_ptr, 4 * (n_cta + 1), cudaMemcpyDeviceToHost, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    if (h[0] != 0 || h[n_cta] != it.n_chunks) return fail("nfa_bplan_set_p2_split: the split must cover [0, n_chunks]");
    for (int b = 0; b < n_cta; ++b)
        if (h[b] > h[b + 1]) return fail("nfa_bplan_set_p2_split: the split must be non-decreasing");
    if (it.cta_ptr.alloc(4 * (n_cta + 1))) return 2;
    NFA_CUDA(cudaMemcpyAsync(it.cta_ptr.p, cta_ptr, 4 * (n_cta + 1), cudaMemcpyDeviceToDevice, s));
    NFA_CUDA(cudaStreamSynchronize(s));
    it.n_split = n_cta;
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

static int assemble_p2tet(const nfa_bplan *bp, const double *coords, const int32_t *cell_nodes, const double *lam,
                          int64_t ls, const double *mu, int64_t ms, double *vals, const double *re4, double *rhs,
                          int rhs_overwrite, nfa_stream stream) {
    if (!bp) return fail("nfa_assemble_p2tet_elasticity: plan is NULL");
    const nfa_bplan::P2Items &it = bp->p2;
    if (it.n_chunks <= 0) return fail("nfa_assemble_p2tet_elasticity: the plan carries no row items");
    cudaStream_t s = (cudaStream_t)stream;
    p2tet_geo_kernel<<<grid_for(bp->n_cells, 128), 128, 0, s>>>(coords, cell_nodes, bp->n_cells, lam, ls, mu, ms,
                                                                (const uint8_t *)it.cell_mask.p, (double *)it.rec.p);
    NFA_LAUNCH_CHECK();
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
                            double *, unsigned, unsigned, int, const double *, double *, int, const int *);
    kern_t kern = nullptr;
    const bool nb4 = nbuf == 4 && !getenv("NFA_P2TET_RT_NBUF");
    const bool ld = re4 != nullptr && rhs != nullptr;
#define NFA_P2_PICK(CRV, TH)                                                                                          \
    if (cr == CRV && threads == TH)                                                                                   \
        kern = ld ? (nb4 ? p2tet_elasticity_kernel<CRV, TH, 4, true> : p2tet_elasticity_kernel<CRV, TH, 0, true>)     \
                  : (nb4 ? p2tet_elasticity_kernel<CRV, TH, 4, false> : p2tet_elasticity_kernel<CRV, TH, 0, false>);
    NFA_P2_PICK(false, 384) NFA_P2_PICK(false, 448) NFA_P2_PICK(false, 480) NFA_P2_PICK(false, 512) NFA_P2_PICK(false, 640)
    NFA_P2_PICK(true, 384) NFA_P2_PICK(true, 448) NFA_P2_PICK(true, 480) NFA_P2_PICK(true, 512) NFA_P2_PICK(true, 640)
#undef NFA_P2_PICK
    if (!kern) return fail("nfa_assemble_p2tet_elasticity: NFA_P2TET_THREADS must be 384, 448, 480, 512 or 640");
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail("nfa_assemble_p2tet_elasticity: kernel does not fit on an SM");
    int64_t grid = std::min<int64_t>(it.n_chunks, (int64_t)n_sm * per_sm);
    if (const char *e = getenv("NFA_P2TET_GRID")) grid = std::max<int64_t>(1, std::min<int64_t>(grid, atoi(e)));  // tests
    if (timing_enabled()) timing_begin(s);
    kern<<<(unsigned)grid, threads, smem, s>>>(nbuf, (int)it.n_chunks, (const int4 *)it.chunk_hdr.p,
                                                  (const int4 *)it.task_hdr.p, (const uint2 *)it.item_hdr.p,
                                                  (const uint32_t *)it.ell.p, (const double *)it.rec.p, tile_stride, vals,
                                                  (unsigned)it.ell_size, (unsigned)it.n_items, (int)it.n_tasks, re4, rhs,
                                                  rhs_overwrite, (it.n_split == grid) ? (const int *)it.cta_ptr.p : nullptr);
    if (timing_enabled()) timing_end(s);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_assemble_p2tet_elasticity(const nfa_bplan *bp, const double *coords, const int32_t *cell_nodes,
                                             const double *lam, int64_t ls, const double *mu, int64_t ms, double *vals,
                                             nfa_stream stream) {
    return assemble_p2tet(bp, coords, cell_nodes, lam, ls, mu, ms, vals, nullptr, nullptr, 0, stream);
}

extern "C" int nfa_assemble_p2tet_elasticity_load(const nfa_bplan *bp, const double *coords, const int32_t *cell_nodes,
                                                  const double *lam, int64_t ls, const double *mu, int64_t ms,
                                                  double *vals, const double *re4, double *rhs, int rhs_overwrite,
                                                  nfa_stream stream) {
    if (!re4 || !rhs) return fail("nfa_assemble_p2tet_elasticity_load: re4 and rhs must not be NULL");
    return assemble_p2tet(bp, coords, cell_nodes, lam, ls, mu, ms, vals, re4, rhs, rhs_overwrite, stream);
}
