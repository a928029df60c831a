// libnfa: error state, geometry (K1) and the quadrature-loop H1 local kernels (K2, K5).  sm_100a.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "nfa_plan.cuh"

namespace nfa {
static std::mutex g_err_mutex;
static std::string g_err = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const std::string &msg) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_err = msg;
}
void count_launch(int64_t n) { g_launches += n; }
}  // namespace nfa

using namespace nfa;

extern "C" int nfa_version(void) { return 100; }
extern "C" const char *nfa_last_error(void) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    static thread_local std::string copy;
    copy = g_err;
    return copy.c_str();
}
extern "C" int64_t nfa_launch_count(void) { return g_launches.load(); }

// ---- optional device timing of the dominant kernel (bench.py's roofline leg) -------------------------------
namespace nfa {
static bool g_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
bool timing_enabled() { return g_timing; }
void timing_begin(cudaStream_t s) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, s);
    g_events.push_back({a, b});
}
void timing_end(cudaStream_t s) { cudaEventRecord(g_events.back().second, s); }
}  // namespace nfa

extern "C" void nfa_timing_enable(int on) { g_timing = on != 0; }
extern "C" int nfa_timing_read(double *total_ms, int64_t *count) {
    double total = 0.0;
    for (auto &ev : g_events) {
        if (cudaEventSynchronize(ev.second) != cudaSuccess) return fail("nfa_timing_read: event sync failed");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev.first, ev.second);
        total += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (total_ms) *total_ms = total;
    if (count) *count = (int64_t)g_events.size();
    g_events.clear();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K1 geometry
template <int DIM>
__global__ void geom_kernel(const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes,
                            int64_t n_cells, double *__restrict__ jac, double *__restrict__ det,
                            double *__restrict__ inv) {
    int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    double X[DIM + 1][3];
    load_cell_coords<DIM>(coords, cell_nodes, cell, X);
    Geom<DIM> g;
    compute_geom<DIM>(X, g);
    if (jac)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < DIM; ++c) jac[(cell * 3 + r) * DIM + c] = g.J[r][c];
    if (det) det[cell] = g.det;
    if (inv)
#pragma unroll
        for (int c = 0; c < DIM; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) inv[(cell * DIM + c) * 3 + r] = g.Jinv[c][r];
}

extern "C" int nfa_geom_simplex(int dim, const double *coords, const int32_t *cell_nodes, int64_t n_cells,
                                double *jac, double *det, double *inv, nfa_stream stream) {
    if (n_cells == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int block = 128, grid = grid_for(n_cells, block);
    if (dim == 3)
        geom_kernel<3><<<grid, block, 0, s>>>(coords, cell_nodes, n_cells, jac, det, inv);
    else if (dim == 2)
        geom_kernel<2><<<grid, block, 0, s>>>(coords, cell_nodes, n_cells, jac, det, inv);
    else if (dim == 1)
        geom_kernel<1><<<grid, block, 0, s>>>(coords, cell_nodes, n_cells, jac, det, inv);
    else
        return fail("nfa_geom_simplex: dim must be 1, 2 or 3");
    NFA_LAUNCH_CHECK();
    return 0;
}

__global__ void map_points_kernel(int nv, int n_q, const double *__restrict__ shape1,
                                  const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes,
                                  int64_t n_cells, double *__restrict__ x) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_cells * n_q * 3;
    if (idx >= total) return;
    int r = (int)(idx % 3);
    int q = (int)((idx / 3) % n_q);
    int64_t cell = idx / (3 * (int64_t)n_q);
    double acc = 0.0;
    for (int v = 0; v < nv; ++v) {
        int64_t n = cell_nodes[cell * nv + v];
        acc += shape1[q * nv + v] * __ldg(coords + 3 * n + r);
    }
    x[idx] = acc;
}

extern "C" int nfa_map_points(int dim, int n_q, const double *shape1, const double *coords,
                              const int32_t *cell_nodes, int64_t n_cells, double *x, nfa_stream stream) {
    if (dim < 0 || dim > 3) return fail("nfa_map_points: dim must be 0..3");
    if (n_cells == 0) return 0;
    int64_t total = n_cells * n_q * 3;
    map_points_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dim + 1, n_q, shape1, coords,
                                                                               cell_nodes, n_cells, x);
    NFA_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K2: H1 volume forms with a quadrature loop (general coefficients).  One warp per cell.
// FORM 0: elasticity (n_comp = DIM), FORM 1: Laplace (n_comp = 1).
struct Coef {
    const double *p;
    int64_t sc, sq;
};

constexpr int WARPS_PER_CTA = 4;

template <int DIM, int FORM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
    local_h1_kernel(int n_phi, int n_q, const double *__restrict__ tab, const double *__restrict__ weights,
                    const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                    Coef c0, Coef c1, const double *__restrict__ f, const double *__restrict__ alpha,
                    const int32_t *__restrict__ dest, double *__restrict__ Ke, double *__restrict__ re) {
    constexpr int NC = (FORM == 0) ? DIM : 1;
    extern __shared__ double smem[];
    const int n_loc = n_phi * NC;
    // CTA-shared: tab [1+DIM][n_q][n_phi], weights [n_q]
    double *s_tab = smem;
    double *s_w = s_tab + (1 + DIM) * n_q * n_phi;
    double *s_warp = s_w + n_q;
    const int per_warp = n_q * n_phi * DIM + 2 * n_q + n_loc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s_g = s_warp + warp * per_warp;  // [n_q][n_phi][DIM] physical gradients
    double *s_c0 = s_g + n_q * n_phi * DIM;  // [n_q] w det * coef0 (lambda | kappa)
    double *s_c1 = s_c0 + n_q;               // [n_q] w det * coef1 (mu)
    double *s_a = s_c1 + n_q;                // [n_loc] local alpha

    for (int i = threadIdx.x; i < (1 + DIM) * n_q * n_phi; i += blockDim.x) s_tab[i] = tab[i];
    for (int i = threadIdx.x; i < n_q; i += blockDim.x) s_w[i] = weights[i];
    __syncthreads();

    for (int64_t cell = (int64_t)blockIdx.x * WARPS_PER_CTA + warp; cell < n_cells;
         cell += (int64_t)gridDim.x * WARPS_PER_CTA) {
        double X[DIM + 1][3];
        load_cell_coords<DIM>(coords, cell_nodes, cell, X);
        Geom<DIM> g;
        compute_geom<DIM>(X, g);
        __syncwarp();
        for (int i = lane; i < n_q * n_phi; i += 32) {
            int q = i / n_phi, a = i - q * n_phi;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) v += g.Jinv[k][d] * s_tab[((1 + k) * n_q + q) * n_phi + a];
                s_g[(q * n_phi + a) * DIM + d] = v;
            }
        }
        for (int q = lane; q < n_q; q += 32) {
            double wd = s_w[q] * g.det;
            s_c0[q] = wd * c0.p[cell * c0.sc + q * c0.sq];
            if (FORM == 0) s_c1[q] = wd * c1.p[cell * c1.sc + q * c1.sq];
        }
        if (alpha)
            for (int i = lane; i < n_loc; i += 32) s_a[i] = alpha[dest[cell * n_loc + i]];
        __syncwarp();

        double *K = Ke + cell * (int64_t)n_loc * n_loc;
        for (int e = lane; e < n_loc * n_loc; e += 32) {
            int ra = e / n_loc, rb = e - ra * n_loc;
            int a = ra / NC, i = ra - a * NC, b = rb / NC, j = rb - b * NC;
            double acc = 0.0;
            if (FORM == 0) {
                for (int q = 0; q < n_q; ++q) {
                    const double *ga = s_g + (q * n_phi + a) * DIM;
                    const double *gb = s_g + (q * n_phi + b) * DIM;
                    double t = s_c0[q] * ga[i] * gb[j] + s_c1[q] * ga[j] * gb[i];
                    if (i == j) {
                        double dot = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; ++d) dot += ga[d] * gb[d];
                        t += s_c1[q] * dot;
                    }
                    acc += t;
                }
            } else {
                for (int q = 0; q < n_q; ++q) {
                    const double *ga = s_g + (q * n_phi + a) * DIM;
                    const double *gb = s_g + (q * n_phi + b) * DIM;
                    double dot = 0.0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d) dot += ga[d] * gb[d];
                    acc += s_c0[q] * dot;
                }
            }
            K[e] = acc;
        }
        __syncwarp();
        // residual: load (+ for elasticity, - for Laplace) + K alpha
        for (int ra = lane; ra < n_loc; ra += 32) {
            int a = ra / NC, i = ra - a * NC;
            double acc = 0.0;
            if (f) {
                for (int q = 0; q < n_q; ++q)
                    acc += s_w[q] * g.det * s_tab[q * n_phi + a] * f[(cell * n_q + q) * NC + i];
                if (FORM == 1) acc = -acc;
            }
            if (alpha) {
                double ka = 0.0;
                for (int rb = 0; rb < n_loc; ++rb) ka += K[ra * n_loc + rb] * s_a[rb];
                acc += ka;
            }
            re[cell * n_loc + ra] = acc;
        }
        __syncwarp();
    }
}

template <int DIM, int FORM>
static int launch_local_h1(int n_phi, int n_q, const double *tab, const double *weights, const double *coords,
                           const int32_t *cell_nodes, int64_t n_cells, Coef c0, Coef c1, const double *f,
                           const double *alpha, const int32_t *dest, double *Ke, double *re, cudaStream_t s) {
    constexpr int NC = (FORM == 0) ? DIM : 1;
    int n_loc = n_phi * NC;
    size_t smem = sizeof(double) * ((size_t)(1 + DIM) * n_q * n_phi + n_q +
                                    (size_t)WARPS_PER_CTA * (n_q * n_phi * DIM + 2 * n_q + n_loc));
    if (smem > 200 * 1024) return fail("nfa_local_h1: element too large for the shared-memory staging");
    auto kern = local_h1_kernel<DIM, FORM>;
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ctas = (n_cells + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    kern<<<grid, WARPS_PER_CTA * 32, smem, s>>>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f,
                                                alpha, dest, Ke, re);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_h1_elasticity(int dim, int n_phi, int n_q, const double *tab, const double *weights,
                                       const double *coords, const int32_t *cell_nodes, int64_t n_cells,
                                       const double *lam, int64_t lsc, int64_t lsq, const double *mu, int64_t msc,
                                       int64_t msq, const double *f, const double *alpha, const int32_t *dest,
                                       double *Ke, double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (alpha && !dest) return fail("nfa_local_h1_elasticity: alpha given without dest");
    Coef c0{lam, lsc, lsq}, c1{mu, msc, msq};
    cudaStream_t s = (cudaStream_t)stream;
    if (dim == 3)
        return launch_local_h1<3, 0>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f, alpha, dest,
                                     Ke, re, s);
    if (dim == 2)
        return launch_local_h1<2, 0>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f, alpha, dest,
                                     Ke, re, s);
    return fail("nfa_local_h1_elasticity: dim must be 2 or 3");
}

extern "C" int nfa_local_h1_laplace(int dim, int n_phi, int n_q, const double *tab, const double *weights,
                                    const double *coords, const int32_t *cell_nodes, int64_t n_cells,
                                    const double *kappa, int64_t ksc, int64_t ksq, const double *f,
                                    const double *alpha, const int32_t *dest, double *Ke, double *re,
                                    nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (alpha && !dest) return fail("nfa_local_h1_laplace: alpha given without dest");
    Coef c0{kappa, ksc, ksq}, c1{nullptr, 0, 0};
    cudaStream_t s = (cudaStream_t)stream;
    if (dim == 3)
        return launch_local_h1<3, 1>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f, alpha, dest,
                                     Ke, re, s);
    if (dim == 2)
        return launch_local_h1<2, 1>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f, alpha, dest,
                                     Ke, re, s);
    if (dim == 1)
        return launch_local_h1<1, 1>(n_phi, n_q, tab, weights, coords, cell_nodes, n_cells, c0, c1, f, alpha, dest,
                                     Ke, re, s);
    return fail("nfa_local_h1_laplace: dim must be 1, 2 or 3");
}

// ------------------------------------------------------------------------------------------------
// K5: facet mass / load.  One warp per facet.
template <int FDIM>
__global__ void __launch_bounds__(128)
    local_h1_facet_kernel(int n_phi, int n_q, int n_comp, const double *__restrict__ tab0,
                          const double *__restrict__ weights, const double *__restrict__ coords,
                          const int32_t *__restrict__ facet_nodes, int64_t n_facets,
                          const double *__restrict__ gval, const double *__restrict__ comp_scale, double beta,
                          double r_sign, int with_matrix, double *__restrict__ Ke, double *__restrict__ re) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_loc = n_phi * n_comp;
    for (int64_t fc = (int64_t)blockIdx.x * 4 + warp; fc < n_facets; fc += (int64_t)gridDim.x * 4) {
        double det = 1.0;
        if constexpr (FDIM > 0) {
            double X[FDIM + 1][3];
            load_cell_coords<FDIM>(coords, facet_nodes, fc, X);
            Geom<FDIM> g;
            compute_geom<FDIM>(X, g);
            det = g.det;
        }
        if (Ke) {
            double *K = Ke + fc * (int64_t)n_loc * n_loc;
            for (int e = lane; e < n_loc * n_loc; e += 32) {
                int ra = e / n_loc, rb = e - ra * n_loc;
                int a = ra / n_comp, i = ra - a * n_comp, b = rb / n_comp, j = rb - b * n_comp;
                double acc = 0.0;
                if (with_matrix && i == j) {
                    for (int q = 0; q < n_q; ++q)
                        acc += beta * det * weights[q] * (tab0[q * n_phi + a] * tab0[q * n_phi + b]);
                    if (comp_scale) acc *= comp_scale[fc * n_comp + i];
                }
                K[e] = acc;
            }
        }
        if (re) {
            for (int ra = lane; ra < n_loc; ra += 32) {
                int a = ra / n_comp, i = ra - a * n_comp;
                double acc = 0.0;
                if (gval)
                    for (int q = 0; q < n_q; ++q)
                        acc += beta * det * weights[q] * gval[(fc * n_q + q) * n_comp + i] * tab0[q * n_phi + a];
                re[fc * n_loc + ra] = r_sign * acc;
            }
        }
    }
}

extern "C" int nfa_local_h1_facet(int facet_dim, int n_phi, int n_q, int n_comp, const double *tab0,
                                  const double *weights, const double *coords, const int32_t *facet_nodes,
                                  int64_t n_facets, const double *gval, const double *comp_scale, double beta,
                                  double r_sign, int with_matrix, double *Ke, double *re, nfa_stream stream) {
    if (n_facets == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t ctas = (n_facets + 3) / 4;
    int grid = (int)(ctas < 148 * 16 ? ctas : 148 * 16);
    if (facet_dim == 3)  // a volume cell: the same kernel is the Lagrange mass matrix / load of L2ProjectorWeakForm
        local_h1_facet_kernel<3><<<grid, 128, 0, s>>>(n_phi, n_q, n_comp, tab0, weights, coords, facet_nodes,
                                                      n_facets, gval, comp_scale, beta, r_sign, with_matrix, Ke, re);
    else if (facet_dim == 2)
        local_h1_facet_kernel<2><<<grid, 128, 0, s>>>(n_phi, n_q, n_comp, tab0, weights, coords, facet_nodes,
                                                      n_facets, gval, comp_scale, beta, r_sign, with_matrix, Ke, re);
    else if (facet_dim == 1)
        local_h1_facet_kernel<1><<<grid, 128, 0, s>>>(n_phi, n_q, n_comp, tab0, weights, coords, facet_nodes,
                                                      n_facets, gval, comp_scale, beta, r_sign, with_matrix, Ke, re);
    else if (facet_dim == 0)
        local_h1_facet_kernel<0><<<grid, 128, 0, s>>>(n_phi, n_q, n_comp, tab0, weights, coords, facet_nodes,
                                                      n_facets, gval, comp_scale, beta, r_sign, with_matrix, Ke, re);
    else
        return fail("nfa_local_h1_facet: facet_dim must be 0..3");
    NFA_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Load vector only (fused path): re[cell][a*nc + i] = sign * sum_q w_q det phi_a(q) f[cell][q][i].
// A CTA takes tiles of TILE consecutive cells: their f blocks are one contiguous span of TILE * n_q * n_comp doubles,
// copied to shared memory by all threads at once (many loads in flight -- the per-warp version was latency bound,
// profiles/r01_load_kernel_ncu_summary.txt); thread (cell, a) then contracts its cell's block with column a of the
// premultiplied table w_q phi_a(q) for all components and writes n_comp consecutive outputs.  HBM-bound on f.
template <int DIM>
__global__ void __launch_bounds__(1024, 1)
    local_h1_load_kernel(int n_phi, int n_q, int n_comp, int tile, const double *__restrict__ tab0,
                         const double *__restrict__ weights, const double *__restrict__ coords,
                         const int32_t *__restrict__ cell_nodes, int64_t n_cells,
                         const double *__restrict__ f, double sign, double *__restrict__ re) {
    extern __shared__ double smem[];
    double *s_wt = smem;                 // [n_q][n_phi]  w_q phi_a(q)
    double *s_f = s_wt + n_q * n_phi;    // [tile][n_q][n_comp]
    const int fn = n_q * n_comp, n_loc = n_phi * n_comp;
    for (int i = threadIdx.x; i < n_q * n_phi; i += blockDim.x) s_wt[i] = weights[i / n_phi] * tab0[i];
    const int lc = threadIdx.x / n_phi, a = threadIdx.x - lc * n_phi;  // worker = (local cell, basis function)
    const int64_t n_tiles = (n_cells + tile - 1) / tile;
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const int64_t c0 = tl * tile;
        const int nc_here = (int)((n_cells - c0) < tile ? (n_cells - c0) : tile);
        __syncthreads();  // previous tile's readers are done (also orders the s_wt fill on the first trip)
        if (f) {
            const double *src = f + c0 * fn;
            for (int t = threadIdx.x; t < nc_here * fn; t += blockDim.x) s_f[t] = __ldg(src + t);
        }
        double det = 0.0;
        if (lc < nc_here) {
            double X[DIM + 1][3];
            load_cell_coords<DIM>(coords, cell_nodes, c0 + lc, X);
            det = compute_det<DIM>(X);
        }
        __syncthreads();
        if (lc < nc_here) {
            double acc[3] = {0.0, 0.0, 0.0};
            if (f) {
                const double *fc = s_f + lc * fn;
                for (int q = 0; q < n_q; ++q) {
                    const double wv = s_wt[q * n_phi + a];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < n_comp) acc[i] += wv * fc[q * n_comp + i];
                }
            }
            double *dst = re + (c0 + lc) * n_loc + a * n_comp;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < n_comp) dst[i] = sign * (acc[i] * det);
        }
    }
}

// Streaming variant for the common element sizes: the premultiplied table w_q phi_a(q) sits in constant memory (DFMA
// reads it as an operand: no load instruction), a thread owns (cell, component), reads its N_Q load values from the
// shared-memory tile and keeps the N_PHI sums in registers.  Tiles of TILE cells (one contiguous span of f) arrive by
// bulk asynchronous copies (cp.async.bulk global -> shared, mbarrier completion), double buffered.  HBM-bound on f.
__constant__ double c_load_wt[64 * 20];

// MERGE (nfa_bplan_load_rhs): the tile's local load vectors stay in shared memory and are summed per NODE of the tile
// (the plan's (tile, node) pairs, entries in ascending cell order); only the pair sums [n_pairs][N_COMP] go to HBM --
// about 3.2 node records per cell instead of 10 -- and a second kernel adds a node's pair sums in ascending tile order.
constexpr int LOAD_SPLIT = 1;  // measured: 2 (five sums per thread, 24 warps per SM) is slower, 266 us against 204
struct LoadMerge {
    const int32_t *tile_ptr;   // [n_tiles + 1]
    const uint32_t *pair_ptr;  // [n_pairs + 1]
    const uint16_t *tile_ent;  // [n_cells * N_PHI]
    double *pair_sum;          // [n_pairs][N_COMP]
};

// sums a = A0 .. A0 + NA of one (cell, component): sc * sum_q w_q phi_a(q) f_q, written to dst[a * pitch]
template <int N_PHI, int NA, int N_COMP, int A0>
__device__ __forceinline__ void load_sums(const double *fc, int n_q, double sc, double *dst, int pitch) {
    double acc[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[a] = 0.0;
    for (int q = 0; q < n_q; ++q) {
        const double v = fc[q * N_COMP];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = fma(c_load_wt[q * N_PHI + A0 + a], v, acc[a]);
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) dst[(A0 + a) * pitch] = sc * acc[a];
}

template <int DIM, int N_PHI, int N_COMP, int TILE, bool MERGE>
__global__ void __launch_bounds__(TILE *N_COMP *LOAD_SPLIT)
    local_h1_load_stream_kernel(int n_q, const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes,
                                int64_t n_cells, const double *__restrict__ f, double sign, double *__restrict__ re,
                                int comp_pitch, LoadMerge mg) {
    extern __shared__ __align__(128) double smem[];  // two tiles [TILE][n_q][N_COMP], two mbarriers, (MERGE) re tile + lists
    constexpr int NT = TILE * N_COMP * LOAD_SPLIT;  // LOAD_SPLIT threads share a (cell, component): N_PHI / LOAD_SPLIT sums each
    constexpr int NA = N_PHI / LOAD_SPLIT;
    static_assert(N_PHI % LOAD_SPLIT == 0, "basis functions must split evenly");
    constexpr int LIST_WORDS = (TILE * N_PHI + 2) / 2 + TILE * N_PHI / 2;  // 32-bit words of one tile's lists
    const int fn = n_q * N_COMP;
    const int tile_doubles = TILE * fn;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * (size_t)tile_doubles);
    double *re_tile = smem + 2 * (size_t)tile_doubles + 2;  // MERGE: [TILE][N_PHI][N_COMP]
    // MERGE: the lists of two tiles (current / next) -- pair offsets relative to the tile's first entry (uint16
    // [TILE * N_PHI + 2]), then the entries' local ids (uint16, stored two a word)
    uint32_t *lists = reinterpret_cast<uint32_t *>(re_tile + TILE * N_PHI * N_COMP);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bar);
    const int part = threadIdx.x / (TILE * N_COMP), within = threadIdx.x - part * (TILE * N_COMP);
    const int lc = within / N_COMP, comp = within - lc * N_COMP;
    const int64_t n_tiles = (n_cells + TILE - 1) / TILE;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t tl, int buf) {  // thread 0 only
        const int64_t c0 = tl * TILE;
        const int nc_here = (int)((n_cells - c0) < TILE ? (n_cells - c0) : TILE);
        const int n_dbl = nc_here * fn;
        const uint32_t bytes = (uint32_t)((n_dbl & ~1) * 8);  // bulk copies move multiples of 16 bytes
        double *tile = smem + (size_t)buf * tile_doubles;
        if (n_dbl & 1) tile[n_dbl - 1] = __ldg(f + c0 * fn + n_dbl - 1);  // odd tail (last tile only), see the loop
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * buf), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(f + c0 * fn), "r"(bytes), "r"(bar0 + 8 * buf)
            : "memory");
    };
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    // Geometry pipeline: the vertex ids of a tile are requested two iterations ahead, its coordinates one iteration ahead,
    // and the determinant is formed at the end of the previous iteration -- the two dependent global round trips
    // (cell_nodes -> coords) never stall the iteration that needs them.
    int32_t ids_next[DIM + 1];  // vertex ids of this thread's cell in the next tile
    double det = 0.0;           // determinant of this thread's cell in the current tile
    auto load_ids = [&](int64_t tile, int32_t (&ids)[DIM + 1]) {
        const int64_t cell = tile * TILE + lc;
#pragma unroll
        for (int v = 0; v <= DIM; ++v) ids[v] = (tile < n_tiles && cell < n_cells) ? __ldg(cell_nodes + cell * (DIM + 1) + v) : 0;
    };
    auto load_X = [&](const int32_t (&ids)[DIM + 1], double (&X)[DIM + 1][3]) {
#pragma unroll
        for (int v = 0; v <= DIM; ++v)
#pragma unroll
            for (int r = 0; r < 3; ++r) X[v][r] = __ldg(coords + 3 * (int64_t)ids[v] + r);
    };
    // List pipeline (MERGE), same idea: pair range two tiles ahead, list words one tile ahead (registers), stored into the
    // other list buffer at the end of the iteration.
    constexpr int NP_IT = MERGE ? (TILE * N_PHI + 1 + NT - 1) / NT : 1, NE_IT = MERGE ? (TILE * N_PHI / 2 + NT - 1) / NT : 1;
    uint32_t lp[NP_IT], le[NE_IT];
    int pc0 = 0, pc1 = 0, pn0 = 0, pn1 = 0;  // pair ranges of the current / next tile
    auto load_range = [&](int64_t tile, int &a, int &b) {
        a = b = 0;
        if (tile < n_tiles) {
            a = __ldg(mg.tile_ptr + tile);
            b = __ldg(mg.tile_ptr + tile + 1);
        }
    };
    auto load_lists = [&](int64_t tile, int q0, int q1) {  // into registers
        const int64_t c0 = tile * TILE;
        const int nc_t = (tile < n_tiles) ? (int)((n_cells - c0) < TILE ? (n_cells - c0) : TILE) : 0;
#pragma unroll
        for (int j = 0; j < NP_IT; ++j) {
            const int i = threadIdx.x + j * NT;
            lp[j] = (nc_t > 0 && i <= q1 - q0) ? __ldg(mg.pair_ptr + q0 + i) : 0u;  // (no arithmetic here: that would wait)
        }
        const uint32_t *ent32 = reinterpret_cast<const uint32_t *>(mg.tile_ent + (size_t)c0 * N_PHI);  // even offset
#pragma unroll
        for (int j = 0; j < NE_IT; ++j) {
            const int i = threadIdx.x + j * NT;
            le[j] = (i < (nc_t * N_PHI) / 2) ? __ldg(ent32 + i) : 0u;
        }
    };
    auto store_lists = [&](int lbuf, int64_t tile) {  // registers -> list buffer lbuf (offsets relative to the tile's entries)
        const uint32_t e_base = (uint32_t)(tile * TILE * N_PHI);
        uint16_t *sp = reinterpret_cast<uint16_t *>(lists + lbuf * LIST_WORDS);
        uint32_t *se = lists + lbuf * LIST_WORDS + (TILE * N_PHI + 2) / 2;
#pragma unroll
        for (int j = 0; j < NP_IT; ++j) {
            const int i = threadIdx.x + j * NT;
            if (i <= TILE * N_PHI) sp[i] = (uint16_t)(lp[j] - e_base);
        }
#pragma unroll
        for (int j = 0; j < NE_IT; ++j) {
            const int i = threadIdx.x + j * NT;
            if (i < TILE * N_PHI / 2) se[i] = le[j];
        }
    };
    {
        int32_t ids0[DIM + 1];
        double X0[DIM + 1][3];
        load_ids(blockIdx.x, ids0);
        load_ids((int64_t)blockIdx.x + gridDim.x, ids_next);
        if constexpr (MERGE) {
            load_range(blockIdx.x, pc0, pc1);
            load_range((int64_t)blockIdx.x + gridDim.x, pn0, pn1);
            load_lists(blockIdx.x, pc0, pc1);
        }
        load_X(ids0, X0);
        det = compute_det<DIM>(X0);
        if constexpr (MERGE) store_lists(0, blockIdx.x);  // visible to the node phase after the barrier that precedes it
    }
    int buf = 0;
    uint32_t phases = 0u;  // bit b = parity the next wait on tile b expects
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, buf ^= 1) {
        const int64_t next = tl + gridDim.x;
        if (threadIdx.x == 0 && next < n_tiles) issue(next, buf ^ 1);  // the other tile was released by the barrier below
        const int64_t c0 = tl * TILE;
        const int nc_here = (int)((n_cells - c0) < TILE ? (n_cells - c0) : TILE);
        double X_next[DIM + 1][3];  // coordinates of the next tile's cell (its ids arrived during the previous iteration)
        load_X(ids_next, X_next);
        load_ids(next + gridDim.x, ids_next);
        int pnn0 = 0, pnn1 = 0;
        if constexpr (MERGE) {
            load_lists(next, pn0, pn1);
            load_range(next + gridDim.x, pnn0, pnn1);
        }
        {  // wait for this tile's bytes
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar0 + 8 * buf), "r"((phases >> buf) & 1u)
                    : "memory");
            }
            phases ^= 1u << buf;
        }
        if ((nc_here * fn) & 1) __syncthreads();  // the odd tail element was written by thread 0 with a plain store
        if (lc < nc_here) {
            const double *fc = smem + (size_t)buf * tile_doubles + lc * fn + comp;
            // the branch on `part` is warp-uniform (TILE * N_COMP is a multiple of 32) and makes the table offsets
            // compile-time constants: the weights stay uniform constant-bank operands
            double *dst = MERGE ? re_tile + lc * (N_PHI * N_COMP) + comp : re + (c0 + lc) * (N_PHI * comp_pitch) + comp;
            const int pitch = MERGE ? N_COMP : comp_pitch;  // comp_pitch > N_COMP: padded node records
            const double sc = sign * det;
            if (part == 0)
                load_sums<N_PHI, NA, N_COMP, 0>(fc, n_q, sc, dst, pitch);
            else
                load_sums<N_PHI, NA, N_COMP, NA>(fc, n_q, sc, dst, pitch);
        }
        if constexpr (MERGE) {
            __syncthreads();  // the tile's local load vectors (and its lists, stored an iteration ago) are complete
            const uint16_t *s_ptr = reinterpret_cast<const uint16_t *>(lists + buf * LIST_WORDS);
            const uint16_t *ent = reinterpret_cast<const uint16_t *>(lists + buf * LIST_WORDS + (TILE * N_PHI + 2) / 2);
            const int n_pairs_here = pc1 - pc0;
            for (int w = threadIdx.x; w < n_pairs_here * N_COMP; w += NT) {
                const int k = w / N_COMP, c = w - k * N_COMP;
                const int e0 = s_ptr[k], e1 = s_ptr[k + 1];
                double sum = 0.0;
                for (int e = e0; e < e1; e += 4) {  // four independent loads in flight, added in ascending cell order
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (e + u < e1) ? re_tile[(int)ent[e + u] * N_COMP + c] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (e + u < e1) sum += v[u];
                }
                mg.pair_sum[(size_t)(pc0 + k) * N_COMP + c] = sum;
            }
            store_lists(buf ^ 1, next);  // the next tile's lists (that buffer was last read two barriers ago)
            pc0 = pn0, pc1 = pn1, pn0 = pnn0, pn1 = pnn1;
        }
        det = compute_det<DIM>(X_next);
        __syncthreads();  // every reader of this tile is done: it may be refilled on the next trip
    }
}

__global__ void load_wt_kernel(int n, int n_phi, const double *__restrict__ tab0, const double *__restrict__ weights,
                               double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = weights[i / n_phi] * tab0[i];
}

template <int DIM, int N_PHI, int N_COMP, int TILE = 64, bool MERGE = false>
static int launch_h1_load_stream(int n_q, const double *tab0, const double *weights, const double *coords,
                                 const int32_t *cell_nodes, int64_t n_cells, const double *f, double sign, double *re,
                                 cudaStream_t s, int comp_pitch = N_COMP, LoadMerge mg = LoadMerge{}) {
    double *wt = nullptr;
    NFA_CUDA(cudaGetSymbolAddress((void **)&wt, c_load_wt));
    load_wt_kernel<<<grid_for(n_q * N_PHI, 128), 128, 0, s>>>(n_q * N_PHI, N_PHI, tab0, weights, wt);
    NFA_LAUNCH_CHECK();
    const size_t smem = sizeof(double) * 2 * (size_t)TILE * n_q * N_COMP + 16 +
                        (MERGE ? sizeof(double) * (size_t)TILE * N_PHI * N_COMP + 2 * 4 * ((TILE * N_PHI + 2) / 2 + TILE * N_PHI / 2) : 0);
    auto kern = local_h1_load_stream_kernel<DIM, N_PHI, N_COMP, TILE, MERGE>;
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE * N_COMP * LOAD_SPLIT, smem));
    if (per_sm < 1) return fail("nfa_local_h1_load: streaming kernel does not fit on an SM");
    const int64_t tiles = (n_cells + TILE - 1) / TILE;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)n_sm * per_sm);
    kern<<<grid, TILE * N_COMP * LOAD_SPLIT, smem, s>>>(n_q, coords, cell_nodes, n_cells, f, sign, re, comp_pitch, mg);
    NFA_LAUNCH_CHECK();
    return 0;
}

// second pass of the tile-merged load: row N_COMP N + c = sum of node N's pair sums in ascending tile order
template <int NC>
__global__ void __launch_bounds__(256)
    pair_gather_kernel(int64_t n_active, const int32_t *__restrict__ active, const int32_t *__restrict__ node_pair_ptr,
                       const uint32_t *__restrict__ node_pair, const double *__restrict__ pair_sum, int overwrite,
                       double *__restrict__ rhs) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_active) return;
    const int64_t N = active ? active[t] : t;  // (the list is the identity when every node has cells)
    const int32_t b = node_pair_ptr[N], e = node_pair_ptr[N + 1];
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (int32_t k = b; k < e; ++k) {
        const double *p = pair_sum + (size_t)__ldg(node_pair + k) * NC;
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += __ldg(p + c);
    }
    double *out = rhs + (size_t)N * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) out[c] = overwrite ? acc[c] : out[c] + acc[c];
}

extern "C" int nfa_bplan_load_rhs(const nfa_bplan *bp, int dim, int n_q, const double *tab0, const double *weights,
                                  const double *coords, const int32_t *cell_nodes, const double *f, double sign,
                                  int overwrite, double *rhs, nfa_stream stream) {
    if (!bp || !tab0 || !weights || !coords || !cell_nodes || !f || !rhs) return fail("nfa_bplan_load_rhs: NULL argument");
    if (n_q < 1 || n_q > 64) return fail("nfa_bplan_load_rhs: 1 to 64 quadrature points");
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int LT = nfa_bplan::LOAD_TILE;  // cells per tile of the plan's pair lists (64; NFA_LOAD_TILE for experiments)
    static_assert(LT == 32 || LT == 64 || LT == 128, "the streaming load kernel runs tiles of 32, 64 or 128 cells");
    const LoadMerge mg = {(const int32_t *)bp->tile_ptr.p, (const uint32_t *)bp->pair_ptr.p, (const uint16_t *)bp->tile_ent.p,
                          (double *)bp->pair_sum.p};
    int rc;
    if (dim == 3 && bp->n_phi == 10 && bp->n_comp == 3)
        rc = launch_h1_load_stream<3, 10, 3, LT, true>(n_q, tab0, weights, coords, cell_nodes, bp->n_cells, f, sign, nullptr, s, 3, mg);
    else if (dim == 3 && bp->n_phi == 4 && bp->n_comp == 3)
        rc = launch_h1_load_stream<3, 4, 3, LT, true>(n_q, tab0, weights, coords, cell_nodes, bp->n_cells, f, sign, nullptr, s, 3, mg);
    else if (dim == 2 && bp->n_phi == 6 && bp->n_comp == 2)
        rc = launch_h1_load_stream<2, 6, 2, LT, true>(n_q, tab0, weights, coords, cell_nodes, bp->n_cells, f, sign, nullptr, s, 2, mg);
    else
        return fail("nfa_bplan_load_rhs: vector P2 / P1 on tetrahedra and vector P2 on triangles only");
    if (rc) return rc;
    // only the nodes that have cells are visited: a rank's local plan (global numbering) leaves the other ranks' rows alone
    const unsigned grid = (unsigned)grid_for(bp->n_active, 256);
    const int32_t *act = bp->n_active == bp->n_nodes ? nullptr : (const int32_t *)bp->active_nodes.p;
    const int32_t *npp = (const int32_t *)bp->node_pair_ptr.p;
    const uint32_t *np = (const uint32_t *)bp->node_pair.p;
    if (bp->n_comp == 3)
        pair_gather_kernel<3><<<grid, 256, 0, s>>>(bp->n_active, act, npp, np, (const double *)bp->pair_sum.p, overwrite, rhs);
    else
        pair_gather_kernel<2><<<grid, 256, 0, s>>>(bp->n_active, act, npp, np, (const double *)bp->pair_sum.p, overwrite, rhs);
    NFA_LAUNCH_CHECK();
    return 0;
}

template <int DIM>
static int launch_h1_load(int n_phi, int n_q, int n_comp, const double *tab0, const double *weights,
                          const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *f, double sign,
                          double *re, cudaStream_t s) {
    if (n_comp > 3) return fail("nfa_local_h1_load: at most 3 components");
    int tile = 1024 / n_phi < 32 ? 1024 / n_phi : 32;
    if (tile < 1) return fail("nfa_local_h1_load: element too large");
    int threads = tile * n_phi;
    size_t smem = sizeof(double) * ((size_t)n_q * n_phi + (size_t)tile * n_q * n_comp);
    if (smem > 200 * 1024) return fail("nfa_local_h1_load: element too large for the shared-memory tile");
    auto kern = local_h1_load_kernel<DIM>;
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t tiles = (n_cells + tile - 1) / tile;
    int grid = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
    kern<<<grid, threads, smem, s>>>(n_phi, n_q, n_comp, tile, tab0, weights, coords, cell_nodes, n_cells, f, sign, re);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_local_h1_load_pitched(int dim, int n_phi, int n_q, int n_comp, const double *tab0, const double *weights,
                                         const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *f,
                                         double sign, double *re, int comp_pitch, nfa_stream stream) {
    if (n_cells == 0) return 0;
    if (!f) return fail("nfa_local_h1_load_pitched: the load array is NULL");
    if (!(dim == 3 && n_phi == 10 && n_comp == 3 && n_q <= 64 && comp_pitch >= 3))
        return fail("nfa_local_h1_load_pitched: vector P2 on tetrahedra with at most 64 quadrature points only");
    return launch_h1_load_stream<3, 10, 3>(n_q, tab0, weights, coords, cell_nodes, n_cells, f, sign, re,
                                           (cudaStream_t)stream, comp_pitch);
}

extern "C" int nfa_local_h1_load(int dim, int n_phi, int n_q, int n_comp, const double *tab0, const double *weights,
                                 const double *coords, const int32_t *cell_nodes, int64_t n_cells, const double *f,
                                 double sign, double *re, nfa_stream stream) {
    if (n_cells == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    // streaming kernel: tiles of 64 cells of f start 16-byte aligned whatever n_q * n_comp is
    if (f && n_q <= 64) {
        if (dim == 3 && n_phi == 10 && n_comp == 3)
            return launch_h1_load_stream<3, 10, 3>(n_q, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
        if (dim == 3 && n_phi == 4 && n_comp == 3)
            return launch_h1_load_stream<3, 4, 3>(n_q, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
        if (dim == 2 && n_phi == 6 && n_comp == 2)
            return launch_h1_load_stream<2, 6, 2>(n_q, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
    }
    if (dim == 3) return launch_h1_load<3>(n_phi, n_q, n_comp, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
    if (dim == 2) return launch_h1_load<2>(n_phi, n_q, n_comp, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
    if (dim == 1) return launch_h1_load<1>(n_phi, n_q, n_comp, tab0, weights, coords, cell_nodes, n_cells, f, sign, re, s);
    return fail("nfa_local_h1_load: dim must be 1, 2 or 3");
}
