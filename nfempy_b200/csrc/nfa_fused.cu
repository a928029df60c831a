// libnfa: fused H1 assembly for affine simplices with cell-wise constant coefficients (sm_100a).
//
// One kernel does the local B^T D B contraction AND the global deterministic reduce: no local matrices are
// written to HBM.  Work is organised by OUTPUT: a CTA owns a chunk of whole node-block rows of the CSR, lists
// every (cell, a, b) contribution that lands in its rows, computes each 3x3 (or 1x1) block from the cell's
// 96-byte geometry record and the exact reference tensor Mref held in shared memory, stages the blocks in
// shared memory and sums them per CSR value in ascending cell order -- the order in which the reference's
// Mat.setValue(..., addv=True) calls arrive (R/src/assembly/ScatterFormData.py:5-11).  No atomics.
//
// Per contribution the local entry is (elasticity, R/src/weak_forms/le_primal_weak_form.py:91-126 integrated
// exactly):  T = Jinv^T Mref[a][b] Jinv,  K = (det lambda) T + (det mu) T^T + (det mu) tr(T) I.
#include <algorithm>
#include <cub/cub.cuh>
#include <vector>

#include "nfa_plan.cuh"

using namespace nfa;

namespace {

#ifndef NFA_FUSED_THREADS
#define NFA_FUSED_THREADS 256
#endif
#ifndef NFA_FUSED_MINBLOCKS
#define NFA_FUSED_MINBLOCKS 4
#endif
constexpr int FUSED_THREADS = NFA_FUSED_THREADS;

__global__ void fill_block_keys(const int32_t *__restrict__ node_dest, int64_t n_cells, int n_phi, int64_t n_nodes,
                                int pair_bits, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int p2 = n_phi * n_phi;
    if (e >= n_cells * p2) return;
    int64_t cell = e / p2;
    int r = (int)(e - cell * p2);
    int a = r / n_phi, b = r - a * n_phi;
    uint64_t na = (uint64_t)node_dest[cell * n_phi + a], nb = (uint64_t)node_dest[cell * n_phi + b];
    keys[e] = na * (uint64_t)n_nodes + nb;
    ids[e] = ((uint32_t)cell << pair_bits) | (uint32_t)r;  // ascending id == ascending (cell, a, b)
}

__global__ void block_cols(const uint64_t *__restrict__ ukeys, int64_t n, int64_t n_nodes, int32_t *__restrict__ bcols) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) bcols[u] = (int32_t)(ukeys[u] % (uint64_t)n_nodes);
}

__global__ void block_ptr(const uint64_t *__restrict__ ukeys, int64_t n, int64_t n_nodes, int32_t *__restrict__ bptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_nodes) return;
    uint64_t target = (uint64_t)r * (uint64_t)n_nodes;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (ukeys[mid] < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    bptr[r] = (int32_t)lo;
}

// per block row: number of contributions
__global__ void row_contribs(const int32_t *__restrict__ bptr, const uint32_t *__restrict__ bseg, int64_t n_nodes,
                             uint32_t *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_nodes) out[r] = bseg[bptr[r + 1]] - bseg[bptr[r]];
}

// key = (chunk << 32) | contribution id, value = staging slot within the chunk
__global__ void chunk_keys(const uint32_t *__restrict__ ids_sorted, int64_t n_contrib,
                           const uint32_t *__restrict__ chunk_cstart, int64_t n_chunks,
                           uint64_t *__restrict__ keys, uint32_t *__restrict__ slots) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_contrib) return;
    int64_t lo = 0, hi = n_chunks;  // last chunk with cstart <= k
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (chunk_cstart[mid] <= (uint32_t)k)
            lo = mid;
        else
            hi = mid;
    }
    keys[k] = ((uint64_t)lo << 32) | (uint64_t)ids_sorted[k];
    slots[k] = (uint32_t)(k - chunk_cstart[lo]);
}

// k = position in cell order; slots[k] = rank in output order (both relative to the chunk start):
// contrib[k] = id, gather_pos[chunk start + output rank] = position in cell order
__global__ void unpack_chunk_sorted(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ slots, int64_t n,
                                    const uint32_t *__restrict__ chunk_cstart, uint32_t *__restrict__ contrib,
                                    uint16_t *__restrict__ gather_pos) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    contrib[k] = (uint32_t)(keys[k] & 0xffffffffull);
    uint32_t cs = chunk_cstart[keys[k] >> 32];
    gather_pos[cs + slots[k]] = (uint16_t)(k - cs);
}

// node incidence lists: key = node of (cell, a), id = cell * n_phi + a (ascending id = ascending cell)
__global__ void fill_node_keys(const int32_t *__restrict__ node_dest, int64_t n, uint32_t *__restrict__ keys,
                               uint32_t *__restrict__ ids) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    keys[e] = (uint32_t)node_dest[e];
    ids[e] = (uint32_t)e;
}

// tile-merged load lists: key = tile * n_nodes + node of (cell, a), id = cell * n_phi + a
__global__ void fill_tile_node_keys(const int32_t *__restrict__ node_dest, int64_t n, int n_phi, int tile_cells,
                                    int64_t n_nodes, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int64_t tile = (e / n_phi) / tile_cells;
    keys[e] = (uint64_t)tile * (uint64_t)n_nodes + (uint64_t)node_dest[e];
    ids[e] = (uint32_t)e;
}

__global__ void tile_local_ids(const uint32_t *__restrict__ ids, int64_t n, int n_phi, int tile_cells,
                               uint16_t *__restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t id = ids[e];
    const uint32_t tile = (id / (uint32_t)n_phi) / (uint32_t)tile_cells;
    out[e] = (uint16_t)(id - tile * (uint32_t)(tile_cells * n_phi));
}

// first unique key >= r * stride (r = 0 .. n_rows): pair range of every tile
__global__ void lower_bound_u64(const uint64_t *__restrict__ ukeys, int64_t n, int64_t n_rows, uint64_t stride,
                                int32_t *__restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    const uint64_t target = (uint64_t)r * stride;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (ukeys[mid] < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    out[r] = (int32_t)lo;
}

__global__ void pair_node_keys(const uint64_t *__restrict__ ukeys, int64_t n, int64_t n_nodes, uint32_t *__restrict__ keys,
                               uint32_t *__restrict__ ids) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    keys[p] = (uint32_t)(ukeys[p] % (uint64_t)n_nodes);
    ids[p] = (uint32_t)p;
}

__global__ void nonempty_flags(const int32_t *__restrict__ ptr, int64_t n, uint8_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = ptr[i + 1] > ptr[i] ? 1 : 0;
}

__global__ void node_ptr_kernel(const uint32_t *__restrict__ sorted_keys, int64_t n, int64_t n_nodes,
                                int32_t *__restrict__ node_ptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_nodes) return;
    int64_t lo = 0, hi = n;  // first entry whose node is >= r
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (uint32_t)r)
            lo = mid + 1;
        else
            hi = mid;
    }
    node_ptr[r] = (int32_t)lo;
}

// Right-hand-side rows of node N = the sum of the local load vectors of the (cell, local node) pairs that are N, in
// ascending cell order (res_g[dest] += r_el, R/src/assembly/ScatterFormData.py:14-18).  One thread per node: its entries
// are consecutive in node_ent, every entry is one record re[id * PITCH .. + NC) (PITCH = 4, NC = 3: one 32-byte sector,
// fetched with a single 256-bit load); the NC rows of a node are adjacent, so a warp's stores are one contiguous run.
template <int NC, int PITCH>
__global__ void __launch_bounds__(256)
    node_gather_kernel(int64_t n_nodes, const int32_t *__restrict__ node_ptr, const uint32_t *__restrict__ node_ent,
                       const double *__restrict__ re, int overwrite, double *__restrict__ rhs) {
    const int64_t N = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (N >= n_nodes) return;
    const int32_t b = node_ptr[N], e = node_ptr[N + 1];
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    constexpr int U = 4;  // entries in flight
    for (int32_t k = b; k < e; k += U) {
        double v[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int c = 0; c < NC; ++c) v[u][c] = 0.0;
            if (k + u < e) {
                const double *p = re + (size_t)__ldg(node_ent + k + u) * PITCH;
                if constexpr (PITCH == 4 && NC == 3) {
                    double x, y, z, w;
                    asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
                    v[u][0] = x;
                    v[u][1] = y;
                    v[u][2] = z;
                } else {
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[u][c] = __ldg(p + c);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)  // absent entries add +0.0: same sums as the entry-by-entry loop
            if (k + u < e) {
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[c] += v[u][c];
            }
    }
    double *out = rhs + (size_t)N * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c) out[c] = overwrite ? acc[c] : out[c] + acc[c];
}

__global__ void scalar_pattern(const int32_t *__restrict__ bptr, const int32_t *__restrict__ bcols, int64_t n_nodes,
                               int nc, int32_t *__restrict__ indptr, int32_t *__restrict__ indices) {
    // one thread per (block row, block): writes its nc x nc column indices; first block of a row writes indptr
    int64_t N = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (N > n_nodes) return;
    if (N == n_nodes) {
        indptr[(int64_t)nc * n_nodes] = nc * nc * bptr[n_nodes];
        return;
    }
    int32_t b0 = bptr[N], nb = bptr[N + 1] - b0;
    for (int i = 0; i < nc; ++i) {
        int32_t start = nc * nc * b0 + i * nc * nb;
        indptr[nc * N + i] = start;
        for (int32_t k = 0; k < nb; ++k)
            for (int j = 0; j < nc; ++j) indices[start + nc * k + j] = nc * bcols[b0 + k] + j;
    }
}

}  // namespace

extern "C" void nfa_bplan_destroy(nfa_bplan *bp) { delete bp; }
extern "C" int64_t nfa_bplan_n_active_nodes(const nfa_bplan *bp) { return bp ? bp->n_active : -1; }
extern "C" int64_t nfa_bplan_nnz(const nfa_bplan *bp) { return bp ? bp->n_bnz * bp->n_comp * bp->n_comp : -1; }
extern "C" int64_t nfa_bplan_bytes(const nfa_bplan *bp) {
    if (!bp) return -1;
    return (int64_t)(bp->bptr.n + bp->bcols.n + bp->bseg.n + bp->chunk_row.n + bp->chunk_hdr.n + bp->contrib.n + bp->contrib_slot.n +
                     bp->indptr.n + bp->indices.n + bp->geo.n + bp->node_ptr.n + bp->node_ent.n + bp->tile_ptr.n + bp->pair_ptr.n +
                     bp->tile_ent.n + bp->node_pair_ptr.n + bp->node_pair.n + bp->pair_sum.n);
}

extern "C" int nfa_bplan_create(int64_t n_nodes, int n_phi, int n_comp, const int32_t *node_dest, int64_t n_cells,
                                nfa_bplan **out, nfa_stream stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!out) return fail("nfa_bplan_create: out is NULL");
    *out = nullptr;
    if (n_comp < 1 || n_comp > 3) return fail("nfa_bplan_create: n_comp must be 1, 2 or 3");
    if (n_phi < 1 || n_phi > 20) return fail("nfa_bplan_create: n_phi must be 1..20");
    int64_t p2 = (int64_t)n_phi * n_phi;
    int64_t E = n_cells * p2;
    if (E <= 0 || E >= ((int64_t)1 << 31)) return fail("nfa_bplan_create: contribution count out of range");
    if (n_nodes * n_comp > INT32_MAX) return fail("nfa_bplan_create: n_dof out of int32 range");
    nfa_bplan *bp = new nfa_bplan();
    auto bail = [&](int rc) {
        delete bp;
        return rc;
    };
    bp->n_nodes = n_nodes;
    bp->n_cells = n_cells;
    bp->n_phi = n_phi;
    bp->n_comp = n_comp;
    bp->n_contrib = E;

    DevBuf keys, keys_alt, ids, ids_alt, tmp, ukeys, counts, nruns;
    if (keys.alloc(8 * E) || keys_alt.alloc(8 * E) || ids.alloc(4 * E) || ids_alt.alloc(4 * E)) return bail(2);
    int pair_bits = 1;
    while ((1 << pair_bits) < p2) ++pair_bits;
    if ((n_cells << pair_bits) >= ((int64_t)1 << 32)) return bail(fail("nfa_bplan_create: too many cells for 32-bit contribution ids"));
    bp->pair_bits = pair_bits;
    fill_block_keys<<<grid_for(E, 256), 256, 0, s>>>(node_dest, n_cells, n_phi, n_nodes, pair_bits,
                                                     (uint64_t *)keys.p, (uint32_t *)ids.p);
    count_launch();
    int end_bit = 1;
    while (end_bit < 64 && (((uint64_t)n_nodes * (uint64_t)n_nodes) >> end_bit) != 0) ++end_bit;
    cub::DoubleBuffer<uint64_t> dk((uint64_t *)keys.p, (uint64_t *)keys_alt.p);
    cub::DoubleBuffer<uint32_t> dv((uint32_t *)ids.p, (uint32_t *)ids_alt.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)E, 0, end_bit, s);
    if (tmp.alloc(tb)) return bail(2);
    if (cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int)E, 0, end_bit, s) != cudaSuccess)
        return bail(fail("nfa_bplan_create: radix sort failed"));
    count_launch(8);
    if (ukeys.alloc(8 * E) || counts.alloc(4 * (E + 1)) || nruns.alloc(sizeof(int))) return bail(2);
    cub::DeviceRunLengthEncode::Encode(nullptr, tb, dk.Current(), (uint64_t *)ukeys.p, (uint32_t *)counts.p,
                                       (int *)nruns.p, (int)E, s);
    if (tmp.alloc(tb)) return bail(2);
    cub::DeviceRunLengthEncode::Encode(tmp.p, tb, dk.Current(), (uint64_t *)ukeys.p, (uint32_t *)counts.p,
                                       (int *)nruns.p, (int)E, s);
    count_launch(3);
    int h_runs = 0;
    cudaMemcpyAsync(&h_runs, nruns.p, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: run-length encode failed"));
    bp->n_bnz = h_runs;
    int64_t nb = h_runs;
    if (nb * n_comp * n_comp > INT32_MAX) return bail(fail("nfa_bplan_create: nnz exceeds int32 (PETSc index type)"));
    if (bp->bseg.alloc(4 * (nb + 1)) || bp->bcols.alloc(4 * nb) || bp->bptr.alloc(4 * (n_nodes + 1))) return bail(2);
    cudaMemsetAsync((uint32_t *)counts.p + nb, 0, 4, s);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t *)counts.p, (uint32_t *)bp->bseg.p, (int)(nb + 1), s);
    if (tmp.alloc(tb)) return bail(2);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, (uint32_t *)counts.p, (uint32_t *)bp->bseg.p, (int)(nb + 1), s);
    block_cols<<<grid_for(nb, 256), 256, 0, s>>>((uint64_t *)ukeys.p, nb, n_nodes, (int32_t *)bp->bcols.p);
    block_ptr<<<grid_for(n_nodes + 1, 256), 256, 0, s>>>((uint64_t *)ukeys.p, nb, n_nodes, (int32_t *)bp->bptr.p);
    count_launch(4);
    ukeys.release();
    counts.release();
    // keys (sorted) are no longer needed; dv.Current() holds contribution ids in output order
    // ---- chunking on the host: whole block rows, bounded contributions and blocks per chunk
    DevBuf rowc;
    if (rowc.alloc(4 * n_nodes)) return bail(2);
    row_contribs<<<grid_for(n_nodes, 256), 256, 0, s>>>((int32_t *)bp->bptr.p, (uint32_t *)bp->bseg.p, n_nodes,
                                                        (uint32_t *)rowc.p);
    count_launch();
    std::vector<uint32_t> h_rowc(n_nodes);
    std::vector<int32_t> h_bptr(n_nodes + 1);
    cudaMemcpyAsync(h_rowc.data(), rowc.p, 4 * n_nodes, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(h_bptr.data(), bp->bptr.p, 4 * (n_nodes + 1), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: block pattern build failed"));
    const int nc2 = n_comp * n_comp;
    int cap_c = 512, cap_b = 512;
    if (const char *e = getenv("NFA_FUSED_CAP")) cap_c = cap_b = std::max(64, atoi(e));
    // staging must fit: cap_c * nc2 doubles (+ Mref) within ~200 KB
    // chunks = runs of consecutive NON-EMPTY block rows (a rank's local plan leaves the rows of other ranks empty)
    std::vector<int32_t> chunk_r0, chunk_r1;
    std::vector<uint32_t> chunk_cstart;
    int64_t cur_c = 0, cur_b = 0;
    uint32_t run = 0;
    int max_c = 0, max_b = 0;
    bool open = false;
    auto close_at = [&](int64_t r) {
        chunk_r1.push_back((int32_t)r);
        max_c = std::max<int>(max_c, (int)cur_c);
        max_b = std::max<int>(max_b, (int)cur_b);
        cur_c = cur_b = 0;
        open = false;
    };
    for (int64_t r = 0; r < n_nodes; ++r) {
        int64_t rc = h_rowc[r], rb = h_bptr[r + 1] - h_bptr[r];
        if (rc == 0) {
            if (open) close_at(r);
            continue;
        }
        if (rc > 65535 || rc * nc2 * 8 > 180 * 1024)
            return bail(fail("nfa_bplan_create: a node has too many adjacent cells for the fused staging"));
        if (open && (cur_c + rc > cap_c || cur_b + rb > cap_b)) close_at(r);
        if (!open) {
            chunk_r0.push_back((int32_t)r);
            chunk_cstart.push_back(run);
            open = true;
        }
        cur_c += rc;
        cur_b += rb;
        run += (uint32_t)rc;
    }
    if (open) close_at(n_nodes);
    bp->n_chunks = (int64_t)chunk_r0.size();
    if (bp->n_chunks == 0) return bail(fail("nfa_bplan_create: no contributions"));
    bp->max_chunk_contrib = max_c;
    bp->max_chunk_bnz = max_b;
    DevBuf d_cstart;
    std::vector<int32_t> hdr(8 * (size_t)bp->n_chunks, 0);
    std::vector<int32_t> &chunk_row = chunk_r0;
    for (int64_t c = 0; c < bp->n_chunks; ++c) {
        int32_t ra = chunk_r0[c], rz = chunk_r1[c];
        uint32_t ca = chunk_cstart[c], cz = (c + 1 < bp->n_chunks) ? chunk_cstart[c + 1] : run;
        hdr[8 * c + 0] = ra;
        hdr[8 * c + 1] = rz - ra;
        hdr[8 * c + 2] = h_bptr[ra];
        hdr[8 * c + 3] = h_bptr[rz] - h_bptr[ra];
        hdr[8 * c + 4] = (int32_t)ca;
        hdr[8 * c + 5] = (int32_t)(cz - ca);
    }
    if (bp->chunk_row.alloc(4 * chunk_row.size()) || d_cstart.alloc(4 * chunk_cstart.size()) ||
        bp->chunk_hdr.alloc(4 * hdr.size()))
        return bail(2);
    cudaMemcpyAsync(bp->chunk_hdr.p, hdr.data(), 4 * hdr.size(), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(bp->chunk_row.p, chunk_row.data(), 4 * chunk_row.size(), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_cstart.p, chunk_cstart.data(), 4 * chunk_cstart.size(), cudaMemcpyHostToDevice, s);
    // ---- reorder each chunk's contributions by (cell, a, b) = ascending id, remembering the staging slot
    uint64_t *ck = (uint64_t *)dk.Alternate();  // reuse the sort buffers
    uint32_t *cs = dv.Alternate();
    chunk_keys<<<grid_for(E, 256), 256, 0, s>>>(dv.Current(), E, (uint32_t *)d_cstart.p, bp->n_chunks, ck, cs);
    count_launch();
    // sort (chunk, id) pairs; ping-pong buffers: keys -> dk.Current(), slots -> need one more u32 buffer
    DevBuf slots_alt;
    if (slots_alt.alloc(4 * E)) return bail(2);
    cub::DoubleBuffer<uint64_t> dk2(ck, dk.Current());
    cub::DoubleBuffer<uint32_t> dv2(cs, (uint32_t *)slots_alt.p);
    int cbits = 1;
    while (cbits < 32 && (((uint64_t)bp->n_chunks) >> cbits) != 0) ++cbits;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, dk2, dv2, (int)E, 0, 32 + cbits, s);
    if (tmp.alloc(tb)) return bail(2);
    if (cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk2, dv2, (int)E, 0, 32 + cbits, s) != cudaSuccess)
        return bail(fail("nfa_bplan_create: chunk sort failed"));
    count_launch(8);
    if (bp->contrib.alloc(4 * E) || bp->contrib_slot.alloc(2 * E)) return bail(2);
    unpack_chunk_sorted<<<grid_for(E, 256), 256, 0, s>>>(dk2.Current(), dv2.Current(), E, (uint32_t *)d_cstart.p,
                                                         (uint32_t *)bp->contrib.p, (uint16_t *)bp->contrib_slot.p);
    count_launch();
    // ---- node incidence lists (stable radix sort of the node ids: entries stay in ascending cell order)
    {
        const int64_t M = n_cells * n_phi;
        DevBuf nk, nk_alt, ni_alt;
        if (nk.alloc(4 * M) || nk_alt.alloc(4 * M) || ni_alt.alloc(4 * M) || bp->node_ent.alloc(4 * M) ||
            bp->node_ptr.alloc(4 * (n_nodes + 1)))
            return bail(2);
        fill_node_keys<<<grid_for(M, 256), 256, 0, s>>>(node_dest, M, (uint32_t *)nk.p, (uint32_t *)ni_alt.p);
        int nbits = 1;
        while (nbits < 32 && (((uint64_t)n_nodes) >> nbits) != 0) ++nbits;
        size_t tb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb2, (uint32_t *)nk.p, (uint32_t *)nk_alt.p, (uint32_t *)ni_alt.p,
                                        (uint32_t *)bp->node_ent.p, (int)M, 0, nbits, s);
        if (tmp.alloc(tb2)) return bail(2);
        if (cub::DeviceRadixSort::SortPairs(tmp.p, tb2, (uint32_t *)nk.p, (uint32_t *)nk_alt.p, (uint32_t *)ni_alt.p,
                                            (uint32_t *)bp->node_ent.p, (int)M, 0, nbits, s) != cudaSuccess)
            return bail(fail("nfa_bplan_create: node incidence sort failed"));
        node_ptr_kernel<<<grid_for(n_nodes + 1, 256), 256, 0, s>>>((uint32_t *)nk_alt.p, M, n_nodes, (int32_t *)bp->node_ptr.p);
        count_launch(6);
        if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: node incidence lists failed"));
    }
    // ---- tile-merged load lists (nfa_bplan_load_rhs)
    {
        constexpr int TL = nfa_bplan::LOAD_TILE;
        const int64_t M = n_cells * n_phi, n_tiles = (n_cells + TL - 1) / TL;
        DevBuf k64, k64_alt, id_a, id_b, uk, cnt, nr;
        if (k64.alloc(8 * M) || k64_alt.alloc(8 * M) || id_a.alloc(4 * M) || id_b.alloc(4 * M)) return bail(2);
        fill_tile_node_keys<<<grid_for(M, 256), 256, 0, s>>>(node_dest, M, n_phi, TL, n_nodes, (uint64_t *)k64.p,
                                                             (uint32_t *)id_a.p);
        int kb = 1;
        while (kb < 64 && (((uint64_t)n_tiles * (uint64_t)n_nodes) >> kb) != 0) ++kb;
        cub::DoubleBuffer<uint64_t> tk((uint64_t *)k64.p, (uint64_t *)k64_alt.p);
        cub::DoubleBuffer<uint32_t> tv((uint32_t *)id_a.p, (uint32_t *)id_b.p);
        size_t tb3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb3, tk, tv, (int)M, 0, kb, s);
        if (tmp.alloc(tb3)) return bail(2);
        if (cub::DeviceRadixSort::SortPairs(tmp.p, tb3, tk, tv, (int)M, 0, kb, s) != cudaSuccess)
            return bail(fail("nfa_bplan_create: tile / node sort failed"));
        if (uk.alloc(8 * M) || cnt.alloc(4 * (M + 1)) || nr.alloc(sizeof(int))) return bail(2);
        cub::DeviceRunLengthEncode::Encode(nullptr, tb3, tk.Current(), (uint64_t *)uk.p, (uint32_t *)cnt.p, (int *)nr.p,
                                           (int)M, s);
        if (tmp.alloc(tb3)) return bail(2);
        cub::DeviceRunLengthEncode::Encode(tmp.p, tb3, tk.Current(), (uint64_t *)uk.p, (uint32_t *)cnt.p, (int *)nr.p,
                                           (int)M, s);
        int h_pairs = 0;
        cudaMemcpyAsync(&h_pairs, nr.p, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: tile / node pairs failed"));
        const int64_t P = h_pairs;
        bp->n_pairs = P;
        if (bp->pair_ptr.alloc(4 * (P + 1)) || bp->tile_ent.alloc(2 * M) || bp->tile_ptr.alloc(4 * (n_tiles + 1)) ||
            bp->node_pair_ptr.alloc(4 * (n_nodes + 1)) || bp->node_pair.alloc(4 * P) ||
            bp->pair_sum.alloc(sizeof(double) * P * n_comp))
            return bail(2);
        cudaMemsetAsync((uint32_t *)cnt.p + P, 0, 4, s);
        cub::DeviceScan::ExclusiveSum(nullptr, tb3, (uint32_t *)cnt.p, (uint32_t *)bp->pair_ptr.p, (int)(P + 1), s);
        if (tmp.alloc(tb3)) return bail(2);
        cub::DeviceScan::ExclusiveSum(tmp.p, tb3, (uint32_t *)cnt.p, (uint32_t *)bp->pair_ptr.p, (int)(P + 1), s);
        tile_local_ids<<<grid_for(M, 256), 256, 0, s>>>(tv.Current(), M, n_phi, TL, (uint16_t *)bp->tile_ent.p);
        lower_bound_u64<<<grid_for(n_tiles + 1, 256), 256, 0, s>>>((uint64_t *)uk.p, P, n_tiles, (uint64_t)n_nodes,
                                                                   (int32_t *)bp->tile_ptr.p);
        // a node's pairs in ascending tile order: stable sort of the pair ids by node
        uint32_t *pk = (uint32_t *)id_a.p, *pk_alt = (uint32_t *)id_b.p;  // M >= P: reuse
        uint32_t *pi = (uint32_t *)k64.p, *pi_out = (uint32_t *)bp->node_pair.p;
        pair_node_keys<<<grid_for(P, 256), 256, 0, s>>>((uint64_t *)uk.p, P, n_nodes, pk, pi);
        int nbits = 1;
        while (nbits < 32 && (((uint64_t)n_nodes) >> nbits) != 0) ++nbits;
        cub::DeviceRadixSort::SortPairs(nullptr, tb3, pk, pk_alt, pi, pi_out, (int)P, 0, nbits, s);
        if (tmp.alloc(tb3)) return bail(2);
        if (cub::DeviceRadixSort::SortPairs(tmp.p, tb3, pk, pk_alt, pi, pi_out, (int)P, 0, nbits, s) != cudaSuccess)
            return bail(fail("nfa_bplan_create: pair sort failed"));
        node_ptr_kernel<<<grid_for(n_nodes + 1, 256), 256, 0, s>>>(pk_alt, P, n_nodes, (int32_t *)bp->node_pair_ptr.p);
        // the nodes that have cells at all (a rank's local plan numbers the GLOBAL nodes: the per-step kernels walk this list)
        DevBuf fl, na;
        if (fl.alloc(n_nodes) || na.alloc(sizeof(int)) || bp->active_nodes.alloc(4 * n_nodes)) return bail(2);
        nonempty_flags<<<grid_for(n_nodes, 256), 256, 0, s>>>((const int32_t *)bp->node_pair_ptr.p, n_nodes, (uint8_t *)fl.p);
        cub::CountingInputIterator<int32_t> iota(0);
        cub::DeviceSelect::Flagged(nullptr, tb3, iota, (uint8_t *)fl.p, (int32_t *)bp->active_nodes.p, (int *)na.p, (int)n_nodes, s);
        if (tmp.alloc(tb3)) return bail(2);
        cub::DeviceSelect::Flagged(tmp.p, tb3, iota, (uint8_t *)fl.p, (int32_t *)bp->active_nodes.p, (int *)na.p, (int)n_nodes, s);
        int h_active = 0;
        cudaMemcpyAsync(&h_active, na.p, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: active node list failed"));
        bp->n_active = h_active;
        count_launch(27);
        if (cudaStreamSynchronize(s) != cudaSuccess) return bail(fail("nfa_bplan_create: tile-merged load lists failed"));
    }
    // ---- scalar CSR pattern
    int64_t n_dof = n_nodes * n_comp, nnz = nb * nc2;
    if (bp->indptr.alloc(4 * (n_dof + 1)) || bp->indices.alloc(4 * nnz)) return bail(2);
    if (bp->geo.alloc(sizeof(double) * 12 * n_cells)) return bail(2);  // per-assembly geometry records
    scalar_pattern<<<grid_for(n_nodes + 1, 128), 128, 0, s>>>((int32_t *)bp->bptr.p, (int32_t *)bp->bcols.p, n_nodes,
                                                              n_comp, (int32_t *)bp->indptr.p, (int32_t *)bp->indices.p);
    count_launch();
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return bail(fail(std::string("nfa_bplan_create: ") + cudaGetErrorString(e)));
    *out = bp;
    return 0;
}

extern "C" int nfa_bplan_gather_load(const nfa_bplan *bp, const double *re, int comp_pitch, int overwrite, double *rhs,
                                     nfa_stream stream) {
    if (!bp || !re || !rhs) return fail("nfa_bplan_gather_load: NULL argument");
    if (comp_pitch != bp->n_comp && !(comp_pitch == 4 && bp->n_comp == 3))
        return fail("nfa_bplan_gather_load: comp_pitch must be n_comp (or 4 for three components)");
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t *np = (const int32_t *)bp->node_ptr.p;
    const uint32_t *ne = (const uint32_t *)bp->node_ent.p;
    const unsigned grid = (unsigned)grid_for(bp->n_nodes, 256);
    if (bp->n_comp == 3 && comp_pitch == 4)
        node_gather_kernel<3, 4><<<grid, 256, 0, s>>>(bp->n_nodes, np, ne, re, overwrite, rhs);
    else if (bp->n_comp == 3)
        node_gather_kernel<3, 3><<<grid, 256, 0, s>>>(bp->n_nodes, np, ne, re, overwrite, rhs);
    else if (bp->n_comp == 2)
        node_gather_kernel<2, 2><<<grid, 256, 0, s>>>(bp->n_nodes, np, ne, re, overwrite, rhs);
    else
        node_gather_kernel<1, 1><<<grid, 256, 0, s>>>(bp->n_nodes, np, ne, re, overwrite, rhs);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_plan_from_bplan(const nfa_bplan *bp, nfa_plan **out, nfa_stream) {
    if (!bp || !out) return fail("nfa_plan_from_bplan: NULL argument");
    nfa_plan *plan = new nfa_plan();
    plan->n_dof = bp->n_nodes * bp->n_comp;
    plan->nnz = bp->n_bnz * bp->n_comp * bp->n_comp;
    plan->ext_indptr = (const int32_t *)bp->indptr.p;
    plan->ext_indices = (const int32_t *)bp->indices.p;
    *out = plan;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// geometry records: [Jinv (DIM x DIM, the first DIM columns of the pseudo-inverse), det*c0, det*c1]
template <int DIM, int FORM>
struct GeoRec {
    static constexpr int N = (FORM == 0) ? (DIM * DIM + 2) : (DIM * (DIM + 1) / 2);
    static constexpr int STRIDE = (N + 1) & ~1;  // 16-byte aligned records
};

template <int DIM, int FORM>
__global__ void geo_records_kernel(const double *__restrict__ coords, const int32_t *__restrict__ cell_nodes,
                                   int64_t n_cells, const double *__restrict__ c0, int64_t s0,
                                   const double *__restrict__ c1, int64_t s1, double *__restrict__ geo) {
    int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    double X[DIM + 1][3];
    load_cell_coords<DIM>(coords, cell_nodes, cell, X);
    Geom<DIM> g;
    compute_geom<DIM>(X, g);
    double *rec = geo + cell * GeoRec<DIM, FORM>::STRIDE;
    if (FORM == 0) {
#pragma unroll
        for (int k = 0; k < DIM; ++k)
#pragma unroll
            for (int d = 0; d < DIM; ++d) rec[k * DIM + d] = g.Jinv[k][d];
        rec[DIM * DIM] = g.det * c0[cell * s0];
        rec[DIM * DIM + 1] = g.det * c1[cell * s1];
    } else {
        // G = det * kappa * Jinv Jinv^T (symmetric, upper triangle row-major); physical columns 0..DIM-1 only
        double w = g.det * c0[cell * s0];
        int t = 0;
#pragma unroll
        for (int k = 0; k < DIM; ++k)
#pragma unroll
            for (int l = k; l < DIM; ++l) {
                double v = 0.0;
#pragma unroll
                for (int d = 0; d < DIM; ++d) v += g.Jinv[k][d] * g.Jinv[l][d];
                rec[t++] = w * v;
            }
    }
}

// Persistent CTAs stride over the chunks.  Per chunk:
//   phase 1  one contribution per thread in cell order (geometry loads broadcast inside a warp); the block is staged
//            at the thread's own position (consecutive 16-byte aligned records: no bank conflicts);
//   phase 2a one CSR block per thread: gathers its contributions through gather_pos in ascending cell order into
//            registers;
//   phase 2b after a barrier the sums are written in CSR layout over the (now free) staging area;
//   phase 2c linear, fully coalesced streaming copy of that tile to the CSR values in HBM.
template <int DIM, int FORM, int MAXB>
__global__ void __launch_bounds__(FUSED_THREADS, NFA_FUSED_MINBLOCKS)
    fused_h1_kernel(int n_chunks, const int4 *__restrict__ chunk_hdr, const int32_t *__restrict__ bptr,
                    const uint32_t *__restrict__ bseg, const uint32_t *__restrict__ contrib,
                    const uint16_t *__restrict__ gather_pos, const double *__restrict__ geo,
                    const double *__restrict__ Mref, int n_phi, int pair_bits, int stage_cap, int block_cap,
                    double *__restrict__ vals) {
    constexpr int NC = (FORM == 0) ? DIM : 1;
    constexpr int NC2 = NC * NC;
    constexpr int SR = (NC2 + 1) & ~1;        // staging record stride (doubles), 16-byte aligned
    constexpr int D2 = DIM * DIM;
    constexpr int MR = (D2 + 1) & ~1;         // Mref record stride in shared memory
    constexpr int GS = GeoRec<DIM, FORM>::STRIDE;
    extern __shared__ double smem[];
    const int p2 = n_phi * n_phi;
    double *s_M = smem;                                        // [p2][MR]
    double *s_stage = s_M + p2 * MR;                           // [stage_cap][SR]; reused as the output tile
    int *s_bs = (int *)(s_stage + (size_t)stage_cap * SR);     // [n_blocks + 1] first contribution of each block
    int *s_rows = s_bs + block_cap + 2;                        // [n_rows + 1] block offsets (relative to b0)

    for (int i = threadIdx.x; i < p2 * D2; i += FUSED_THREADS) s_M[(i / D2) * MR + (i % D2)] = __ldg(Mref + i);
    const uint32_t pair_mask = (1u << pair_bits) - 1u;

    int ch = blockIdx.x;
    int4 h0 = make_int4(0, 0, 0, 0), h1 = h0;
    if (ch < n_chunks) {
        h0 = __ldg(chunk_hdr + 2 * ch);
        h1 = __ldg(chunk_hdr + 2 * ch + 1);
    }
    __syncthreads();

    for (; ch < n_chunks; ch += gridDim.x) {
        // chunk header: {r0, n_rows, b0, n_blocks}, {c0, n_contrib, -, -}
        const int r0 = h0.x, n_rows = h0.y, b0 = h0.z, n_blocks = h0.w;
        const uint32_t c0 = (uint32_t)h1.x, n_contrib = (uint32_t)h1.y;
        const int next = ch + gridDim.x;
        if (next < n_chunks) {  // prefetch the next header; consumed after this chunk's last barrier
            h0 = __ldg(chunk_hdr + 2 * next);
            h1 = __ldg(chunk_hdr + 2 * next + 1);
        }

        uint32_t k = threadIdx.x;
        uint32_t id = 0;
        if (k < n_contrib) id = __ldg(contrib + c0 + k);
        for (int i = threadIdx.x; i <= n_rows; i += FUSED_THREADS) s_rows[i] = __ldg(bptr + r0 + i) - b0;
        for (int i = threadIdx.x; i <= n_blocks; i += FUSED_THREADS) s_bs[i] = (int)(__ldg(bseg + b0 + i) - c0);

        // ---- phase 1
        while (k < n_contrib) {
            const uint32_t cell = id >> pair_bits;
            const int pair = (int)(id & pair_mask);
            const double2 *rec = reinterpret_cast<const double2 *>(geo + (int64_t)cell * GS);
            double g[GS];
#pragma unroll
            for (int q = 0; q < GS / 2; ++q) {
                double2 v = __ldg(rec + q);
                g[2 * q] = v.x;
                g[2 * q + 1] = v.y;
            }
            double2 *out = reinterpret_cast<double2 *>(s_stage + (size_t)k * SR);
            k += FUSED_THREADS;
            if (k < n_contrib) id = __ldg(contrib + c0 + k);  // prefetch while this one computes
            const double2 *M2 = reinterpret_cast<const double2 *>(s_M + pair * MR);
            double Mr[MR];
#pragma unroll
            for (int q = 0; q < MR / 2; ++q) {
                double2 v = M2[q];
                Mr[2 * q] = v.x;
                Mr[2 * q + 1] = v.y;
            }
            double res[SR];
            if constexpr (SR > NC2) res[SR - 1] = 0.0;
            if constexpr (FORM == 0) {
                const double wl = g[D2], wm = g[D2 + 1];
                // P = Jinv^T M (P[i][l] = sum_k Jinv[k][i] M[k][l]);  T = P Jinv (T[i][j] = sum_l P[i][l] Jinv[l][j])
                double P[DIM][DIM], T[DIM][DIM];
#pragma unroll
                for (int i = 0; i < DIM; ++i)
#pragma unroll
                    for (int l = 0; l < DIM; ++l) {
                        double v = 0.0;
#pragma unroll
                        for (int kk = 0; kk < DIM; ++kk) v += g[kk * DIM + i] * Mr[kk * DIM + l];
                        P[i][l] = v;
                    }
                double tr = 0.0;
#pragma unroll
                for (int i = 0; i < DIM; ++i) {
#pragma unroll
                    for (int j = 0; j < DIM; ++j) {
                        double v = 0.0;
#pragma unroll
                        for (int l = 0; l < DIM; ++l) v += P[i][l] * g[l * DIM + j];
                        T[i][j] = v;
                    }
                    tr += T[i][i];
                }
#pragma unroll
                for (int i = 0; i < DIM; ++i)
#pragma unroll
                    for (int j = 0; j < DIM; ++j) {
                        double v = wl * T[i][j] + wm * T[j][i];
                        if (i == j) v += wm * tr;
                        res[i * DIM + j] = v;
                    }
            } else {
                double v = 0.0;
                int t = 0;
#pragma unroll
                for (int kk = 0; kk < DIM; ++kk)
#pragma unroll
                    for (int l = kk; l < DIM; ++l) {
                        v += g[t] * (kk == l ? Mr[kk * DIM + l] : (Mr[kk * DIM + l] + Mr[l * DIM + kk]));
                        ++t;
                    }
                res[0] = v;
            }
#pragma unroll
            for (int q = 0; q < SR / 2; ++q) out[q] = make_double2(res[2 * q], res[2 * q + 1]);
        }
        __syncthreads();

        // ---- phase 2a: one block per thread, contributions in ascending cell order
        const uint16_t *gp = gather_pos + c0;
        double acc[MAXB][NC2];
        int where[MAXB], stride[MAXB];
#pragma unroll
        for (int m = 0; m < MAXB; ++m) {
            const int blk = threadIdx.x + m * FUSED_THREADS;
            where[m] = -1;
            stride[m] = 0;
#pragma unroll
            for (int e = 0; e < NC2; ++e) acc[m][e] = 0.0;
            if (blk < n_blocks) {
                int lo = 0, hi = n_rows;  // block row: s_rows[lo] <= blk < s_rows[lo + 1]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_rows[mid] <= blk)
                        lo = mid;
                    else
                        hi = mid;
                }
                const int rb = s_rows[lo], nb = s_rows[lo + 1] - rb;
                where[m] = NC2 * rb + NC * (blk - rb);
                stride[m] = NC * nb;
                const int k0 = s_bs[blk], k1 = s_bs[blk + 1];
                int pos = __ldg(gp + k0);
                for (int kk = k0; kk < k1; ++kk) {
                    const double2 *src = reinterpret_cast<const double2 *>(s_stage + (size_t)pos * SR);
                    if (kk + 1 < k1) pos = __ldg(gp + kk + 1);
                    double v[SR];
#pragma unroll
                    for (int q = 0; q < SR / 2; ++q) {
                        double2 t = src[q];
                        v[2 * q] = t.x;
                        v[2 * q + 1] = t.y;
                    }
#pragma unroll
                    for (int e = 0; e < NC2; ++e) acc[m][e] += v[e];
                }
            }
        }
        __syncthreads();
        // ---- phase 2b: sums in CSR layout over the staging area
#pragma unroll
        for (int m = 0; m < MAXB; ++m) {
            if (where[m] >= 0) {
#pragma unroll
                for (int i = 0; i < NC; ++i)
#pragma unroll
                    for (int j = 0; j < NC; ++j) s_stage[where[m] + i * stride[m] + j] = acc[m][i * NC + j];
            }
        }
        __syncthreads();
        // ---- phase 2c: coalesced streaming copy of the tile
        double *dst = vals + (int64_t)NC2 * b0;
        for (int v = threadIdx.x; v < n_blocks * NC2; v += FUSED_THREADS) __stcs(dst + v, s_stage[v]);
        __syncthreads();
    }
}

template <int DIM, int FORM>
static int run_fused(const nfa_bplan *bp, const double *Mref, const double *coords, const int32_t *cell_nodes,
                     const double *c0, int64_t s0, const double *c1, int64_t s1, double *vals, cudaStream_t s) {
    constexpr int NC = (FORM == 0) ? DIM : 1;
    if (bp->n_comp != NC) return fail("fused H1: plan n_comp does not match the form");
    constexpr int GS = GeoRec<DIM, FORM>::STRIDE;
    static_assert(GS <= 12 && GS % 2 == 0, "geometry record exceeds the plan's scratch stride");
    const DevBuf &geo = bp->geo;
    geo_records_kernel<DIM, FORM><<<grid_for(bp->n_cells, 128), 128, 0, s>>>(coords, cell_nodes, bp->n_cells, c0, s0,
                                                                             c1, s1, (double *)geo.p);
    NFA_LAUNCH_CHECK();
    int p2 = bp->n_phi * bp->n_phi;
    int stage_cap = bp->max_chunk_contrib, block_cap = bp->max_chunk_bnz;
    constexpr int SR = (NC * NC + 1) & ~1, MR = (DIM * DIM + 1) & ~1;
    size_t smem = sizeof(double) * ((size_t)p2 * MR + (size_t)stage_cap * SR) + sizeof(int) * (2 * (size_t)block_cap + 8);
    const int maxb = (block_cap + FUSED_THREADS - 1) / FUSED_THREADS;
    if (maxb > 4) return fail("fused H1: a chunk holds more blocks than the kernel is instantiated for");
    auto kern = maxb <= 1 ? fused_h1_kernel<DIM, FORM, 1>
                          : (maxb == 2 ? fused_h1_kernel<DIM, FORM, 2> : fused_h1_kernel<DIM, FORM, 4>);
    NFA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, device = 0, n_sm = 148;
    NFA_CUDA(cudaGetDevice(&device));
    NFA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    NFA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FUSED_THREADS, smem));
    if (per_sm < 1) return fail("fused H1: kernel does not fit on an SM");
    int64_t grid = std::min<int64_t>(bp->n_chunks, (int64_t)n_sm * per_sm);
    if (timing_enabled()) timing_begin(s);
    kern<<<(unsigned)grid, FUSED_THREADS, smem, s>>>(
        (int)bp->n_chunks, (const int4 *)bp->chunk_hdr.p, (const int32_t *)bp->bptr.p, (const uint32_t *)bp->bseg.p,
        (const uint32_t *)bp->contrib.p, (const uint16_t *)bp->contrib_slot.p, (const double *)geo.p, Mref, bp->n_phi,
        bp->pair_bits, stage_cap, block_cap, vals);
    if (timing_enabled()) timing_end(s);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_assemble_h1_elasticity_affine(const nfa_bplan *bp, int dim, const double *Mref,
                                                 const double *coords, const int32_t *cell_nodes, const double *lam,
                                                 int64_t ls, const double *mu, int64_t ms, double *vals,
                                                 nfa_stream stream) {
    if (!bp) return fail("nfa_assemble_h1_elasticity_affine: plan is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (dim == 3) return run_fused<3, 0>(bp, Mref, coords, cell_nodes, lam, ls, mu, ms, vals, s);
    if (dim == 2) return run_fused<2, 0>(bp, Mref, coords, cell_nodes, lam, ls, mu, ms, vals, s);
    return fail("nfa_assemble_h1_elasticity_affine: dim must be 2 or 3");
}

extern "C" int nfa_assemble_h1_laplace_affine(const nfa_bplan *bp, int dim, const double *Mref, const double *coords,
                                              const int32_t *cell_nodes, const double *kappa, int64_t ks, double *vals,
                                              nfa_stream stream) {
    if (!bp) return fail("nfa_assemble_h1_laplace_affine: plan is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (dim == 3) return run_fused<3, 1>(bp, Mref, coords, cell_nodes, kappa, ks, nullptr, 0, vals, s);
    if (dim == 2) return run_fused<2, 1>(bp, Mref, coords, cell_nodes, kappa, ks, nullptr, 0, vals, s);
    if (dim == 1) return run_fused<1, 1>(bp, Mref, coords, cell_nodes, kappa, ks, nullptr, 0, vals, s);
    return fail("nfa_assemble_h1_laplace_affine: dim must be 1, 2 or 3");
}
