// Plan objects behind the opaque handles of include/nfa.h.
#pragma once
#include "nfa_common.cuh"

#define NFA_MAX_SEGMENTS 16

namespace nfa {

struct DevBuf {
    void *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t bytes);
    void release();
};

// Deterministic gather list: output u sums entries perm[seg_ptr[u] .. seg_ptr[u+1]) (ascending entry id =
// reference call order) into position pos[u] (identity when pos is empty).
struct GatherSet {
    int64_t n_unique = 0;
    DevBuf seg_ptr;  // uint32 [n_unique + 1]
    DevBuf perm;     // uint32 [n_entries]
    DevBuf pos;      // int32  [n_unique] or empty
};

}  // namespace nfa

struct nfa_plan {
    int64_t n_dof = 0, nnz = 0;
    nfa::DevBuf indptr, indices;     // int32, owned unless base/bplan-backed
    const nfa_plan *base = nullptr;  // pattern owner when merged into another plan
    const int32_t *ext_indptr = nullptr, *ext_indices = nullptr;  // pattern owned by a block plan
    int n_segments = 0;
    int64_t n_items[NFA_MAX_SEGMENTS] = {0};
    int32_t n_loc[NFA_MAX_SEGMENTS] = {0};
    int with_matrix[NFA_MAX_SEGMENTS] = {0};
    int with_rhs[NFA_MAX_SEGMENTS] = {0};  // 0: the segment's right-hand side is summed elsewhere (nfa_bplan_gather_load)
    int64_t mat_off[NFA_MAX_SEGMENTS + 1] = {0}, rhs_off[NFA_MAX_SEGMENTS + 1] = {0};
    nfa::GatherSet mat, rhs;

    const int32_t *d_indptr() const {
        if (base) return base->d_indptr();
        return ext_indptr ? ext_indptr : (const int32_t *)indptr.p;
    }
    const int32_t *d_indices() const {
        if (base) return base->d_indices();
        return ext_indices ? ext_indices : (const int32_t *)indices.p;
    }
};

// Node-block plan of the fused H1 path.  Block row N = node N; its columns (node ids, ascending) are
// bcols[bptr[N] .. bptr[N+1]).  The scalar CSR follows: rows n_comp*N + i, columns n_comp*M + j.
// Work is cut into CTA chunks of whole block rows; a chunk lists its contributions (cell, a, b) ordered by
// cell so that geometry loads broadcast, each with the staging slot that orders the sum by ascending cell.
struct nfa_bplan {
    int64_t n_nodes = 0, n_cells = 0, n_bnz = 0, n_contrib = 0;
    int n_phi = 0, n_comp = 0, pair_bits = 0;
    int64_t n_chunks = 0;
    int max_chunk_contrib = 0, max_chunk_bnz = 0;
    nfa::DevBuf bptr;         // int32 [n_nodes + 1]
    nfa::DevBuf bcols;        // int32 [n_bnz]
    nfa::DevBuf bseg;         // uint32 [n_bnz + 1]  contribution range of each block (output order)
    nfa::DevBuf chunk_row;    // int32 [n_chunks + 1] first block row of each chunk
    nfa::DevBuf chunk_hdr;    // int32 [n_chunks][8]: r0, n_rows, b0, n_blocks, c0, n_contrib, 0, 0
    nfa::DevBuf contrib;      // uint32 [n_contrib]  (cell-ordered within chunk): cell * n_phi^2 + a * n_phi + b
    nfa::DevBuf contrib_slot; // uint16 [n_contrib]  gather_pos: for the k-th contribution in output order, its
                              // position in the chunk's cell order (= its staging slot)
    nfa::DevBuf indptr, indices;  // scalar CSR pattern int32
    nfa::DevBuf geo;              // scratch: per-cell geometry records [n_cells][12], rewritten every assembly
    // node incidence lists (nfa_bplan_gather_load): node N is local node a of cell c for the entries c * n_phi + a in
    // node_ent[node_ptr[N] .. node_ptr[N+1]), ascending = the order of the reference's res_g[dest] += r_el
    nfa::DevBuf node_ptr;         // int32 [n_nodes + 1]
    nfa::DevBuf node_ent;         // uint32 [n_cells * n_phi]
    // Tile-merged load (nfa_bplan_load_rhs): cells are cut into tiles of LOAD_TILE consecutive cells; a PAIR is a (tile,
    // node) with at least one (cell, local node) of the tile on that node.  Pairs are ordered by (tile, node); the entries
    // of a pair (ascending cell) are tile_ent[pair_ptr[p] .. pair_ptr[p+1]) as uint16 local ids (cell - tile * LOAD_TILE) *
    // n_phi + a.  A node's pairs (ascending tile) are node_pair[node_pair_ptr[N] .. node_pair_ptr[N+1]).
#ifndef NFA_LOAD_TILE
#define NFA_LOAD_TILE 64
#endif
    static constexpr int LOAD_TILE = NFA_LOAD_TILE;
    int64_t n_pairs = 0;
    nfa::DevBuf tile_ptr;         // int32 [n_tiles + 1]  first pair of each tile
    nfa::DevBuf pair_ptr;         // uint32 [n_pairs + 1]
    nfa::DevBuf tile_ent;         // uint16 [n_cells * n_phi]
    nfa::DevBuf node_pair_ptr;    // int32 [n_nodes + 1]
    nfa::DevBuf node_pair;        // uint32 [n_pairs]
    nfa::DevBuf pair_sum;         // scratch: double [n_pairs][n_comp], rewritten every assembly
    int64_t n_active = 0;         // nodes with at least one cell (all of them on one GPU; a rank's share when sharded)
    nfa::DevBuf active_nodes;     // int32 [n_active], ascending
    // Row-item lists of the vector-P2 tetrahedron kernel (nfa_p2tet.cu), built by nfempy_b200/assembly/p2_items.py
    struct P2Items {
        int64_t n_chunks = 0, n_tasks = 0, n_items = 0, ell_size = 0;
        int max_tile_blocks = 0;
        nfa::DevBuf chunk_hdr;  // int32 [n_chunks][4]: b0, n_blocks, task0, n_tasks
        nfa::DevBuf task_hdr;   // int32 [n_tasks][4]: type | kmax << 8, ell_base, item_base, n_items
        nfa::DevBuf item_hdr;   // uint16 [n_slots][4]: tile offsets of up to three blocks, row pitch; type E: + a second slot
        nfa::DevBuf ell;        // uint32 [ell_size]: cell << 6 | i << 4 | j << 2 | x, 32 lanes interleaved
        nfa::DevBuf rec;        // scratch: geometry records [n_cells][4][4], rewritten every assembly
        nfa::DevBuf cta_ptr;    // optional int32 [n_split + 1]: chunk range of every CTA when the grid has n_split CTAs
        int64_t n_split = 0;
        nfa::DevBuf cell_mask;  // optional uint8 [n_cells]: 0 = the cell contributes zeros (halo cell of a sharded plan)
    } p2;
};
