// libnfa: streamed scatter for forms whose local matrices cannot all be resident (sm_100a).
//
// BASELINE config 4 (LCEScaledDualWeakForm at BDM_3 x RT_3 x P_2 x P_2, R/src/weak_forms/lce_scaled_dual_weak_form.py:445-699)
// has 348 local dofs: a cell's local matrix is 0.97 MB and the gather lists of the two-stage plan (nfa_plan_create: one
// 32-bit id per local entry, sorted) would be larger than the matrix itself.  In these spaces every dof belongs to at
// most TWO cells (facet dofs of the H(div) fields; everything else is cell-local), so the CSR position of a local entry
// (cell c, a, b) follows from small tables instead of a global sort:
//     row     = dest[c][a]
//     columns = the sorted union of dest[c] and dest[partner of c across the facet that owns a]  (dest[c] alone for a
//               cell-local row)
//     pos     = indptr[row] + rank of dest[c][b] in that union        (rank tables: one uint16 per (table, b))
// The host layer (nfempy_b200/assembly/streamed.py) builds the tables once; cells are then walked in chunks: local kernel
// -> nfa_stream_scatter pass 0 (stores) -> pass 1 (adds).  An entry is STORED by the lower cell of its row, or by the
// higher cell when the lower cell has no such column; the higher cell ADDS to what the lower one stored.  Cells are walked
// in ascending order, so every value is the sum of its (at most two) contributions in ascending cell order -- the order
// of the reference's Mat.setValue(..., addv=True) calls, R/src/assembly/ScatterFormData.py:5-11 -- without atomics.
// indptr is 64-bit: 73 002 tetrahedra already hold 8.7e9 stored entries.
#include "nfa_common.cuh"

using namespace nfa;

namespace {

// one warp per row: out[indptr[r] .. indptr[r+1]) = src[src_off[r] ..)
__global__ void rows_copy_i32_kernel(int64_t n_rows, const int64_t *__restrict__ indptr, const int64_t *__restrict__ src_off,
                                     const int32_t *__restrict__ src, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows;
         r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t b = indptr[r], n = indptr[r + 1] - b, s = src_off[r];
        for (int64_t k = lane; k < n; k += 32) out[b + k] = src[s + k];
    }
}

// one CTA per (cell, local row a); threads walk the local columns b
__global__ void __launch_bounds__(128)
    stream_scatter_kernel(int n_loc, const int32_t *__restrict__ dest, const int32_t *__restrict__ tab_of,
                          const uint16_t *__restrict__ tabs, const int64_t *__restrict__ indptr,
                          const double *__restrict__ Ke, double *__restrict__ vals, int pass) {
    const int64_t ca = blockIdx.x;  // cell * n_loc + a
    const int64_t row = dest[ca];
    const uint16_t *tab = tabs + (int64_t)tab_of[ca] * n_loc;
    const int64_t base = indptr[row];
    const double *k = Ke + ca * n_loc;
    for (int b = threadIdx.x; b < n_loc; b += blockDim.x) {
        const unsigned t = tab[b];
        const int64_t pos = base + (t & 0x7FFFu);
        if ((int)(t >> 15) != pass) continue;
        if (pass == 0)
            vals[pos] = k[b];
        else
            vals[pos] += k[b];
    }
}

}  // namespace

extern "C" int nfa_rows_copy_i32(int64_t n_rows, const int64_t *indptr, const int64_t *src_off, const int32_t *src,
                                 int32_t *out, nfa_stream stream) {
    if (n_rows <= 0) return 0;
    if (!indptr || !src_off || !src || !out) return fail("nfa_rows_copy_i32: NULL argument");
    const int grid = (int)std::min<int64_t>(grid_for(n_rows * 32, 256), 148 * 16);
    rows_copy_i32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_rows, indptr, src_off, src, out);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_stream_scatter(int64_t n_cells, int n_loc, const int32_t *dest, const int32_t *tab_of,
                                  const uint16_t *tabs, const int64_t *indptr, const double *Ke, double *vals, int pass,
                                  nfa_stream stream) {
    if (n_cells <= 0) return 0;
    if (!dest || !tab_of || !tabs || !indptr || !Ke || !vals) return fail("nfa_stream_scatter: NULL argument");
    if (pass != 0 && pass != 1) return fail("nfa_stream_scatter: pass must be 0 (stores) or 1 (adds)");
    if (n_loc < 1 || n_loc > 32767) return fail("nfa_stream_scatter: n_loc out of range");
    const int64_t blocks = n_cells * n_loc;
    if (blocks > 0x7FFFFFFF) return fail("nfa_stream_scatter: chunk too large (n_cells * n_loc must stay below 2^31)");
    stream_scatter_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(n_loc, dest, tab_of, tabs, indptr, Ke, vals,
                                                                              pass);
    NFA_LAUNCH_CHECK();
    return 0;
}
