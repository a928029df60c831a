// Shared helpers for libnfa (sm_100a).  See include/nfa.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/nfa.h"

namespace nfa {

void set_error(const std::string &msg);
void count_launch(int64_t n = 1);
bool timing_enabled();
void timing_begin(cudaStream_t s);
void timing_end(cudaStream_t s);

inline int fail(const std::string &msg) {
    set_error(msg);
    return 1;
}

#define NFA_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            nfa::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)

#define NFA_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        nfa::count_launch();                                                                        \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            nfa::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));                \
            return 3;                                                                               \
        }                                                                                           \
    } while (0)

// Per-cell affine geometry of a dim-simplex embedded in 3-space.
// J[r][c] = X_{c+1}[r] - X_0[r] (column c = dx/dxi_c), det = sqrt(det(J^T J)) >= 0 (|det J| for dim = 3),
// Jinv = (J^T J)^{-1} J^T  (dim x 3).  Matches R/src/geometry/mapping.py:23-36 up to round-off: the QR there
// yields |det R| = sqrt(det(J^T J)) and R^{-1} Q^T = J^+.
template <int DIM>
struct Geom {
    double J[3][DIM];
    double Jinv[DIM][3];
    double det;
};

template <int DIM>
__device__ __forceinline__ void compute_geom(const double (&X)[DIM + 1][3], Geom<DIM> &g) {
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) g.J[r][c] = X[c + 1][r] - X[0][r];
    if constexpr (DIM == 3) {
        const double (&J)[3][3] = g.J;
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        double d = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        double id = 1.0 / d;
        g.Jinv[0][0] = c00 * id;
        g.Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
        g.Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
        g.Jinv[1][0] = c01 * id;
        g.Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
        g.Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
        g.Jinv[2][0] = c02 * id;
        g.Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
        g.Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
        g.det = fabs(d);
    } else if constexpr (DIM == 2) {
        double a = 0, b = 0, c = 0;  // G = J^T J = [[a, b], [b, c]]
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            a += g.J[r][0] * g.J[r][0];
            b += g.J[r][0] * g.J[r][1];
            c += g.J[r][1] * g.J[r][1];
        }
        double dg = a * c - b * b;
        double id = 1.0 / dg;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            g.Jinv[0][r] = (c * g.J[r][0] - b * g.J[r][1]) * id;
            g.Jinv[1][r] = (a * g.J[r][1] - b * g.J[r][0]) * id;
        }
        g.det = sqrt(dg);
    } else {
        double a = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) a += g.J[r][0] * g.J[r][0];
#pragma unroll
        for (int r = 0; r < 3; ++r) g.Jinv[0][r] = g.J[r][0] / a;
        g.det = sqrt(a);
    }
}

// |det J| only (no inverse, no division): sqrt(det(J^T J)), i.e. |det J| for DIM = 3.
template <int DIM>
__device__ __forceinline__ double compute_det(const double (&X)[DIM + 1][3]) {
    double J[3][DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) J[r][c] = X[c + 1][r] - X[0][r];
    if constexpr (DIM == 3) {
        double d = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                   J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        return fabs(d);
    } else if constexpr (DIM == 2) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            a += J[r][0] * J[r][0];
            b += J[r][0] * J[r][1];
            c += J[r][1] * J[r][1];
        }
        return sqrt(a * c - b * b);
    } else {
        double a = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) a += J[r][0] * J[r][0];
        return sqrt(a);
    }
}

template <int DIM>
__device__ __forceinline__ void load_cell_coords(const double *__restrict__ coords,
                                                 const int32_t *__restrict__ cell_nodes, int64_t cell,
                                                 double (&X)[DIM + 1][3]) {
#pragma unroll
    for (int v = 0; v <= DIM; ++v) {
        int64_t n = cell_nodes[cell * (DIM + 1) + v];
#pragma unroll
        for (int r = 0; r < 3; ++r) X[v][r] = __ldg(coords + 3 * n + r);
    }
}

inline int grid_for(int64_t work, int block) { return (int)((work + block - 1) / block); }

}  // namespace nfa
