// libnfa: interface-row exchange of the cell-partitioned assembly over NVLink peer memory (sm_100a).
//
// After the local fused assembly a rank holds partial sums for the matrix rows a LOWER rank owns (the
// reference is single-process, this is the multi-GPU extension).  Instead of packing them into a staging buffer and
// handing that to ncclSend / ncclRecv from the host, the sending rank's kernel stores the rows' CSR segments and
// right-hand-side entries DIRECTLY into the owner's receive area (peer stores through NVLink / NVSwitch, the area is a
// cudaMalloc allocation mapped into the sender with CUDA IPC) and then publishes an epoch flag there; the owner's kernel
// waits for the flag and adds the payload to its rows.  Everything is stream-ordered device work -- no host wait, no
// collective call -- so the whole step (local kernels + exchange) replays from one CUDA graph.
//
// Protocol per directed pair (sender s -> owner o), all counters monotonic 64-bit epochs:
//   s: wait ack[o<-s] >= e            (the owner has consumed payload e, so the area may be overwritten)
//      store payload, __threadfence_system(); the last block publishes flag[s->o] = e + 1 (release, system scope)
//   o: wait flag[s->o] >= e + 1 (acquire, system scope); add payload to vals / rhs (rows are distinct: deterministic;
//      the caller launches the pulls in ascending sender rank); the last block publishes ack[o<-s] = e + 1 at the sender
// The sender's / owner's own epoch e lives in their local state words and advances with every launch.
#include "nfa_plan.cuh"

using namespace nfa;

namespace {

// A peer that never arrives must not hang the device: a wait gives up after ~2^25 polls (tens of seconds) and counts the
// event here; nfa_p2p_timeouts() reports it (the sums are then incomplete and the caller must treat the step as failed).
__device__ unsigned int g_p2p_timeouts = 0;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// state[0] = launches completed, state[1] = blocks finished in the current launch
__global__ void push_rows_kernel(const int32_t *__restrict__ indptr, const double *__restrict__ vals,
                                 const double *__restrict__ rhs, const int32_t *__restrict__ rows,
                                 const int64_t *__restrict__ buf_off, int64_t n_rows, double *remote_buf,
                                 unsigned long long *remote_flag, const unsigned long long *local_ack,
                                 unsigned long long *state) {
    __shared__ unsigned long long epoch;
    if (threadIdx.x == 0) {
        const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(state);
        unsigned int polls = 0;
        while (ld_acquire_sys(local_ack) < e) {
            __nanosleep(500);
            if (++polls > (1u << 25)) {
                atomicAdd(&g_p2p_timeouts, 1u);
                break;
            }
        }
        epoch = e;
    }
    __syncthreads();
    const int64_t n_mat = buf_off[n_rows];
    const int lane = threadIdx.x & 31;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_rows;
         k += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int32_t r = rows[k];
        const int32_t b = indptr[r], e = indptr[r + 1];
        const int64_t o = buf_off[k];
        for (int32_t t = b + lane; t < e; t += 32) remote_buf[o + (t - b)] = vals[t];
        if (lane == 0) remote_buf[n_mat + k] = rhs[r];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(state + 1, 1ULL);
        if (done == gridDim.x - 1) {
            state[1] = 0;
            __threadfence_system();
            st_release_sys(remote_flag, epoch + 1);
            *reinterpret_cast<volatile unsigned long long *>(state) = epoch + 1;
        }
    }
}

__global__ void pull_add_rows_kernel(const int32_t *__restrict__ indptr, double *__restrict__ vals, double *__restrict__ rhs,
                                     const int32_t *__restrict__ rows, const int64_t *__restrict__ buf_off, int64_t n_rows,
                                     const double *local_buf, const unsigned long long *local_flag,
                                     unsigned long long *remote_ack, unsigned long long *state) {
    __shared__ unsigned long long epoch;
    if (threadIdx.x == 0) {
        const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(state);
        unsigned int polls = 0;
        while (ld_acquire_sys(local_flag) < e + 1) {
            __nanosleep(500);
            if (++polls > (1u << 25)) {
                atomicAdd(&g_p2p_timeouts, 1u);
                break;
            }
        }
        epoch = e;
    }
    __syncthreads();
    const int64_t n_mat = buf_off[n_rows];
    const int lane = threadIdx.x & 31;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_rows;
         k += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int32_t r = rows[k];
        const int32_t b = indptr[r], e = indptr[r + 1];
        const int64_t o = buf_off[k];
        for (int32_t t = b + lane; t < e; t += 32) vals[t] += __ldcg(local_buf + o + (t - b));
        if (lane == 0) rhs[r] += __ldcg(local_buf + n_mat + k);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(state + 1, 1ULL);
        if (done == gridDim.x - 1) {
            state[1] = 0;
            __threadfence_system();
            st_release_sys(remote_ack, epoch + 1);
            *reinterpret_cast<volatile unsigned long long *>(state) = epoch + 1;
        }
    }
}

}  // namespace

// ---- peer-mappable allocations -----------------------------------------------------------------------------------------
extern "C" int nfa_p2p_alloc(int64_t bytes, void **d_ptr, void *handle64) {
    if (bytes <= 0 || !d_ptr || !handle64) return fail("nfa_p2p_alloc: bad argument");
    void *p = nullptr;
    NFA_CUDA(cudaMalloc(&p, (size_t)bytes));
    NFA_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
        cudaFree(p);
        return fail(std::string("nfa_p2p_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(err));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return 0;
}

extern "C" int nfa_p2p_open(const void *handle64, void **d_peer_ptr) {
    if (!handle64 || !d_peer_ptr) return fail("nfa_p2p_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) return fail(std::string("nfa_p2p_open: cudaIpcOpenMemHandle: ") + cudaGetErrorString(err));
    *d_peer_ptr = p;
    return 0;
}

extern "C" int nfa_p2p_close(void *d_peer_ptr) {
    if (d_peer_ptr) NFA_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
    return 0;
}

extern "C" int nfa_p2p_free(void *d_ptr) {
    if (d_ptr) NFA_CUDA(cudaFree(d_ptr));
    return 0;
}

extern "C" int64_t nfa_p2p_timeouts() {
    unsigned int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_p2p_timeouts, sizeof(v)) != cudaSuccess) return -1;
    return (int64_t)v;
}

// ---- the exchange kernels ----------------------------------------------------------------------------------------------
extern "C" int nfa_push_rows(const int32_t *indptr, const double *vals, const double *rhs, const int32_t *rows,
                             const int64_t *buf_off, int64_t n_rows, double *remote_buf, uint64_t *remote_flag,
                             const uint64_t *local_ack, uint64_t *state, nfa_stream stream) {
    if (n_rows <= 0) return 0;
    if (!remote_buf || !remote_flag || !local_ack || !state) return fail("nfa_push_rows: NULL exchange pointer");
    const int grid = (int)std::min<int64_t>(grid_for(n_rows * 32, 256), 296);
    push_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(indptr, vals, rhs, rows, buf_off, n_rows, remote_buf,
                                                             (unsigned long long *)remote_flag,
                                                             (const unsigned long long *)local_ack,
                                                             (unsigned long long *)state);
    NFA_LAUNCH_CHECK();
    return 0;
}

extern "C" int nfa_pull_add_rows(const int32_t *indptr, double *vals, double *rhs, const int32_t *rows,
                                 const int64_t *buf_off, int64_t n_rows, const double *local_buf,
                                 const uint64_t *local_flag, uint64_t *remote_ack, uint64_t *state, nfa_stream stream) {
    if (n_rows <= 0) return 0;
    if (!local_buf || !local_flag || !remote_ack || !state) return fail("nfa_pull_add_rows: NULL exchange pointer");
    const int grid = (int)std::min<int64_t>(grid_for(n_rows * 32, 256), 296);
    pull_add_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(indptr, vals, rhs, rows, buf_off, n_rows, local_buf,
                                                                 (const unsigned long long *)local_flag,
                                                                 (unsigned long long *)remote_ack,
                                                                 (unsigned long long *)state);
    NFA_LAUNCH_CHECK();
    return 0;
}
