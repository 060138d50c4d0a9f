// psk_p2p.cu -- peer-to-peer ghost-cell exchange for slab-decomposed grids (sm_100a).
//
// BASELINE.json configs[3]: ONE periodic Burgers grid cut into contiguous slabs, one process
// (and one B200) per slab.  Every RHS needs the g = 3 cells next to each slab edge from the
// neighbouring GPU: 24 bytes per side per stage, so the exchange is pure latency.  Instead of a
// send/recv pair per neighbour (host-launched NCCL kernels plus staging copies), the owner of
// the edge cells stores them straight into the neighbour's ghost slots through NVLink peer
// memory and then raises an epoch flag in the neighbour's memory; the neighbour spins on its
// own (local) flag before it launches the kernels that read the ghost cells.
//
//   psk_p2p_alloc / open / close / free   peer-visible device memory (CUDA IPC handles; the
//                                         64-byte handle travels through torch.distributed)
//   psk_halo_push                         edge cells -> neighbours' ghost slots, fence, flags
//   psk_halo_wait                         spin (bounded) until both local flags reach an epoch
//
// The reference is single-device (pyshocks/__init__.py:66 forces one CPU device); there is no
// reference counterpart to cite beyond the ghost-cell layout of grid.py:83-117.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "psk_common.cuh"

namespace psk {

__global__ void halo_push_kernel(const double *__restrict__ src_lo, double *dst_lo,
                                 const double *__restrict__ src_hi, double *dst_hi, int count,
                                 long long *flag_lo, long long *flag_hi, long long epoch) {
  for (int k = threadIdx.x; k < 2 * count; k += blockDim.x) {
    if (k < count) {
      if (dst_lo != nullptr) dst_lo[k] = src_lo[k];
    } else {
      if (dst_hi != nullptr) dst_hi[k - count] = src_hi[k - count];
    }
  }
  // every writer orders its peer stores before the flag that announces them
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (flag_lo != nullptr)
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flag_lo), "l"(epoch) : "memory");
    if (flag_hi != nullptr)
      asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(flag_hi), "l"(epoch) : "memory");
  }
}

__device__ __forceinline__ long long load_acquire_sys(const long long *p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One thread spins; a dead or late peer cannot hang the GPU: after timeout_ns the kernel gives
// up and records it (the host checks the flag when it next synchronises).
__global__ void halo_wait_kernel(const long long *flag_a, const long long *flag_b, long long epoch,
                                 unsigned long long timeout_ns, int *timed_out) {
  if (threadIdx.x != 0) return;
  const unsigned long long t0 = global_timer_ns();
  bool ok_a = (flag_a == nullptr), ok_b = (flag_b == nullptr);
  while (true) {
    if (!ok_a) ok_a = load_acquire_sys(flag_a) >= epoch;
    if (!ok_b) ok_b = load_acquire_sys(flag_b) >= epoch;
    if (ok_a && ok_b) return;
    // sticky: once a wait has given up, no later wait spends its timeout again
    if (timed_out != nullptr && *reinterpret_cast<volatile int *>(timed_out) != 0) return;
    if (global_timer_ns() - t0 > timeout_ns) {
      if (timed_out != nullptr) atomicExch(timed_out, 1);
      return;
    }
    __nanosleep(64);
  }
}

}  // namespace psk

using namespace psk;

extern "C" {

int psk_p2p_alloc(uint64_t bytes, void **ptr, unsigned char *handle) {
  if (ptr == nullptr || handle == nullptr || bytes == 0) return PSK_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == PSK_IPC_HANDLE_BYTES, "IPC handle size");
  void *p = nullptr;
  PSK_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e);
  }
  std::memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return PSK_OK;
}

int psk_p2p_free(void *ptr) {
  if (ptr == nullptr) return PSK_OK;
  PSK_CUDA_OK(cudaFree(ptr));
  return PSK_OK;
}

int psk_p2p_open(const unsigned char *handle, void **ptr) {
  if (ptr == nullptr || handle == nullptr) return PSK_E_INVALID;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void *p = nullptr;
  PSK_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return PSK_OK;
}

int psk_p2p_close(void *ptr) {
  if (ptr == nullptr) return PSK_OK;
  PSK_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return PSK_OK;
}

int psk_halo_push(const double *src_lo, double *dst_lo, const double *src_hi, double *dst_hi,
                  int32_t count, int64_t *flag_lo, int64_t *flag_hi, int64_t epoch,
                  psk_stream_t stream) {
  if (count <= 0) return PSK_E_INVALID;
  if ((dst_lo != nullptr && src_lo == nullptr) || (dst_hi != nullptr && src_hi == nullptr))
    return PSK_E_INVALID;
  const int threads = count <= 16 ? 32 : (count <= 64 ? 128 : 256);
  halo_push_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      src_lo, dst_lo, src_hi, dst_hi, count, reinterpret_cast<long long *>(flag_lo),
      reinterpret_cast<long long *>(flag_hi), static_cast<long long>(epoch));
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

int psk_halo_wait(const int64_t *flag_a, const int64_t *flag_b, int64_t epoch, int64_t timeout_ns,
                  int32_t *timed_out, psk_stream_t stream) {
  if (flag_a == nullptr && flag_b == nullptr) return PSK_E_INVALID;
  if (timeout_ns <= 0) return PSK_E_INVALID;
  halo_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long *>(flag_a), reinterpret_cast<const long long *>(flag_b),
      static_cast<long long>(epoch), static_cast<unsigned long long>(timeout_ns), timed_out);
  PSK_CUDA_OK(cudaGetLastError());
  return PSK_OK;
}

}  // extern "C"
