// halo.cu -- cross-process synchronisation for the NVLink peer-to-peer halo exchange of the
// row-slab driver (yolohtli_b200/slab.py, transport "p2p"): ghost rows are written straight
// into the neighbour's HBM through a CUDA-IPC mapping (cudaMemcpyAsync over NVLink), followed
// by a release of a sequence number in the neighbour's flag word; the neighbour's edge stream
// acquires it before its edge kernels read the ghosts.  No NCCL on that path.
#include "yh_common.cuh"

namespace {

__global__ void flag_set_kernel(volatile int *flag, int value) {
  __threadfence_system();          // the ghost-row copy ahead of us in the stream is complete; publish
  *flag = value;
  __threadfence_system();
}

// Spin until *flag >= value.  A peer that never signals must not hang the GPU: after ~4 s
// (clock64) the kernel records a timeout in *status and returns.
__global__ void flag_wait_kernel(volatile int *flag, int value, int *status) {
  const long long t0 = clock64();
  while (*flag < value) {
    if (clock64() - t0 > 8000000000ll) { *status = 1; break; }
    __nanosleep(200);
  }
  __threadfence_system();
}

}  // namespace

extern "C" {

int yh_flag_set(int *flag_peer, int value, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(flag_peer != nullptr, "null flag");
  flag_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_peer, value);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_flag_wait(int *flag_local, int value, int *status_local, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(flag_local && status_local, "null flag");
  flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_local, value, status_local);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_memcpy_async(void *dst, const void *src, size_t bytes, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(dst && src, "null pointer");
  YH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return YH_OK;
}

int yh_enable_peer_access(int peer_device) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  int dev = 0, can = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return YH_OK;
  YH_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) { yh_set_error("device %d cannot access peer %d", dev, peer_device); return YH_ERR_UNSUPPORTED; }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return YH_OK; }
  YH_CUDA(e);
  return YH_OK;
}

}  // extern "C"
