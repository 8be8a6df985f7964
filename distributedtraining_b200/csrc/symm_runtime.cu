// Symmetric peer-memory runtime (host side, C ABI): one window per rank allocated with cudaMalloc, exported with CUDA
// IPC, mapped by every other rank of the box.  Kernels then issue ld/st on the mapped peer addresses directly over
// NVLink 5 / NVSwitch (see optim_avg.cu gather_avg_kernel, sm100_gemm.cu with a peer-resident B operand).
//
// This replaces the reference's tensor plane -- torch.save -> git-lfs push -> hf_hub_download -> torch.load
// (reference hivetrain/hf_manager.py:91-136, 186-197) -- for ranks co-located on one box.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

extern "C" int dtb_set_device(int dev) { return int(cudaSetDevice(dev)); }
extern "C" int dtb_get_device() {
  int d = -1;
  cudaGetDevice(&d);
  return d;
}
extern "C" int dtb_symm_alloc(size_t bytes, void** out) {
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) return int(e);
  return int(cudaMemset(*out, 0, bytes));
}
extern "C" int dtb_symm_free(void* p) { return int(cudaFree(p)); }
extern "C" int dtb_ipc_get_handle(void* p, char* out64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return int(e);
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(out64, &h, 64);
  return 0;
}
extern "C" int dtb_ipc_open_handle(const char* in64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, 64);
  return int(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
}
extern "C" int dtb_ipc_close_handle(void* p) { return int(cudaIpcCloseMemHandle(p)); }
extern "C" int dtb_can_access_peer(int dev, int peer) {
  int ok = 0;
  cudaDeviceCanAccessPeer(&ok, dev, peer);
  return ok;
}
extern "C" int dtb_memcpy_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  return int(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
}
extern "C" int dtb_last_error() { return int(cudaGetLastError()); }
extern "C" int dtb_read_u32(const uint32_t* dev_ptr, uint32_t* out) {
  return int(cudaMemcpy(out, dev_ptr, 4, cudaMemcpyDeviceToHost));
}
