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

// ------------------------------------------------------------------------------------------------------------------
// VMM + NVLS multicast runtime.  A window is a physical allocation (cuMemCreate, exportable as a POSIX fd) that every rank
// maps (a) at a unicast address per peer (P2P loads / stores over NVLink) and (b) -- when the box has NVSwitch multicast --
// through ONE multicast object that all ranks bind their window to: a `multimem.st` to the multicast address lands in every
// rank's window at the same offset, a `multimem.ld_reduce` returns the sum over the ranks (csrc/optim_avg.cu).  cudaMalloc +
// CUDA-IPC windows (above) cannot be bound to a multicast object, which is why round 1 needed a second, library-allocated
// arena for its NVLS plane.  The fd hand-over between the processes is done by the Python side (SCM_RIGHTS over a unix
// socket: parallel/symm.py); this file holds every driver call.  Driver entry points are resolved at run time
// (cudaGetDriverEntryPoint), so the library neither links libcuda nor needs it on a CPU-only box.
// ------------------------------------------------------------------------------------------------------------------
#include <cuda.h>

namespace {
template <typename F>
F drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<F>(p);
}
#define DRV(name) static auto f_##name = drv<decltype(&name)>(#name); if (!f_##name) return -100
CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
int map_rw(CUdeviceptr* va, size_t size, size_t align, CUmemGenericAllocationHandle h, int dev) {
  DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
  CUresult r = f_cuMemAddressReserve(va, size, align, 0, 0);
  if (r != CUDA_SUCCESS) return int(r);
  r = f_cuMemMap(*va, size, 0, h, 0);
  if (r != CUDA_SUCCESS) return int(r);
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  return int(f_cuMemSetAccess(*va, size, &acc, 1));
}
}  // namespace

// granularity the window size must be a multiple of (max of allocation and multicast granularities; 0 on error)
extern "C" size_t dtb_vmm_granularity(int dev, int world) {
  auto f_gran = drv<decltype(&cuMemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
  if (!f_gran) return 0;
  CUmemAllocationProp prop = alloc_prop(dev);
  size_t g = 0;
  if (f_gran(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return 0;
  auto f_mg = drv<decltype(&cuMulticastGetGranularity)>("cuMulticastGetGranularity");
  if (f_mg && world > 1) {
    CUmulticastObjectProp mp{};
    mp.numDevices = unsigned(world);
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (f_mg(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  return g;
}
extern "C" int dtb_mc_supported(int dev) {
  auto f_attr = drv<decltype(&cuDeviceGetAttribute)>("cuDeviceGetAttribute");
  if (!f_attr) return 0;
  int v = 0;
  if (f_attr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
  return v;
}
// allocate + map my window; returns the allocation handle (opaque), the unicast address and an exportable fd
extern "C" int dtb_vmm_alloc(size_t bytes, size_t align, int dev, unsigned long long* handle, void** ptr, int* fd) {
  DRV(cuMemCreate); DRV(cuMemExportToShareableHandle);
  CUmemAllocationProp prop = alloc_prop(dev);
  CUmemGenericAllocationHandle h;
  CUresult r = f_cuMemCreate(&h, bytes, &prop, 0);
  if (r != CUDA_SUCCESS) return int(r);
  CUdeviceptr va = 0;
  int rc = map_rw(&va, bytes, align, h, dev);
  if (rc) return rc;
  r = f_cuMemExportToShareableHandle(fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return int(r);
  *handle = (unsigned long long)h;
  *ptr = reinterpret_cast<void*>(va);
  return int(cudaMemset(*ptr, 0, bytes));
}
// map a peer's window (fd received from that rank) into my address space
extern "C" int dtb_vmm_import(int fd, size_t bytes, size_t align, int dev, void** ptr) {
  DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  CUresult r = f_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return int(r);
  CUdeviceptr va = 0;
  int rc = map_rw(&va, bytes, align, h, dev);
  *ptr = reinterpret_cast<void*>(va);
  return rc;
}
// multicast object: created by ONE rank (returns an fd to hand to the others), imported by the rest
extern "C" int dtb_mc_create(int world, size_t bytes, unsigned long long* mc, int* fd) {
  DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
  CUmulticastObjectProp mp{};
  mp.numDevices = unsigned(world);
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  CUresult r = f_cuMulticastCreate(&h, &mp);
  if (r != CUDA_SUCCESS) return int(r);
  r = f_cuMemExportToShareableHandle(fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return int(r);
  *mc = (unsigned long long)h;
  return 0;
}
extern "C" int dtb_mc_import(int fd, unsigned long long* mc) {
  DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  CUresult r = f_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return int(r);
  *mc = (unsigned long long)h;
  return 0;
}
extern "C" int dtb_mc_add_device(unsigned long long mc, int dev) {
  DRV(cuMulticastAddDevice);
  return int(f_cuMulticastAddDevice((CUmemGenericAllocationHandle)mc, dev));
}
// after EVERY rank has added its device: bind my window and map the multicast object
extern "C" int dtb_mc_bind_map(unsigned long long mc, unsigned long long mem, size_t bytes, size_t align, int dev, void** mc_ptr) {
  DRV(cuMulticastBindMem);
  CUresult r = f_cuMulticastBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem, 0, bytes, 0);
  if (r != CUDA_SUCCESS) return int(r);
  CUdeviceptr va = 0;
  int rc = map_rw(&va, bytes, align, (CUmemGenericAllocationHandle)mc, dev);
  *mc_ptr = reinterpret_cast<void*>(va);
  return rc;
}
