// Shared pieces of the flat-arena averaging kernels (optim_avg.cu, meta_avg.cu): limits, typed delta loaders, bounded flag waits.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

namespace dtb {

using bf16 = __nv_bfloat16;
constexpr int kMaxMiners = 64;
constexpr int kMaxOut = 16;

// Bounded device-side spin: ~200 ns per poll.  The limit is per translation unit (no -rdc) and set through
// dtb_set_flag_timeout_*(); default 2^26 polls ~= 13 s.  A timeout raises *error_flag (checked by the host once per round:
// parallel/symm.py ErrorMonitor) instead of letting a dead peer hang the box.
static __device__ long long g_spin_limit = 1ll << 26;

DTB_DEVICE bool wait_flag_ge(const uint32_t* flag, uint32_t value, int* error_flag) {
  long long spins = 0;
  const long long limit = g_spin_limit;
  while (ld_acquire_sys(flag) < value) {
    if (++spins > limit) {
      if (error_flag) *error_flag = 1;
      return false;
    }
    __nanosleep(200);
  }
  return true;
}

// 8 consecutive delta elements of one miner starting at element e, decoded to fp32.
// MODE 0: fp32, 1: bf16, 2: e4m3 with one fp32 scale per 32 elements.  Plain ld.global: measured 2x the NVLink
// throughput of sys-scope loads; visibility comes from the acquire on the miner's publish flag + round-parity buffers.
template <int MODE>
DTB_DEVICE void load_delta8_plain(const void* dptr, const float* sptr, size_t e, float* d) {
  if (MODE == 0) {
    const float4 q0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dptr) + e);
    const float4 q1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dptr) + e + 4);
    d[0] = q0.x; d[1] = q0.y; d[2] = q0.z; d[3] = q0.w; d[4] = q1.x; d[5] = q1.y; d[6] = q1.z; d[7] = q1.w;
  } else if (MODE == 1) {
    const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(dptr) + e);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __bfloat1622float2(h[k]);
      d[2 * k] = t.x;
      d[2 * k + 1] = t.y;
    }
  } else {
    const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(dptr) + e);
    const float sc = sptr[e >> 5];
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&q);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half_raw hr = __nv_cvt_fp8_to_halfraw(b[k], __NV_E4M3);
      d[k] = __half2float(*reinterpret_cast<const __half*>(&hr)) * sc;
    }
  }
}

// NVLS: one store replicated by the NVSwitch into every rank's window (multicast address), one load returning the sum.
DTB_DEVICE void multimem_st_v4(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

}  // namespace dtb
