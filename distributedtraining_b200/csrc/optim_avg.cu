// Flat-arena kernels of the local-SGD algorithm (sm_100a):
//   * fused multi-tensor AdamW over the whole parameter arena (+ bf16 compute copy, + optional fused delta emit)
//   * delta emit  d = theta - theta_base   (fp32 / bf16 / block-scaled fp8 e4m3, 32-element blocks, fp32 scales)
//   * fused kernel (a): gather -> learned weighted sum -> base add, in ONE pass:
//         theta_new[e] = s_j * theta_base[e] + sum_i w[i,j] * delta_i[e],   s_j = sum_i w[i,j]
//     delta_i pointers may be PEER (NVLink-mapped) addresses: the kernel issues the P2P loads itself, optionally
//     waits on per-miner publish flags (ld.acquire.sys) and pushes the result to several destinations (peer stores),
//     which turns it into reduce-scatter + all-gather when every rank runs it on its shard.
//   * segmented multi-dot: G[i,j] = <g_j, theta_base_j + delta_ij - theta_avg_j>  (meta-gradient of the learned mixer)
//   * cross-GPU flag publish / device-side barrier
//
// Parity: reference hivetrain/training_manager.py:391,417-421 (AdamW, delta emit), hivetrain/averaging_logic.py:
// 422-448 (weighted average), :513-528 (meta-gradient), :121-127 (NaN screen); SURVEY.md K10, K13, K19-K24.
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "avg_common.cuh"

namespace dtb {

// ------------------------------------------------------------------------------------------------------------------
// AdamW
// ------------------------------------------------------------------------------------------------------------------
// hyper (device, fp32[12]): lr, beta1, beta2, eps, weight_decay, grad_scale, bias_corr1, bias_corr2, fresh, -, -, -
// fresh = 1 on the FIRST step after an optimizer (re-)creation: the moments are zero by definition, so the step kernel neither
// reads them nor needs them cleared beforehand -- the reset the reference performs after every base pull
// (hivetrain/training_manager.py:371-373) costs no memory traffic at all (1 GB less written per round, 1 GB less read per first step).
__global__ void adam_prep_kernel(int* step, float* hyper) {
  const int t = *step + 1;
  *step = t;
  hyper[6] = 1.f - powf(hyper[1], float(t));
  hyper[7] = 1.f - powf(hyper[2], float(t));
  hyper[8] = (t == 1) ? 1.f : 0.f;
}

template <int DELTA_MODE>  // 0: none, 1: fp32 delta, 2: bf16 delta
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ master, bf16* __restrict__ p16,
                                                    const float* __restrict__ grad, float* __restrict__ m,
                                                    float* __restrict__ v, const float* __restrict__ hyper,
                                                    const float* __restrict__ base, void* __restrict__ delta, size_t n4,
                                                    const float* __restrict__ fresh_src) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gs = hyper[5];
  const float step_size = lr * sqrtf(hyper[7]) / hyper[6];
  const float decay = 1.f - lr * wd;
  const bool fresh = hyper[8] != 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 mm = make_float4(0.f, 0.f, 0.f, 0.f), vv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!fresh) {
      mm = reinterpret_cast<float4*>(m)[i];
      vv = reinterpret_cast<float4*>(v)[i];
    }
    // first step after a base pull: theta IS theta_base -- read it from there, so the round never has to write the master arena
    float4 p = (fresh && fresh_src) ? reinterpret_cast<const float4*>(fresh_src)[i] : reinterpret_cast<float4*>(master)[i];
    float* gp = &g.x; float* mp = &mm.x; float* vp = &vv.x; float* pp = &p.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * gs;
      mp[k] = b1 * mp[k] + (1.f - b1) * gk;
      vp[k] = b2 * vp[k] + (1.f - b2) * gk * gk;
      pp[k] = (pp[k] - step_size * mp[k] / (sqrtf(vp[k]) + eps)) * decay;
    }
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(master)[i] = p;
    if (p16) {
      uint2 o;
      __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
    if (DELTA_MODE != 0) {
      const float4 b = reinterpret_cast<const float4*>(base)[i];
      const float4 d = make_float4(p.x - b.x, p.y - b.y, p.z - b.z, p.w - b.w);
      if (DELTA_MODE == 1) {
        reinterpret_cast<float4*>(delta)[i] = d;
      } else {
        uint2 o;
        __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(delta)[i] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// delta emit / casts
// ------------------------------------------------------------------------------------------------------------------
// mode 0: fp32, 1: bf16, 2: fp8 e4m3 with one fp32 scale per 32 elements (scale = amax/448; q = x/scale)
__global__ void __launch_bounds__(256) delta_emit_kernel(const float* __restrict__ master, const float* __restrict__ base,
                                                         void* __restrict__ out, float* __restrict__ scales, size_t n,
                                                         int mode, int* __restrict__ bad) {
  // each thread handles 8 consecutive elements; 4 threads cooperate on one 32-element fp8 block
  const size_t n8 = n / 8;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
    const float4 a0 = reinterpret_cast<const float4*>(master)[2 * i], a1 = reinterpret_cast<const float4*>(master)[2 * i + 1];
    const float4 b0 = reinterpret_cast<const float4*>(base)[2 * i], b1 = reinterpret_cast<const float4*>(base)[2 * i + 1];
    float d[8] = {a0.x - b0.x, a0.y - b0.y, a0.z - b0.z, a0.w - b0.w, a1.x - b1.x, a1.y - b1.y, a1.z - b1.z, a1.w - b1.w};
    if (bad) {  // NaN/Inf screen at the SOURCE (reference averaging_logic.py:121-127 screens after the download): the verdict
                // travels with the publish flag, so every consumer agrees on it without re-reading the delta
      float z = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) z += d[k] * 0.f;
      if (z != 0.f) *bad = 1;
    }
    if (mode == 0) {
      reinterpret_cast<float4*>(out)[2 * i] = make_float4(d[0], d[1], d[2], d[3]);
      reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(d[4], d[5], d[6], d[7]);
    } else if (mode == 1) {
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(d[2 * k], d[2 * k + 1]);
      reinterpret_cast<uint4*>(out)[i] = o;
    } else {
      float amax = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(d[k]));
      // the 4 threads of a block are lanes 4q..4q+3 (i is contiguous across lanes, n8 % 4 == 0 by arena alignment)
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
      const float scale = amax > 0.f ? amax * (1.f / 448.f) : 1.f;
      const float inv = 1.f / scale;
      uint2 o;
      uint8_t* q = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = __nv_cvt_float_to_fp8(d[k] * inv, __NV_SATFINITE, __NV_E4M3);
      reinterpret_cast<uint2*>(out)[i] = o;
      if ((i & 3) == 0) scales[i >> 2] = scale;
    }
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 p = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

// Round reset after a base pull: master = base, p16 = bf16(base), m = v = 0 (the reference re-creates AdamW after
// every pull: hivetrain/training_manager.py:371-377).  One pass.
__global__ void __launch_bounds__(256) round_reset_kernel(const float* __restrict__ base, float* __restrict__ master,
                                                          bf16* __restrict__ p16, float* __restrict__ m,
                                                          float* __restrict__ v, size_t n4, int reset_moments) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 p = reinterpret_cast<const float4*>(base)[i];
    reinterpret_cast<float4*>(master)[i] = p;
    if (p16) {
      uint2 o;
      __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
    if (reset_moments) {
      reinterpret_cast<float4*>(m)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(v)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fused kernel (a): gather -> weighted sum -> base add (-> multi-destination store)
// ------------------------------------------------------------------------------------------------------------------
struct AvgParams {
  const void* delta[kMaxMiners];       // per-miner delta window (local or peer-mapped)
  const float* dscale[kMaxMiners];     // per-miner fp8 block scales (mode 2)
  float* out_f32[kMaxOut];             // destinations (local and/or peer) for the fp32 result
  bf16* out_bf16[kMaxOut];             // optional bf16 compute copies
  const uint32_t* wait_flag[kMaxMiners];  // optional: spin until *wait_flag[i] >= wait_value before touching delta[i]
  const float* base;
  const float* w;                      // [N, P] row-major mixing weights
  const int64_t* chunk_start;
  const int32_t* chunk_len;
  const int32_t* chunk_tid;
  int* nan_flags;                      // [N], set to 1 if miner i's delta holds a non-finite value
  int* error_flag;                     // set to 1 on a flag-wait timeout
  int N, P, n_out, mode;               // mode: 0 fp32, 1 bf16, 2 fp8-block
  int chunk_begin, chunk_end;          // shard of the chunk table (or of chunk_ids) processed by this launch
  const int32_t* chunk_ids;            // optional indirection: process chunk_ids[chunk_begin..chunk_end)
  int unit_base;                       // 1: theta = base + sum_i w_i delta_i (delta APPLY), 0: s_j = sum_i w_ij (averaging)
  int ld_mode;                         // peer-load flavour, see ld_peer_v4
  uint32_t wait_value;
  const int* active;                   // optional [N] device mask (round_prepare_kernel): inactive miners are neither read nor weighted
  float* out_mc_f32;                   // optional MULTICAST destinations: one multimem.st lands the result in every rank's
  bf16* out_mc_bf16;                   //   window (the averaged-base broadcast rides on the averaging kernel, no second pass)
};

// Peer-load flavour (set once per process through DTB200_PEER_LD = sys | nc | weak; default sys):
//   sys  : ld.relaxed.sys     -- coherent with data a peer published before its release flag (strictly correct)
//   nc   : ld.global.nc       -- non-coherent streaming path (L1 no-allocate); valid because a window buffer is never
//                                 rewritten while a kernel that reads it is in flight (round-parity double buffering)
//   weak : plain ld.global
__device__ int g_peer_ld_mode = 0;
DTB_DEVICE uint4 ld_peer_v4(const void* ptr, int mode) {
  if (mode == 1) return ld_nc_v4(ptr);
  if (mode == 2) return *reinterpret_cast<const uint4*>(ptr);
  return ld_relaxed_sys_v4(ptr);
}

template <int MODE>
__device__ __forceinline__ void load_delta8(const AvgParams& p, int i, size_t e, float* d) {
  const int lm = p.ld_mode;
  if (MODE == 0) {
    const uint4 q0 = ld_peer_v4(reinterpret_cast<const float*>(p.delta[i]) + e, lm);
    const uint4 q1 = ld_peer_v4(reinterpret_cast<const float*>(p.delta[i]) + e + 4, lm);
    d[0] = __uint_as_float(q0.x); d[1] = __uint_as_float(q0.y); d[2] = __uint_as_float(q0.z); d[3] = __uint_as_float(q0.w);
    d[4] = __uint_as_float(q1.x); d[5] = __uint_as_float(q1.y); d[6] = __uint_as_float(q1.z); d[7] = __uint_as_float(q1.w);
  } else if (MODE == 1) {
    const uint4 q = ld_peer_v4(reinterpret_cast<const bf16*>(p.delta[i]) + e, lm);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __bfloat1622float2(h[k]);
      d[2 * k] = t.x;
      d[2 * k + 1] = t.y;
    }
  } else {
    uint2 q;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];"
                 : "=r"(q.x), "=r"(q.y)
                 : "l"(reinterpret_cast<const uint8_t*>(p.delta[i]) + e)
                 : "memory");
    float sc;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(sc) : "l"(p.dscale[i] + (e >> 5)) : "memory");
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&q);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half_raw hr = __nv_cvt_fp8_to_halfraw(b[k], __NV_E4M3);
      d[k] = __half2float(*reinterpret_cast<const __half*>(&hr)) * sc;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) gather_avg_kernel(const __grid_constant__ AvgParams p) {
  __shared__ float s_w[kMaxMiners + 1];
  __shared__ int s_idx[kMaxMiners];  // compacted list of the miners that take part (active mask)
  __shared__ int s_n;
  // --- wait for the producers' publish flags (fused "barrier-by-flag" instead of the reference's SHA polling) ---
  if (p.wait_value != 0) {
    if (threadIdx.x < p.N && p.wait_flag[threadIdx.x] != nullptr && !(p.active && !p.active[threadIdx.x]))
      wait_flag_ge(p.wait_flag[threadIdx.x], p.wait_value, p.error_flag);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < p.N; ++i)
      if (!p.active || p.active[i]) s_idx[n++] = i;
    s_n = n;
  }
  __syncthreads();
  const int NA = s_n;
  int bad = 0;  // bitmask (per thread) of miners with non-finite data; N <= 64 -> two 32-bit words
  int bad_hi = 0;
  for (int ci = p.chunk_begin + blockIdx.x; ci < p.chunk_end; ci += gridDim.x) {
    const int c = p.chunk_ids ? p.chunk_ids[ci] : ci;
    const int j = p.chunk_tid[c];
    const size_t start = size_t(p.chunk_start[c]);
    const int len = p.chunk_len[c];
    __syncthreads();
    if (threadIdx.x < NA) s_w[threadIdx.x] = p.w[size_t(s_idx[threadIdx.x]) * p.P + j];
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < NA; ++i) s += s_w[i];
      s_w[kMaxMiners] = p.unit_base ? 1.f : s;
    }
    __syncthreads();
    const float s_sum = s_w[kMaxMiners];
    for (int v8 = threadIdx.x; v8 * 8 < len; v8 += blockDim.x) {
      const size_t e = start + size_t(v8) * 8;
      const float4 b0 = reinterpret_cast<const float4*>(p.base + e)[0];
      const float4 b1 = reinterpret_cast<const float4*>(p.base + e)[1];
      float acc[8] = {b0.x * s_sum, b0.y * s_sum, b0.z * s_sum, b0.w * s_sum, b1.x * s_sum, b1.y * s_sum, b1.z * s_sum, b1.w * s_sum};
      int i = 0;
      for (; i + 4 <= NA; i += 4) {  // 4 miners' loads in flight per thread before the FMAs
        const int m0 = s_idx[i], m1 = s_idx[i + 1], m2 = s_idx[i + 2], m3 = s_idx[i + 3];
        float d0[8], d1[8], d2[8], d3[8];
        load_delta8<MODE>(p, m0, e, d0);
        load_delta8<MODE>(p, m1, e, d1);
        load_delta8<MODE>(p, m2, e, d2);
        load_delta8<MODE>(p, m3, e, d3);
        const float w0 = s_w[i], w1 = s_w[i + 1], w2 = s_w[i + 2], w3 = s_w[i + 3];
        float chk0 = 0.f, chk1 = 0.f, chk2 = 0.f, chk3 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[k] += w0 * d0[k] + w1 * d1[k] + w2 * d2[k] + w3 * d3[k];
          chk0 += d0[k] * 0.f; chk1 += d1[k] * 0.f; chk2 += d2[k] * 0.f; chk3 += d3[k] * 0.f;  // NaN/Inf -> NaN
        }
        if (chk0 != 0.f) { if (m0 < 32) bad |= 1 << m0; else bad_hi |= 1 << (m0 - 32); }
        if (chk1 != 0.f) { if (m1 < 32) bad |= 1 << m1; else bad_hi |= 1 << (m1 - 32); }
        if (chk2 != 0.f) { if (m2 < 32) bad |= 1 << m2; else bad_hi |= 1 << (m2 - 32); }
        if (chk3 != 0.f) { if (m3 < 32) bad |= 1 << m3; else bad_hi |= 1 << (m3 - 32); }
      }
      for (; i < NA; ++i) {
        const int m0 = s_idx[i];
        float d0[8];
        load_delta8<MODE>(p, m0, e, d0);
        const float w0 = s_w[i];
        float chk0 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[k] += w0 * d0[k];
          chk0 += d0[k] * 0.f;
        }
        if (chk0 != 0.f) { if (m0 < 32) bad |= 1 << m0; else bad_hi |= 1 << (m0 - 32); }
      }
      const float4 o0 = make_float4(acc[0], acc[1], acc[2], acc[3]), o1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
      uint4 ob;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&ob);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
      for (int o = 0; o < p.n_out; ++o) {
        if (p.out_f32[o]) {
          reinterpret_cast<float4*>(p.out_f32[o] + e)[0] = o0;
          reinterpret_cast<float4*>(p.out_f32[o] + e)[1] = o1;
        }
        if (p.out_bf16[o]) *reinterpret_cast<uint4*>(p.out_bf16[o] + e) = ob;
      }
      if (p.out_mc_f32) {
        multimem_st_v4(p.out_mc_f32 + e, *reinterpret_cast<const uint4*>(&o0));
        multimem_st_v4(p.out_mc_f32 + e + 4, *reinterpret_cast<const uint4*>(&o1));
      }
      if (p.out_mc_bf16) multimem_st_v4(p.out_mc_bf16 + e, ob);
    }
  }
  if (p.nan_flags) {
    if (bad) for (int i = 0; i < 32 && i < p.N; ++i) if (bad & (1 << i)) p.nan_flags[i] = 1;
    if (bad_hi) for (int i = 32; i < p.N; ++i) if (bad_hi & (1 << (i - 32))) p.nan_flags[i] = 1;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// all-gather-by-PULL fused with the round reset.  After the reduce-scatter launch of gather_avg_kernel every rank holds
// the new base only for ITS shard of the chunk table (in its window).  This kernel reads each chunk from its owner's
// window (P2P loads: measured 780 GB/s, vs ~210 GB/s for the push form) and in the same pass writes the local fp32 base,
// the fp32 master, the bf16 compute copy and clears the Adam moments -- the optimizer re-creation after a base pull
// (reference hivetrain/training_manager.py:365-378) rides on the all-gather.
// ------------------------------------------------------------------------------------------------------------------
struct ShardPullParams {
  const float* shard_src[kMaxMiners];     // base window of every rank (peer-mapped)
  const uint32_t* wait_flag[kMaxMiners];  // local flag words: shard owner r has published round wait_value
  const int64_t* chunk_start;
  const int32_t* chunk_len;
  float* base_out;
  float* master;
  bf16* p16;
  float* m;
  float* v;
  int* error_flag;
  int world, num_chunks, chunks_per_rank, reset_moments;
  uint32_t wait_value;
};

__global__ void __launch_bounds__(256) shard_pull_reset_kernel(const __grid_constant__ ShardPullParams p) {
  if (p.wait_value != 0) {
    if (threadIdx.x < p.world && p.wait_flag[threadIdx.x] != nullptr)
      wait_flag_ge(p.wait_flag[threadIdx.x], p.wait_value, p.error_flag);
    __syncthreads();
  }
  for (int c = blockIdx.x; c < p.num_chunks; c += gridDim.x) {
    const int owner = min(c / p.chunks_per_rank, p.world - 1);
    const float* src = p.shard_src[owner];
    const size_t start = size_t(p.chunk_start[c]);
    const int len = p.chunk_len[c];
    // a full chunk is 4096 elements = 1024 float4 = 4 per thread: all four (peer) loads are issued before the first store
    const int n4 = len >> 2;
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v4 = threadIdx.x + k * 256;
      if (v4 < n4) x[k] = *reinterpret_cast<const float4*>(src + start + size_t(v4) * 4);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v4 = threadIdx.x + k * 256;
      if (v4 >= n4) continue;
      const size_t e = start + size_t(v4) * 4;
      if (p.base_out) *reinterpret_cast<float4*>(p.base_out + e) = x[k];
      *reinterpret_cast<float4*>(p.master + e) = x[k];
      if (p.p16) {
        uint2 o;
        __nv_bfloat162 lo = __floats2bfloat162_rn(x[k].x, x[k].y), hi = __floats2bfloat162_rn(x[k].z, x[k].w);
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.p16 + e) = o;
      }
      if (p.reset_moments) {
        *reinterpret_cast<float4*>(p.m + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(p.v + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// device-side model checksum (replaces the reference's sha256 over a 498 MB D2H copy per miner: validation_logic.py:133,
// 198-203).  64-bit Fletcher-style pair (sum of words, sum of position-weighted words) over the raw fp32 bits; the
// integer atomics make it order independent, hence deterministic.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) checksum_kernel(const uint32_t* __restrict__ x, size_t n, unsigned long long* out) {
  unsigned long long a = 0, b = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const unsigned long long v = x[i];
    a += v;
    b += v * ((i & 0xFFFFFull) + 1ull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, a);
    atomicAdd(out + 1, b);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// cross-GPU flags
// ------------------------------------------------------------------------------------------------------------------
struct FlagParams {
  uint32_t* dst[kMaxMiners];  // one flag word per destination rank (peer-mapped)
  int n;
  uint32_t value;
  const int* cond;            // optional device predicate: publish ``value`` if *cond != 0, else 0
  uint32_t* mc_dst;           // optional MULTICAST address of the flag word: ONE release store through the NVSwitch reaches every
                              // rank's flag page -- the same path the multimem.st DATA of the preceding kernel took, and a
                              // release, so the flag cannot overtake the data it announces
};
// Publish: make all prior writes of this GPU visible system-wide, then release-store the round number to every peer.
__global__ void publish_flag_kernel(const __grid_constant__ FlagParams p) {
  __threadfence_system();
  const uint32_t v = (p.cond == nullptr || *p.cond != 0) ? p.value : 0u;
  if (p.mc_dst) {
    if (threadIdx.x == 0) asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(p.mc_dst), "r"(v) : "memory");
    return;
  }
  if (threadIdx.x < p.n && p.dst[threadIdx.x]) st_release_sys(p.dst[threadIdx.x], v);
}
// Wait until all of flags[0..n) >= value (local polling).
__global__ void wait_flags_kernel(const uint32_t* flags, int n, int stride, uint32_t value, int* error_flag) {
  if (threadIdx.x < n) wait_flag_ge(flags + size_t(threadIdx.x) * stride, value, error_flag);
}

// Wait until flags[0..n) >= *target (target lives in device memory: the same launch serves every round of a captured graph).
__global__ void wait_flags_dev_kernel(const uint32_t* flags, int n, const uint32_t* target, int* error_flag) {
  if (threadIdx.x < n) wait_flag_ge(flags + threadIdx.x, *reinterpret_cast<const volatile uint32_t*>(target), error_flag);
}

}  // namespace dtb

using namespace dtb;

// ------------------------------------------------------------------------------------------------------------------
// NVLS path: in-switch reduction + multicast broadcast (NVSwitch SHARP) for the UNIFORM / pre-scaled mixers.
//   every rank owns one contiguous shard of the arena and, for it,
//     sum   = multimem.ld_reduce.add(delta_mc)      -- ONE load returns sum_i delta_i: the switch pulls the N copies
//     theta = base_scale * base + scale * sum
//     multimem.st(out_mc, theta)                    -- ONE store lands in every rank's landing buffer
// so a rank receives |arena| / N reduced bytes and |arena| (N-1)/N broadcast bytes, half the ingress of the pull round
// (which reads N-1 full shards twice), and issues no per-peer loop at all.  delta_mc / out_mc are MULTICAST addresses of
// symmetric allocations (torch.distributed._symmetric_memory: cuMulticast objects), base is this rank's local copy.
// Replaces the reference's N downloads + N x 148 ATen axpys (hivetrain/averaging_logic.py:431-448) for mixers that
// do not need per-miner weights learned on the averager.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__global__ void __launch_bounds__(256) nvls_avg_kernel(const float* __restrict__ delta_mc, float* __restrict__ out_mc,
                                                       const float* __restrict__ base, size_t lo4, size_t hi4, size_t base4,
                                                       float scale, float base_scale) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = lo4 + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    // the symmetric buffers are padded to a multiple of the world size; the local base copy is not
    const float4 b = i < base4 ? __ldg(reinterpret_cast<const float4*>(base) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 sum = multimem_ld_reduce_add(delta_mc + 4 * i);
    float4 o;
    o.x = fmaf(base_scale, b.x, scale * sum.x);
    o.y = fmaf(base_scale, b.y, scale * sum.y);
    o.z = fmaf(base_scale, b.z, scale * sum.z);
    o.w = fmaf(base_scale, b.w, scale * sum.w);
    multimem_st(out_mc + 4 * i, o);
  }
}

#define KCHECK() (cudaGetLastError() == cudaSuccess ? 0 : 1)

extern "C" int dtb_adam_prep(int* step, float* hyper, cudaStream_t s) {
  adam_prep_kernel<<<1, 1, 0, s>>>(step, hyper);
  return KCHECK();
}
extern "C" int dtb_adamw(float* master, void* p16, const float* grad, float* m, float* v, const float* hyper, const float* base,
                         void* delta, int delta_mode, size_t n, int num_sms, cudaStream_t s, const float* fresh_src) {
  const size_t n4 = n / 4;
  const int grid = num_sms * 8;
  if (delta_mode == 1) adamw_kernel<1><<<grid, 256, 0, s>>>(master, (bf16*)p16, grad, m, v, hyper, base, delta, n4, fresh_src);
  else if (delta_mode == 2) adamw_kernel<2><<<grid, 256, 0, s>>>(master, (bf16*)p16, grad, m, v, hyper, base, delta, n4, fresh_src);
  else adamw_kernel<0><<<grid, 256, 0, s>>>(master, (bf16*)p16, grad, m, v, hyper, nullptr, nullptr, n4, fresh_src);
  return KCHECK();
}
extern "C" int dtb_delta_emit(const float* master, const float* base, void* out, float* scales, size_t n, int mode, int num_sms,
                              cudaStream_t s, int* bad) {
  delta_emit_kernel<<<num_sms * 8, 256, 0, s>>>(master, base, out, scales, n, mode, bad);
  return KCHECK();
}
extern "C" int dtb_cast_f32_bf16(const float* src, void* dst, size_t n, int num_sms, cudaStream_t s) {
  cast_f32_bf16_kernel<<<num_sms * 8, 256, 0, s>>>(src, (bf16*)dst, n / 4);
  return KCHECK();
}
extern "C" int dtb_round_reset(const float* base, float* master, void* p16, float* m, float* v, size_t n, int reset_moments,
                               int num_sms, cudaStream_t s) {
  round_reset_kernel<<<num_sms * 8, 256, 0, s>>>(base, master, (bf16*)p16, m, v, n / 4, reset_moments);
  return KCHECK();
}

// deltas / dscales / wait_flags: host arrays of N device pointers.  outs_f32 / outs_bf16: host arrays of n_out pointers.
extern "C" int dtb_gather_avg(const void** deltas, const float** dscales, const uint32_t** wait_flags, uint32_t wait_value,
                              const float* base, const float* w, const int64_t* chunk_start, const int32_t* chunk_len,
                              const int32_t* chunk_tid, int chunk_begin, int chunk_end, float** outs_f32, void** outs_bf16,
                              int n_out, int* nan_flags, int* error_flag, int N, int P, int mode, int grid, cudaStream_t s,
                              const int32_t* chunk_ids, int unit_base, const int* active, float* out_mc_f32, void* out_mc_bf16) {
  if (N > kMaxMiners || n_out > kMaxOut) return 3;
  AvgParams p{};
  p.active = active;
  p.out_mc_f32 = out_mc_f32;
  p.out_mc_bf16 = (bf16*)out_mc_bf16;
  for (int i = 0; i < N; ++i) {
    p.delta[i] = deltas[i];
    p.dscale[i] = dscales ? dscales[i] : nullptr;
    p.wait_flag[i] = wait_flags ? wait_flags[i] : nullptr;
  }
  for (int o = 0; o < n_out; ++o) {
    p.out_f32[o] = outs_f32 ? outs_f32[o] : nullptr;
    p.out_bf16[o] = outs_bf16 ? (bf16*)outs_bf16[o] : nullptr;
  }
  p.base = base; p.w = w; p.chunk_start = chunk_start; p.chunk_len = chunk_len; p.chunk_tid = chunk_tid;
  p.nan_flags = nan_flags; p.error_flag = error_flag; p.N = N; p.P = P; p.n_out = n_out; p.mode = mode;
  p.chunk_begin = chunk_begin; p.chunk_end = chunk_end; p.wait_value = wait_flags ? wait_value : 0;
  p.chunk_ids = chunk_ids; p.unit_base = unit_base;
  {
    static int mode = -1;
    if (mode < 0) {
      const char* e = getenv("DTB200_PEER_LD");
      mode = (e && e[0] == 'n') ? 1 : ((e && e[0] == 's') ? 0 : 2);  // default: weak (2x the link throughput of sys-scope loads, measured)
    }
    p.ld_mode = mode;
  }
  if (grid > chunk_end - chunk_begin) grid = chunk_end - chunk_begin;
  if (grid < 1) return 0;
  if (mode == 0) gather_avg_kernel<0><<<grid, 256, 0, s>>>(p);
  else if (mode == 1) gather_avg_kernel<1><<<grid, 256, 0, s>>>(p);
  else gather_avg_kernel<2><<<grid, 256, 0, s>>>(p);
  return KCHECK();
}

extern "C" int dtb_set_flag_timeout_optim(double seconds) {
  const long long polls = seconds <= 0 ? (1ll << 62) : (long long)(seconds / 200e-9);
  return cudaMemcpyToSymbol(g_spin_limit, &polls, sizeof(polls)) == cudaSuccess ? 0 : 1;
}
extern "C" int dtb_publish_flag(uint32_t** dsts, int n, uint32_t value, cudaStream_t s, const int* cond, void* mc_dst) {
  if (n > kMaxMiners) return 3;
  FlagParams p{};
  for (int i = 0; i < n; ++i) p.dst[i] = dsts[i];
  p.n = n; p.value = value; p.cond = cond; p.mc_dst = (uint32_t*)mc_dst;
  publish_flag_kernel<<<1, 64, 0, s>>>(p);
  return KCHECK();
}
extern "C" int dtb_wait_flags(const uint32_t* flags, int n, int stride, uint32_t value, int* error_flag, cudaStream_t s) {
  wait_flags_kernel<<<1, 64, 0, s>>>(flags, n, stride, value, error_flag);
  return KCHECK();
}

extern "C" int dtb_wait_flags_dev(const uint32_t* flags, int n, const uint32_t* target, int* error_flag, cudaStream_t s) {
  wait_flags_dev_kernel<<<1, 64, 0, s>>>(flags, n, target, error_flag);
  return KCHECK();
}

extern "C" int dtb_shard_pull_reset(const float** shard_src, const uint32_t** wait_flags, uint32_t wait_value,
                                    const int64_t* chunk_start, const int32_t* chunk_len, int num_chunks, int chunks_per_rank,
                                    int world, float* base_out, float* master, void* p16, float* m, float* v, int reset_moments,
                                    int* error_flag, int grid, cudaStream_t s) {
  if (world > kMaxMiners) return 3;
  ShardPullParams p{};
  for (int r = 0; r < world; ++r) {
    p.shard_src[r] = shard_src[r];
    p.wait_flag[r] = wait_flags ? wait_flags[r] : nullptr;
  }
  p.chunk_start = chunk_start; p.chunk_len = chunk_len; p.base_out = base_out; p.master = master; p.p16 = (bf16*)p16;
  p.m = m; p.v = v; p.error_flag = error_flag; p.world = world; p.num_chunks = num_chunks; p.chunks_per_rank = chunks_per_rank;
  p.reset_moments = reset_moments; p.wait_value = wait_flags ? wait_value : 0;
  if (grid > num_chunks) grid = num_chunks;
  shard_pull_reset_kernel<<<grid, 256, 0, s>>>(p);
  return KCHECK();
}

// delta_mc / out_mc: multicast addresses; [lo4, hi4): this rank's shard in units of float4
extern "C" int dtb_nvls_avg(const float* delta_mc, float* out_mc, const float* base, size_t lo4, size_t hi4, size_t base4,
                            float scale, float base_scale, int grid, cudaStream_t s) {
  if (hi4 <= lo4) return 0;
  nvls_avg_kernel<<<grid, 256, 0, s>>>(delta_mc, out_mc, base, lo4, hi4, base4, scale, base_scale);
  return KCHECK();
}

extern "C" int dtb_checksum(const void* x, size_t n_words, unsigned long long* out2, int num_sms, cudaStream_t s) {
  cudaMemsetAsync(out2, 0, 16, s);
  checksum_kernel<<<num_sms * 4, 256, 0, s>>>((const uint32_t*)x, n_words, out2);
  return KCHECK();
}
