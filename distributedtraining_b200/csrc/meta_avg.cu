// Distributed learned mixer (sm_100a): the averager's meta-learning loop sharded over ALL ranks of the box.
//
// Reference: hivetrain/averaging_logic.py:490-541 runs, on ONE machine, meta_epochs^2 x ceil(100/B) sequential SGD steps on
// the mixing matrix w[N, P]; every step rebuilds theta_bar = sum_i w_ij (theta_base_j + delta_ij) from 2N disk loads, does a
// fwd/bwd on a validation batch and N x P dot products G_ij = <g_j, theta_ij - theta_bar_j>.  Here (parallel/meta.py):
//
//   round start   round_prepare_kernel   wait for every miner's publish flag, read the NaN verdicts the miners attached to
//                                        their publishes, build the active mask, (re-)initialise w = 1/N_active
//                 shard_transpose_kernel rank k pulls ITS shard [e0, e1) of every miner's delta over NVLink ONCE (all-to-all
//                                        by pull, decoded to fp32) -> every later pass over the deltas is local HBM, 1/R of it
//   meta-step     gather_avg_kernel      (optim_avg.cu) theta_bar shard from the local delta shards -> fp32 shard + bf16 shard
//                 shard_pull16_kernel    all-gather of the bf16 theta_bar shards by pull (the only per-step NVLink traffic in
//                                        replicate mode: |theta| * 2 B * (R-1)/R per rank)
//                 fwd/bwd                replicated on every rank, or data-parallel over the validation rows
//                 seg_dot_kernel         G partial of rank k's shard: <sum_r c_r g_r, delta_i> for all i + the common term
//                                        <g, base - theta_bar>; in data-parallel mode the R gradient arenas are PEER pointers,
//                                        i.e. the gradient reduce-scatter is fused into the dot (no all-reduce of g)
//                 seg_dot_finish_kernel  per-tensor reduction, result stored into EVERY rank's slot table (peer stores, KBs)
//                 w_update_kernel        waits for all R partials, w -= lr * sum_r partial_r in a fixed order -> w stays
//                                        bit-identical on all ranks without a broadcast
//
// SURVEY.md K19-K23, section 7.4.5.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "avg_common.cuh"

namespace dtb {

// ------------------------------------------------------------------------------------------------------------------
// round prepare
// ------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const uint32_t* delta_flag[kMaxMiners];  // local flag words: miner i published round `round`
  const uint32_t* bad_flag[kMaxMiners];    // local flag words: == round iff miner i's delta holds NaN/Inf (set by its emit)
  int* active;                             // [N] out
  int* n_active;                           // [1] out
  float* w;                                // [N, P], (re-)initialised to 1/n_active on the active rows when init_w
  int* error_flag;
  int N, P, init_w;
  uint32_t round;
};

__global__ void __launch_bounds__(256) round_prepare_kernel(const __grid_constant__ PrepParams p) {
  __shared__ int s_act[kMaxMiners];
  __shared__ int s_n;
  if (threadIdx.x < p.N) {
    bool ok = true;
    if (p.delta_flag[threadIdx.x]) ok = wait_flag_ge(p.delta_flag[threadIdx.x], p.round, p.error_flag);  // timeout == failed download
    const bool bad = p.bad_flag[threadIdx.x] && ld_acquire_sys(p.bad_flag[threadIdx.x]) == p.round;
    s_act[threadIdx.x] = (ok && !bad) ? 1 : 0;
    p.active[threadIdx.x] = s_act[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < p.N; ++i) n += s_act[i];
    s_n = n;
    *p.n_active = n;
  }
  __syncthreads();
  if (p.init_w && p.w) {
    const float v = s_n > 0 ? 1.f / float(s_n) : 0.f;  // softmax(ones[N_active, P], dim=0) (reference :423-430)
    for (int idx = threadIdx.x; idx < p.N * p.P; idx += blockDim.x) p.w[idx] = s_act[idx / p.P] ? v : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// delta all-to-all by pull: dst[i][e - e0] = fp32(delta_i[e]) for e in [e0, e1), for every miner i
// ------------------------------------------------------------------------------------------------------------------
struct TransParams {
  const void* delta[kMaxMiners];    // peer-mapped delta windows
  const float* dscale[kMaxMiners];  // fp8 block scales (mode 2)
  float* dst[kMaxMiners];           // LOCAL shard buffers
  const int* active;
  size_t e0, e1;
  int N, mode;
};

template <int MODE>
__global__ void __launch_bounds__(256) shard_transpose_kernel(const __grid_constant__ TransParams p) {
  const size_t n8 = (p.e1 - p.e0) / 8;
  for (size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x; v < n8; v += size_t(gridDim.x) * blockDim.x) {
    const size_t e = p.e0 + v * 8;
    int i = 0;
    for (; i + 4 <= p.N; i += 4) {  // 4 peers' loads in flight per thread
      float d[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!p.active || p.active[i + k]) load_delta8_plain<MODE>(p.delta[i + k], p.dscale[i + k], e, d[k]);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) d[k][q] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4* o = reinterpret_cast<float4*>(p.dst[i + k] + v * 8);
        o[0] = make_float4(d[k][0], d[k][1], d[k][2], d[k][3]);
        o[1] = make_float4(d[k][4], d[k][5], d[k][6], d[k][7]);
      }
    }
    for (; i < p.N; ++i) {
      float d[8];
      if (!p.active || p.active[i]) load_delta8_plain<MODE>(p.delta[i], p.dscale[i], e, d);
      else {
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = 0.f;
      }
      float4* o = reinterpret_cast<float4*>(p.dst[i] + v * 8);
      o[0] = make_float4(d[0], d[1], d[2], d[3]);
      o[1] = make_float4(d[4], d[5], d[6], d[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// bf16 all-gather by pull: chunk c comes from src[owner(c)]; the caller's own shard is skipped (already local)
// ------------------------------------------------------------------------------------------------------------------
struct Pull16Params {
  const bf16* src[kMaxMiners];            // every rank's bf16 theta_bar window
  const uint32_t* wait_flag[kMaxMiners];  // local flag words: rank r's shard of step wait_value is complete
  const int64_t* chunk_start;
  const int32_t* chunk_len;
  bf16* dst;
  int* error_flag;
  int world, self, num_chunks, chunks_per_rank;
  uint32_t wait_value;
};

__global__ void __launch_bounds__(256) shard_pull16_kernel(const __grid_constant__ Pull16Params p) {
  if (p.wait_value != 0) {
    if (threadIdx.x < p.world && threadIdx.x != p.self && p.wait_flag[threadIdx.x] != nullptr)
      wait_flag_ge(p.wait_flag[threadIdx.x], p.wait_value, p.error_flag);
    __syncthreads();
  }
  const int own0 = p.self * p.chunks_per_rank, own1 = (p.self == p.world - 1) ? p.num_chunks : own0 + p.chunks_per_rank;
  const int n_other = p.num_chunks - (own1 - own0);
  for (int k = blockIdx.x; k < n_other; k += gridDim.x) {
    const int c = k < own0 ? k : k + (own1 - own0);
    const int owner = min(c / p.chunks_per_rank, p.world - 1);
    const bf16* src = p.src[owner];
    const size_t start = size_t(p.chunk_start[c]);
    const int len = p.chunk_len[c];
    for (int v8 = threadIdx.x; v8 * 8 < len; v8 += blockDim.x) {
      const size_t e = start + size_t(v8) * 8;
      *reinterpret_cast<uint4*>(p.dst + e) = *reinterpret_cast<const uint4*>(src + e);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// segmented multi-dot over a chunk range (meta-gradient).  Per chunk c in [c0, c1):
//     partial[c - c0, i] = sum_e gs[e] * delta_i[e]      (i < N),     partial[c - c0, N] = sum_e gs[e] * (base[e] - avg[e])
//     gs[e] = sum_r gscale[r] * g_r[e]                   (R = 1: the local gradient; R = world: fused reduce-scatter of g)
// Register accumulators for 8 miners per pass (N > 8: more passes, g comes from L2), plain 128-bit loads.
// ------------------------------------------------------------------------------------------------------------------
struct SegDotParams {
  const float* g[kMaxMiners];
  float gscale[kMaxMiners];
  const uint32_t* wait_flag[kMaxMiners];  // optional: g_r of step wait_value is complete
  const void* delta[kMaxMiners];          // virtual bases: element e of miner i lives at delta[i] + e (typed by MODE)
  const float* dscale[kMaxMiners];
  const float* base;
  const float* avg;
  const int64_t* chunk_start;
  const int32_t* chunk_len;
  float* partial;
  int* error_flag;
  int N, R, c0, c1;
  uint32_t wait_value;
};

template <int MODE>
__global__ void __launch_bounds__(256) seg_dot_kernel(const __grid_constant__ SegDotParams p) {
  __shared__ float red[8][9];
  if (p.wait_value != 0) {
    if (threadIdx.x < p.R && p.wait_flag[threadIdx.x] != nullptr) wait_flag_ge(p.wait_flag[threadIdx.x], p.wait_value, p.error_flag);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int c = p.c0 + blockIdx.x; c < p.c1; c += gridDim.x) {
    const size_t start = size_t(p.chunk_start[c]);
    const int len = p.chunk_len[c];
    for (int mg = 0; mg < p.N; mg += 8) {
      const int nm = min(8, p.N - mg);
      float acc[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[k] = 0.f;
      for (int v8 = threadIdx.x; v8 * 8 < len; v8 += blockDim.x) {
        const size_t e = start + size_t(v8) * 8;
        float gs[8];
        {
          const float s0 = p.gscale[0];
          const float4 a = *reinterpret_cast<const float4*>(p.g[0] + e), b = *reinterpret_cast<const float4*>(p.g[0] + e + 4);
          gs[0] = s0 * a.x; gs[1] = s0 * a.y; gs[2] = s0 * a.z; gs[3] = s0 * a.w;
          gs[4] = s0 * b.x; gs[5] = s0 * b.y; gs[6] = s0 * b.z; gs[7] = s0 * b.w;
        }
        for (int r = 1; r < p.R; ++r) {
          const float sr = p.gscale[r];
          const float4 a = *reinterpret_cast<const float4*>(p.g[r] + e), b = *reinterpret_cast<const float4*>(p.g[r] + e + 4);
          gs[0] += sr * a.x; gs[1] += sr * a.y; gs[2] += sr * a.z; gs[3] += sr * a.w;
          gs[4] += sr * b.x; gs[5] += sr * b.y; gs[6] += sr * b.z; gs[7] += sr * b.w;
        }
        if (mg == 0) {
          const float4 b0 = *reinterpret_cast<const float4*>(p.base + e), b1 = *reinterpret_cast<const float4*>(p.base + e + 4);
          const float4 a0 = *reinterpret_cast<const float4*>(p.avg + e), a1 = *reinterpret_cast<const float4*>(p.avg + e + 4);
          acc[8] += gs[0] * (b0.x - a0.x) + gs[1] * (b0.y - a0.y) + gs[2] * (b0.z - a0.z) + gs[3] * (b0.w - a0.w) +
                    gs[4] * (b1.x - a1.x) + gs[5] * (b1.y - a1.y) + gs[6] * (b1.z - a1.z) + gs[7] * (b1.w - a1.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < nm) {
            float d[8];
            load_delta8_plain<MODE>(p.delta[mg + k], p.dscale[mg + k], e, d);
            acc[k] += gs[0] * d[0] + gs[1] * d[1] + gs[2] * d[2] + gs[3] * d[3] + gs[4] * d[4] + gs[5] * d[5] + gs[6] * d[6] +
                      gs[7] * d[7];
          }
        }
      }
      // block reduce 9 values (fixed order -> deterministic)
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[wib][k] = v;
      }
      __syncthreads();
      if (threadIdx.x < 9) {
        float s = 0.f;
        for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
        float* row = p.partial + size_t(c - p.c0) * (p.N + 1);
        if (threadIdx.x < nm) row[mg + threadIdx.x] = s;
        else if (threadIdx.x == 8 && mg == 0) row[p.N] = s;
      }
      __syncthreads();
    }
  }
}

// one block per tensor j: G[i, j] = sum over this launch's chunks of tensor j (partial[., i] + partial[., N]); stored to every
// destination table (local and/or peer); entry [N * P] carries this rank's share of the validation loss.
struct SegDotFinParams {
  float* dst[kMaxMiners];
  const float* partial;
  const int32_t* first_chunk;  // [P + 1]
  const int* active;
  const float* loss;
  float loss_scale;
  int N, P, c0, c1, n_dst;
};

__global__ void __launch_bounds__(256) seg_dot_finish_kernel(const __grid_constant__ SegDotFinParams p) {
  __shared__ float red[8][kMaxMiners + 1];
  const int j = blockIdx.x;
  const int a = max(p.first_chunk[j], p.c0), b = min(p.first_chunk[j + 1], p.c1);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int i = 0; i <= p.N; ++i) {
    float s = 0.f;
    for (int c = a + threadIdx.x; c < b; c += blockDim.x) s += p.partial[size_t(c - p.c0) * (p.N + 1) + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[wib][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < p.N) {
    float s = 0.f, cm = 0.f;
    for (int wv = 0; wv < 8; ++wv) {
      s += red[wv][threadIdx.x];
      cm += red[wv][p.N];
    }
    const float v = (!p.active || p.active[threadIdx.x]) ? s + cm : 0.f;
    for (int d = 0; d < p.n_dst; ++d) p.dst[d][size_t(threadIdx.x) * p.P + j] = v;
  }
  if (j == 0 && threadIdx.x == 0 && p.loss) {
    const float l = *p.loss * p.loss_scale;
    for (int d = 0; d < p.n_dst; ++d) p.dst[d][size_t(p.N) * p.P] = l;
  }
}

// w -= lr * sum_r slot_r (fixed summation order: identical on every rank); the summed loss shares are accumulated for logging
struct WUpdParams {
  const float* slot[kMaxMiners];
  const uint32_t* wait_flag[kMaxMiners];
  float* w;
  float* loss_acc;   // [2]: running sum of losses, last loss
  int* error_flag;
  float lr;
  int N, P, R;
  uint32_t wait_value;
};

__global__ void __launch_bounds__(256) w_update_kernel(const __grid_constant__ WUpdParams p) {
  if (p.wait_value != 0) {
    if (threadIdx.x < p.R && p.wait_flag[threadIdx.x] != nullptr) wait_flag_ge(p.wait_flag[threadIdx.x], p.wait_value, p.error_flag);
    __syncthreads();
  }
  const int n = p.N * p.P;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    float g = 0.f;
    for (int r = 0; r < p.R; ++r) g += p.slot[r][idx];
    p.w[idx] -= p.lr * g;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.loss_acc) {
    float l = 0.f;
    for (int r = 0; r < p.R; ++r) l += p.slot[r][n];
    p.loss_acc[0] += l;
    p.loss_acc[1] = l;
  }
}

}  // namespace dtb

using namespace dtb;

#define KCHECK() (cudaGetLastError() == cudaSuccess ? 0 : 1)

extern "C" int dtb_set_flag_timeout_meta(double seconds) {
  const long long polls = seconds <= 0 ? (1ll << 62) : (long long)(seconds / 200e-9);
  return cudaMemcpyToSymbol(g_spin_limit, &polls, sizeof(polls)) == cudaSuccess ? 0 : 1;
}

extern "C" int dtb_round_prepare(const uint32_t** delta_flags, const uint32_t** bad_flags, uint32_t round, int* active,
                                 int* n_active, float* w, int N, int P, int init_w, int* error_flag, cudaStream_t s) {
  if (N > kMaxMiners) return 3;
  PrepParams p{};
  for (int i = 0; i < N; ++i) {
    p.delta_flag[i] = delta_flags ? delta_flags[i] : nullptr;
    p.bad_flag[i] = bad_flags ? bad_flags[i] : nullptr;
  }
  p.active = active; p.n_active = n_active; p.w = w; p.error_flag = error_flag; p.N = N; p.P = P; p.init_w = init_w; p.round = round;
  round_prepare_kernel<<<1, 256, 0, s>>>(p);
  return KCHECK();
}

extern "C" int dtb_shard_transpose(const void** deltas, const float** dscales, float** dsts, const int* active, size_t e0,
                                   size_t e1, int N, int mode, int grid, cudaStream_t s) {
  if (N > kMaxMiners) return 3;
  if (e1 <= e0) return 0;
  TransParams p{};
  for (int i = 0; i < N; ++i) {
    p.delta[i] = deltas[i];
    p.dscale[i] = dscales ? dscales[i] : nullptr;
    p.dst[i] = dsts[i];
  }
  p.active = active; p.e0 = e0; p.e1 = e1; p.N = N; p.mode = mode;
  if (mode == 0) shard_transpose_kernel<0><<<grid, 256, 0, s>>>(p);
  else if (mode == 1) shard_transpose_kernel<1><<<grid, 256, 0, s>>>(p);
  else shard_transpose_kernel<2><<<grid, 256, 0, s>>>(p);
  return KCHECK();
}

extern "C" int dtb_shard_pull16(const void** srcs, const uint32_t** wait_flags, uint32_t wait_value, const int64_t* chunk_start,
                                const int32_t* chunk_len, int num_chunks, int chunks_per_rank, int world, int self, void* dst,
                                int* error_flag, int grid, cudaStream_t s) {
  if (world > kMaxMiners) return 3;
  if (world <= 1) return 0;
  Pull16Params p{};
  for (int r = 0; r < world; ++r) {
    p.src[r] = (const bf16*)srcs[r];
    p.wait_flag[r] = wait_flags ? wait_flags[r] : nullptr;
  }
  p.chunk_start = chunk_start; p.chunk_len = chunk_len; p.dst = (bf16*)dst; p.error_flag = error_flag; p.world = world; p.self = self;
  p.num_chunks = num_chunks; p.chunks_per_rank = chunks_per_rank; p.wait_value = wait_flags ? wait_value : 0;
  if (grid > num_chunks) grid = num_chunks;
  shard_pull16_kernel<<<grid, 256, 0, s>>>(p);
  return KCHECK();
}

// gs: R gradient arenas (+ scales, + optional flags); deltas: N typed virtual bases; writes partial[(c1-c0), N+1] then reduces it
// per tensor into n_dst destination tables [N*P + 1].
extern "C" int dtb_seg_dot(const float** gs, const float* gscales, const uint32_t** wait_flags, uint32_t wait_value, int R,
                           const void** deltas, const float** dscales, int N, int mode, const float* base, const float* avg,
                           const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* first_chunk, int c0, int c1, int P,
                           float* partial, float** dsts, int n_dst, const int* active, const float* loss, float loss_scale,
                           int* error_flag, int grid, cudaStream_t s) {
  if (N > kMaxMiners || R > kMaxMiners || n_dst > kMaxMiners || R < 1) return 3;
  SegDotParams p{};
  for (int r = 0; r < R; ++r) {
    p.g[r] = gs[r];
    p.gscale[r] = gscales ? gscales[r] : 1.f;
    p.wait_flag[r] = wait_flags ? wait_flags[r] : nullptr;
  }
  for (int i = 0; i < N; ++i) {
    p.delta[i] = deltas[i];
    p.dscale[i] = dscales ? dscales[i] : nullptr;
  }
  p.base = base; p.avg = avg; p.chunk_start = chunk_start; p.chunk_len = chunk_len; p.partial = partial; p.error_flag = error_flag;
  p.N = N; p.R = R; p.c0 = c0; p.c1 = c1; p.wait_value = wait_flags ? wait_value : 0;
  if (c1 > c0) {
    int g1 = grid > c1 - c0 ? c1 - c0 : grid;
    if (mode == 0) seg_dot_kernel<0><<<g1, 256, 0, s>>>(p);
    else if (mode == 1) seg_dot_kernel<1><<<g1, 256, 0, s>>>(p);
    else seg_dot_kernel<2><<<g1, 256, 0, s>>>(p);
    if (cudaGetLastError() != cudaSuccess) return 1;
  }
  SegDotFinParams f{};
  for (int d = 0; d < n_dst; ++d) f.dst[d] = dsts[d];
  f.partial = partial; f.first_chunk = first_chunk; f.active = active; f.loss = loss; f.loss_scale = loss_scale;
  f.N = N; f.P = P; f.c0 = c0; f.c1 = c1; f.n_dst = n_dst;
  seg_dot_finish_kernel<<<P, 256, 0, s>>>(f);
  return KCHECK();
}

extern "C" int dtb_w_update(const float** slots, const uint32_t** wait_flags, uint32_t wait_value, int R, float* w, float lr, int N,
                            int P, float* loss_acc, int* error_flag, cudaStream_t s) {
  if (R > kMaxMiners || R < 1) return 3;
  WUpdParams p{};
  for (int r = 0; r < R; ++r) {
    p.slot[r] = slots[r];
    p.wait_flag[r] = wait_flags ? wait_flags[r] : nullptr;
  }
  p.w = w; p.loss_acc = loss_acc; p.error_flag = error_flag; p.lr = lr; p.N = N; p.P = P; p.R = R;
  p.wait_value = wait_flags ? wait_value : 0;
  const int n = N * P;
  w_update_kernel<<<(n + 255) / 256, 256, 0, s>>>(p);
  return KCHECK();
}
