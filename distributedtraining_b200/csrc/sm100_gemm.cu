// Persistent warp-specialised tcgen05 GEMM for sm_100a (bf16 in, fp32 accumulate in TMEM).
//
//   C[M,N] (+)= epilogue( A · B^T )        reduction dimension K
//
// * operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through a 4-stage mbarrier ring
// * one elected thread issues tcgen05.mma (UMMA 128 x 256 x 16, cta_group::1), accumulators live in TMEM,
//   double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of tile i+1
// * epilogue warps read TMEM with tcgen05.ld, apply the fused epilogue (bias / GELU / residual / dGELU),
//   stage through swizzled shared memory and leave via TMA store (bf16) or TMA reduce-add (fp32, split-K wgrad)
// * both operand majors are supported so forward (A,B K-major), dgrad (B MN-major) and wgrad (A,B MN-major)
//   run without any transposes.
//
// Role parity: replaces the cuBLAS GEMMs behind the reference's HF GPT-2 forward/backward
// (reference hivetrain/training_manager.py:380-386).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "dropout.cuh"
#include "sm100_ptx.cuh"

namespace dtb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // one 128B swizzle atom of bf16
constexpr int UMMA_K = 16;
constexpr int kStages = 3;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;       // 1-SM: 48 KB x 3 stages
constexpr int kStages2 = 5;                          // 2-SM: (16 KB A + 16 KB half-B) x 5 stages
constexpr int kStageBytes2 = kABytes + kBBytes / 2;
constexpr int kRingBytes = kStages2 * kStageBytes2 > kStages * kStageBytes ? kStages2 * kStageBytes2 : kStages * kStageBytes;
constexpr int kSlabBytes = BLOCK_M * 128;       // 16 KB: 128 rows x 128 B (64 bf16 or 32 fp32 columns)
constexpr int kNumSlabBufs = 4;                 // two per epilogue group: staging is double-buffered
constexpr int kTmemCols = 512;
constexpr int kNumThreads = 352;                // warp0 TMA, warp1 MMA, warps2-9 epilogue (2 groups of 4), warp10 persist
constexpr int kEpiThreads = 256;
constexpr int kSmemBytes = 1024 /*align slack*/ + kRingBytes + kNumSlabBufs * kSlabBytes + BLOCK_N * 4 /*bias*/ + 256 /*barriers*/;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum Epi : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_BIAS_RESID = 3, EPI_DGELU = 4, EPI_RESID = 5 };

struct GemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  CUtensorMap tmap_c;
  CUtensorMap tmap_c2;  // second output (pre-activation) for EPI_BIAS_GELU
  CUtensorMap tmap_aux; // residual / pre-activation input tile for EPI_*RESID / EPI_DGELU
  CUtensorMap tmap_b2;  // optional second B operand: C = A (B + B2)^T computed as two accumulating MMA passes
  CUtensorMap tmap_bp;  // optional local destination: B tiles pulled from a peer are persisted here as a side effect
  int dual_b, persist_b;
  const float* scale_a;  // optional device scalars: effective alpha = alpha * (*scale_a) * (*scale_b)  (fp8 de-quantisation)
  const float* scale_b;
  int fp8;      // operands are e4m3 bytes (K-major both): 128 K-elements per stage, tcgen05.mma kind::f8f6f4 (K = 32)
  int kblk;     // K elements per pipeline stage: 64 (bf16) or 128 (fp8) -- one 128-byte swizzle atom either way
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, num_kb;
  int epi;
  int ldaux;
  const __nv_bfloat16* bias;  // [N]
  const __nv_bfloat16* aux;   // [M, ldaux]
  __nv_bfloat16* c2;          // second output (pre-activation), written with direct 16 B stores in the dual epilogue
  int ldc2;
  float alpha;
  DropArgs drop;  // EPI_BIAS_RESID / EPI_RESID: out = aux + dropout(acc [+ bias])  (GPT-2 resid_pdrop)
  float* colsum;  // optional fp32 [N]: += column sums of the bf16 output (e.g. the bias gradient that equals colsum(dY))
  // Fused broadcast -> first forward GEMM (path (b)): the B operand (a weight matrix of the NEW averaged base) is being landed
  // in local HBM by the shard owners' averaging kernels (multimem.st through the NVSwitch).  The kernel itself acquires the
  // owners' base flags (slots 0..ready_hi of this rank's flag page) against the device-resident target round before any
  // thread touches the weights -- no separate wait kernel, no host involvement; nullptr = plain GEMM.
  const uint32_t* ready_flags;
  const uint32_t* ready_target;
  int ready_hi;
};

// GELU (tanh form, HF "gelu_new") with MUFU.TANH in fp32.  (A packed tanh.approx.bf16x2 variant halves the MUFU count but
// its bf16-precision tanh is not accurate enough for the derivative: (1 - t^2) amplifies the error -- measured, reverted.)
DTB_DEVICE float tanh_fast(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return t;
}
// Pair forms on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2: one issue slot per TWO elements).  The fused epilogues are
// issue-bound, not latency-bound: with scalar math the GELU epilogue needed ~3.3 us of issue time per 128x256 tile against
// 3.2 us of tensor-core time for K = 768, so the fc / dgelu GEMMs ran at 0.9 PF while bias-only epilogues reached 1.3 PF.
DTB_DEVICE float2 f2(float a, float b) { return make_float2(a, b); }
DTB_DEVICE float2 gelu_tanh2(float2 x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float2 a = __fmul2_rn(x, __ffma2_rn(f2(k01, k01), __fmul2_rn(x, x), f2(k0, k0)));
  const float2 t = f2(tanh_fast(a.x), tanh_fast(a.y));
  const float2 h = __fmul2_rn(x, f2(0.5f, 0.5f));
  return __ffma2_rn(h, t, h);
}
// d/dx gelu_tanh for a pair: 0.5(1+t) + 0.5 x (1-t^2) k0 (1 + 3 k1 x^2)
DTB_DEVICE float2 dgelu_tanh2(float2 x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f, k3 = 3.f * 0.7978845608028654f * 0.044715f;
  const float2 s = __fmul2_rn(x, x);
  const float2 a = __fmul2_rn(x, __ffma2_rn(f2(k01, k01), s, f2(k0, k0)));
  const float2 t = f2(tanh_fast(a.x), tanh_fast(a.y));
  const float2 w = __fmul2_rn(__fmul2_rn(x, f2(0.5f, 0.5f)), __ffma2_rn(f2(-t.x, -t.y), t, f2(1.f, 1.f)));
  return __ffma2_rn(w, __ffma2_rn(f2(k3, k3), s, f2(k0, k0)), __ffma2_rn(f2(0.5f, 0.5f), t, f2(0.5f, 0.5f)));
}
DTB_DEVICE uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// explicit shared-memory load: through a generic pointer the bias reads compiled to LD.E.128 (generic), on which the epilogue
// warps of the bias+GELU GEMM spent 13.5 % of their stall samples (profiles/ncu_r2_gemm_bias_gelu.json)
DTB_DEVICE float4 lds_f4(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
DTB_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// CL = 1: one CTA per 128x256 tile (cta_group::1).  Measured ceiling ~1.3 PF: TMA fills (96 B/clk) + tensor-core operand
//         reads (96 B/clk) exceed the 128 B/clk shared-memory bandwidth of one SM.
// CL = 2: 2-SM UMMA (cta_group::2): a CTA pair owns a 256x256 tile; each CTA stages its 128 A rows and HALF of the B tile
//         (the tensor cores of the pair share operands), so per-SM shared-memory traffic drops to 64 + 64 B/clk and the
//         ring deepens to 6 stages.  Only the pair's leader issues MMAs; TMA bytes of both CTAs are credited to the
//         leader's full barrier; stage release / accumulator-ready are multicast commits; the non-leader's epilogue
//         arrives remotely on the leader's TMEM-empty barrier.
// FP8: e4m3 operands (K-major only), kind::f8f6f4.  A TEMPLATE parameter on purpose: a runtime branch in the single-thread
// MMA issue loop cost 20 % of the GEMM throughput (measured: 1088 -> 860 TFLOP/s on 16384x2304x768) -- the issuing thread has
// ~128 cycles per instruction and every extra branch/select in that loop starves the tensor pipe.
// EPIT >= 0: the epilogue mode is a compile-time constant (the hot shapes of the training step); -1: read it from the params.
// The slab loop of the runtime version spends ~6 % of its instructions on mode branches / selects, and the fused epilogues are
// issue-bound.
template <bool A_MN, bool B_MN, bool OUT_F32, int CL, int FP8 = 0, int EPIT = -1>
__global__ void __launch_bounds__(kNumThreads, 1) sm100_gemm_kernel(const __grid_constant__ GemmParams p) {
  const int epi_mode = EPIT >= 0 ? EPIT : p.epi;
  pdl_launch_dependents();  // the NEXT kernel of the stream may begin its own prologue as soon as every CTA is past this point
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kNStages = CL == 2 ? kStages2 : kStages;
  constexpr int kStgBytes = CL == 2 ? kStageBytes2 : kStageBytes;
  uint8_t* smem_ab = smem;
  uint8_t* smem_slab = smem + kRingBytes;
  float* smem_bias = reinterpret_cast<float*>(smem_slab + kNumSlabBufs * kSlabBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_bias + BLOCK_N);
  uint64_t* empty_bar = full_bar + kNStages;
  uint64_t* tmem_full_bar = empty_bar + kNStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* aux_bars = tmem_empty_bar + 2;  // [4] one per epilogue group and staging buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aux_bars + 4);

  const uint32_t warp_idx = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    tma_prefetch_desc(&p.tmap_c);
    tma_prefetch_desc(&p.tmap_aux);
    for (int i = 0; i < kNStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], (CL == 1 && p.persist_b) ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], CL * (kEpiThreads / 32));  // 2-SM: the epilogue warps of BOTH CTAs arrive on the leader
      mbar_init(&aux_bars[2 * i], 1);
      mbar_init(&aux_bars[2 * i + 1], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    if constexpr (CL == 2) {
      tmem_alloc_2sm(tmem_ptr_smem, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, kTmemCols);
      tmem_relinquish();
    }
  }
  // everything above touched only shared memory, TMEM and kernel parameters; from here on global memory written by earlier
  // kernels is read (flags, bias, operands) or overwritten (outputs): wait for the predecessor grid to complete and flush
  pdl_wait();
  if (p.ready_flags != nullptr && threadIdx.x == 64) {  // an epilogue thread: warps 0 / 1 are busy with barriers and TMEM
    const uint32_t tgt = *reinterpret_cast<const volatile uint32_t*>(p.ready_target);
    for (int o = 0; o <= p.ready_hi; ++o) {
      long long spins = 0;
      while (ld_acquire_sys(p.ready_flags + o) < tgt && ++spins < (1ll << 27)) __nanosleep(100);
    }
    fence_proxy_async_all();  // the weights are consumed by TMA (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // peers must not multicast into / arrive on uninitialised barriers
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  const int tiles_mg = (p.tiles_m + CL - 1) / CL;  // M-tile groups; CTA r of a cluster owns m_t = mg * CL + r
  const int total_work = tiles_mg * p.tiles_n * p.splits;
  constexpr uint16_t kMcMask = uint16_t((1u << CL) - 1);
  const bool is_leader = cta_rank == 0;

  if (warp_idx == 0) {
    // ===================== TMA producer (warp-uniform loop; only the issue itself is under elect.sync) =====================
    {
      uint32_t stage = 0, phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int n_t = w % p.tiles_n;
        const int m_t = ((w / p.tiles_n) % tiles_mg) * CL + cta_rank;
        const int sp = w / (p.tiles_n * tiles_mg);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int m0 = m_t * BLOCK_M, n0 = n_t * BLOCK_N;
        const int kreps = p.dual_b ? 2 : 1;
        for (int rep = 0; rep < kreps; ++rep) {
          const CUtensorMap* bmap = rep == 0 ? &p.tmap_b : &p.tmap_b2;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem_ab + stage * kStgBytes;
            uint8_t* sb = sa + kABytes;
            const int k0 = kb * (FP8 ? 128 : BLOCK_K);
            if (elect_one()) {
            if constexpr (CL == 2) {
              // both CTAs load (own A rows + own half of B); all bytes are credited to the leader's full barrier
              if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes2);
              if constexpr (A_MN) {
#pragma unroll
                for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d_2sm(sa + a * (BLOCK_K * 128), &p.tmap_a, &full_bar[stage], m0 + a * 64, k0);
              } else {
                tma_load_2d_2sm(sa, &p.tmap_a, &full_bar[stage], k0, m0);
              }
              const int nh = n0 + cta_rank * (BLOCK_N / 2);
              if constexpr (B_MN) {
#pragma unroll
                for (int a = 0; a < BLOCK_N / 128; ++a) tma_load_2d_2sm(sb + a * (BLOCK_K * 128), bmap, &full_bar[stage], nh + a * 64, k0);
              } else {
                tma_load_2d_2sm(sb, bmap, &full_bar[stage], k0, nh);
              }
            } else {
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              if constexpr (A_MN) {
#pragma unroll
                for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d(sa + a * (BLOCK_K * 128), &p.tmap_a, &full_bar[stage], m0 + a * 64, k0);
              } else {
                tma_load_2d(sa, &p.tmap_a, &full_bar[stage], k0, m0);
              }
              if constexpr (B_MN) {
#pragma unroll
                for (int a = 0; a < BLOCK_N / 64; ++a) tma_load_2d(sb + a * (BLOCK_K * 128), bmap, &full_bar[stage], n0 + a * 64, k0);
              } else {
                tma_load_2d(sb, bmap, &full_bar[stage], k0, n0);
              }
            }
            }  // elect_one
            if (++stage == kNStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // FP8 = 1: e4m3 x e4m3 (forward);  FP8 = 2: A in e5m2 (gradients: range over precision), B in e4m3 (weights) -- dgrad
    constexpr uint32_t idesc = FP8 ? make_idesc(FP8 == 2 ? kFmtE5M2 : kFmtE4M3, kFmtE4M3, false, false, BLOCK_M * CL, BLOCK_N)
                                   : make_idesc(kFmtBF16, kFmtBF16, A_MN, B_MN, BLOCK_M * CL, BLOCK_N);
    // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused (1).  MN-major SW128: 64-element atoms along MN are
    // BLOCK_K*128 B apart (LBO), 8-row K groups 1024 B apart (SBO).
    constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 16, b_lbo = B_MN ? BLOCK_K * 128 : 16;
    constexpr uint32_t a_kadv = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;  // descriptor start-address units (16 B)
    constexpr uint32_t b_kadv = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int w = cluster_id; is_leader && w < total_work; w += num_clusters) {  // 2-SM: only the leader issues
      const int sp = w / (p.tiles_n * tiles_mg);
      const int kb0 = sp * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      const int nk = (kb1 - kb0) * (p.dual_b ? 2 : 1);
      for (int ki = 0; ki < nk; ++ki) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // The whole warp runs this loop convergently and only the tcgen05 instructions sit under elect.sync: every operand is
        // then provably warp-uniform and ptxas keeps the descriptors in uniform registers.  (Issuing under a divergent
        // `if (lane == 0)` made ptxas wrap EVERY UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop; the single
        // issuing thread has ~128 cycles per instruction, so that overhead directly starved the tensor pipe.)
        const uint32_t sa = smem_u32(smem_ab + stage * kStgBytes);
        const uint32_t sb = sa + kABytes;
        const uint64_t da = make_smem_desc(sa, a_lbo, 1024);
        const uint64_t db = make_smem_desc(sb, b_lbo, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t accf = (ki > 0 || k > 0) ? 1u : 0u;
            const uint64_t dak = da + uint64_t(k * a_kadv), dbk = db + uint64_t(k * b_kadv);
            if constexpr (CL == 2) {
              if constexpr (FP8) umma_f8_2sm(tmem_d, dak, dbk, idesc, accf); else umma_f16_2sm(tmem_d, dak, dbk, idesc, accf);
            } else {
              if constexpr (FP8) umma_f8(tmem_d, dak, dbk, idesc, accf); else umma_f16(tmem_d, dak, dbk, idesc, accf);
            }
          }
          if constexpr (CL == 2) {
            umma_commit_2sm_mc(&empty_bar[stage], kMcMask);  // releases the stage in BOTH CTAs
            if (ki == nk - 1) umma_commit_2sm_mc(&tmem_full_bar[acc], kMcMask);
          } else {
            umma_commit(&empty_bar[stage]);
            if (ki == nk - 1) umma_commit(&tmem_full_bar[acc]);
          }
        }
        if (++stage == kNStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp_idx == 10) {
    // ===================== persist warp (fused broadcast -> GEMM) =====================
    // When B is pulled from a PEER window, the CTA that owns the first M-tile of each N-tile TMA-stores every B stage
    // to the local weight copy while the MMA consumes it: the broadcast rides on the first forward GEMM.
    if (CL == 1 && p.persist_b && lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int n_t = w % p.tiles_n;
        const int m_t = ((w / p.tiles_n) % tiles_mg) * CL + cta_rank;
        const int sp = w / (p.tiles_n * tiles_mg);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int n0 = n_t * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (m_t == 0) {
            const uint8_t* sb = smem_ab + stage * kStageBytes + kABytes;
            const int k0 = kb * BLOCK_K;
            if constexpr (B_MN) {
#pragma unroll
              for (int a = 0; a < BLOCK_N / 64; ++a) tma_store_2d(&p.tmap_bp, sb + a * (BLOCK_K * 128), n0 + a * 64, k0);
            } else {
              tma_store_2d(&p.tmap_bp, sb, k0, n0);
            }
            tma_store_commit();
            tma_store_wait_read<0>();
          }
          mbar_arrive(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ===================== epilogue (warps 2..9): two independent groups of 4 warps =====================
    // Group g handles slabs g, g+2, ... of every tile with its own PAIR of staging buffers, named barrier, aux mbarriers
    // and TMA issuer thread.  Two warps per SM sub-partition hide the ALU/MUFU latency of the fused epilogues.  Work items
    // (tile, slab) alternate between the two buffers, so the TMA store of item i drains while item i+1 is computed, and the
    // aux tile (residual / pre-activation) of item i+1 is TMA-loaded into the other buffer while item i is processed (the
    // first slab of the NEXT tile is prefetched across the tile boundary) and transformed in place.  The measured cost of
    // the single-buffered version was one exposed store-drain + one exposed aux load latency per slab (short-K GEMMs with
    // a residual epilogue ran at 0.6 PF).
    const uint32_t ewarp = warp_idx - 2;             // 0..7
    const uint32_t grp = ewarp >> 2;                 // 0 / 1
    const uint32_t quad = warp_idx & 3;              // TMEM lane quadrant this warp may access
    const uint32_t row_l = quad * 32 + lane;         // row within the tile
    const uint32_t epi_tid = threadIdx.x - 64;       // 0..255
    const bool issuer = (lane == 0) && ((ewarp & 3) == 0);  // one issuer thread per group
    const uint32_t bar_id = 1 + grp;                 // named barriers 1 / 2 (group-local, 128 threads)
    uint8_t* const buf0 = smem_slab + (2 * grp) * kSlabBytes;
    uint8_t* const buf1 = buf0 + kSlabBytes;
    uint64_t* const aux_bar = aux_bars + 2 * grp;    // [2], one per buffer
    uint32_t acc = 0, acc_phase = 0, aux_phase = 0 /* bit b = parity of aux_bar[b] */, item = 0;
    const float alpha_eff = p.alpha * (p.scale_a ? (*p.scale_a) * (*p.scale_b) : 1.f);
    constexpr int kColsPerSlab = OUT_F32 ? 32 : 64;
    constexpr int kSlabs = BLOCK_N / kColsPerSlab;
    const bool has_bias = (epi_mode == EPI_BIAS || epi_mode == EPI_BIAS_GELU || epi_mode == EPI_BIAS_RESID);
    const bool has_aux = !OUT_F32 && (epi_mode == EPI_BIAS_RESID || epi_mode == EPI_RESID || epi_mode == EPI_DGELU);
    const bool dual = (epi_mode == EPI_BIAS_GELU);
    const uint32_t dthr = (epi_mode == EPI_BIAS_RESID || epi_mode == EPI_RESID) ? p.drop.thr : 0u;
    const uint32_t dkey = dthr ? drop_key(p.drop.rng, p.drop.stream) : 0u;
    // work item w -> (n tile, m-tile group) WITHOUT a division per tile: the runtime divisions (I2F.RP sequences) sat at the head of
    // every tile's epilogue (5.6 % of the stall samples); the indices advance by a fixed (quotient, remainder) step instead
    const int step_q = num_clusters / p.tiles_n, step_n = num_clusters % p.tiles_n;
    int cur_q = cluster_id / p.tiles_n, cur_n = cluster_id % p.tiles_n;  // q = w / tiles_n
    auto mt_of = [&](int q) { return (OUT_F32 ? (q % tiles_mg) : q) * CL + int(cta_rank); };  // bf16 outputs: splits == 1, q < tiles_mg
    auto issue_aux = [&](int n_t, int m_t, int sl, uint32_t b) {  // issuer only: aux slab (tile, sl) -> staging buffer b
      mbar_expect_tx(&aux_bar[b], kSlabBytes);
      tma_load_2d(b ? buf1 : buf0, &p.tmap_aux, &aux_bar[b], n_t * BLOCK_N + sl * kColsPerSlab, m_t * BLOCK_M);
    };
    if (has_aux && issuer && cluster_id < total_work) issue_aux(cur_n, mt_of(cur_q), int(grp), 0);
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int n_t = cur_n;
      const int m_t = mt_of(cur_q);
      int nxt_n = cur_n + step_n, nxt_q = cur_q + step_q;  // indices of this cluster's next work item
      if (nxt_n >= p.tiles_n) { nxt_n -= p.tiles_n; ++nxt_q; }
      cur_n = nxt_n; cur_q = nxt_q;
      const int m0 = m_t * BLOCK_M, n0 = n_t * BLOCK_N;
      const int row = m0 + int(row_l);
      if (has_bias) {
        named_bar_sync(3, kEpiThreads);  // previous tile's readers of smem_bias are done
        for (int i = epi_tid; i < BLOCK_N; i += kEpiThreads) {
          int c = n0 + i;
          smem_bias[i] = (c < p.N) ? __bfloat162float(p.bias[c]) : 0.f;
        }
        named_bar_sync(3, kEpiThreads);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + acc * BLOCK_N + ((quad * 32u) << 16);
#pragma unroll 1
      for (int s = grp; s < kSlabs; s += 2, ++item) {
        const int c0 = s * kColsPerSlab;  // column offset inside the tile
        const int gcol0 = n0 + c0;
        const bool last = (s + 2 >= kSlabs);
        const uint32_t b = dual ? 0u : (item & 1u);
        uint8_t* const buf = b ? buf1 : buf0;
        uint8_t* rowp = buf + row_l * 128;
        if constexpr (OUT_F32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr_row + c0, r);
          tmem_ld_wait();
          if (last) {  // this warp has drained its share of the accumulator -> hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CL == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
          }
          if (issuer) tma_store_wait_read<1>();  // the store that used this buffer two items ago has drained
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float4 v;
            v.x = __uint_as_float(r[ch * 4 + 0]) * alpha_eff;
            v.y = __uint_as_float(r[ch * 4 + 1]) * alpha_eff;
            v.z = __uint_as_float(r[ch * 4 + 2]) * alpha_eff;
            v.w = __uint_as_float(r[ch * 4 + 3]) * alpha_eff;
            *reinterpret_cast<float4*>(rowp + ((ch ^ (row_l & 7)) << 4)) = v;
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            if (gcol0 < p.N) tma_reduce_add_2d(&p.tmap_c, buf, gcol0, m0);
            tma_store_commit();
          }
        } else {
          uint32_t r[64];
          tmem_ld_32x32b_x32(taddr_row + c0, r);
          tmem_ld_32x32b_x32(taddr_row + c0 + 32, r + 32);
          tmem_ld_wait();
          if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CL == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
          }
          if (has_aux) {
            if (issuer) {  // prefetch the aux slab of the NEXT work item into the other buffer
              int nw = w, ns = s + 2, an = n_t, am = m_t;
              if (ns >= kSlabs) { nw = w + num_clusters; ns = int(grp); an = nxt_n; am = mt_of(nxt_q); }
              if (nw < total_work) {
                tma_store_wait_read<0>();  // the previous item's store (issued one TMEM read ago) has left that buffer
                issue_aux(an, am, ns, b ^ 1u);
              }
            }
            mbar_wait(&aux_bar[b], (aux_phase >> b) & 1u);
            aux_phase ^= (1u << b);
          } else {
            if (issuer) { if (dual) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
            named_bar_sync(bar_id, 128);
          }
          uint8_t* rowp2 = buf1 + row_l * 128;  // dual: gelu(u) tile
          const float2 alpha2 = f2(alpha_eff, alpha_eff);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float2 v[4];  // 8 consecutive columns as 4 packed pairs
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = f2(__uint_as_float(r[ch * 8 + 2 * i]), __uint_as_float(r[ch * 8 + 2 * i + 1]));
            if (has_bias) {
              const float4 b0 = lds_f4(smem_bias + c0 + ch * 8);
              const float4 b1 = lds_f4(smem_bias + c0 + ch * 8 + 4);
              v[0] = __ffma2_rn(v[0], alpha2, f2(b0.x, b0.y));
              v[1] = __ffma2_rn(v[1], alpha2, f2(b0.z, b0.w));
              v[2] = __ffma2_rn(v[2], alpha2, f2(b1.x, b1.y));
              v[3] = __ffma2_rn(v[3], alpha2, f2(b1.z, b1.w));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = __fmul2_rn(v[i], alpha2);
            }
            const uint32_t sw = ((ch ^ (row_l & 7)) << 4);
            if (has_aux) {
              const uint4 q = *reinterpret_cast<const uint4*>(rowp + sw);
              const float2 a[4] = {unpack_bf16x2(q.x), unpack_bf16x2(q.y), unpack_bf16x2(q.z), unpack_bf16x2(q.w)};
              if (epi_mode == EPI_DGELU) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = __fmul2_rn(v[i], dgelu_tanh2(a[i]));
              } else {
                if (dthr) {  // residual-branch dropout: pair index = row * (N / 2) + col / 2 (dropout.cuh)
                  const uint32_t pair0 = uint32_t(row) * uint32_t(p.N >> 1) + (uint32_t(gcol0 + ch * 8) >> 1);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint32_t wd = drop_word(dkey, pair0 + i);
                    v[i] = __fmul2_rn(v[i], f2(drop_mul_lo(wd, dthr, p.drop.scale), drop_mul_hi(wd, dthr, p.drop.scale)));
                  }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = __fadd2_rn(v[i], a[i]);
              }
            }
            uint4 o;
            o.x = pack_bf16x2(v[0].x, v[0].y); o.y = pack_bf16x2(v[1].x, v[1].y);
            o.z = pack_bf16x2(v[2].x, v[2].y); o.w = pack_bf16x2(v[3].x, v[3].y);
            *reinterpret_cast<uint4*>(rowp + sw) = o;
            if (dual) {  // second output gelu(u) goes to the group's other buffer in the same pass
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = gelu_tanh2(v[i]);
              o.x = pack_bf16x2(v[0].x, v[0].y); o.y = pack_bf16x2(v[1].x, v[1].y);
              o.z = pack_bf16x2(v[2].x, v[2].y); o.w = pack_bf16x2(v[3].x, v[3].y);
              *reinterpret_cast<uint4*>(rowp2 + sw) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            if (gcol0 < p.N) {
              tma_store_2d(dual ? &p.tmap_c2 : &p.tmap_c, buf, gcol0, m0);
              if (dual) tma_store_2d(&p.tmap_c, buf1, gcol0, m0);
            }
            tma_store_commit();
          }
          if (p.colsum) {
            // column sums of the staged [128 x 64] output slab (the bf16 values being stored): thread t sums column t % 64
            // over rows (t / 64) * 64 .. + 63 and adds it to colsum[gcol] -- replaces a separate pass over the output.
            // The buffer is rewritten two items later, behind a named barrier every thread passes after these reads.
            const uint32_t gt = epi_tid & 127u;
            const int c = int(gt & 63u), r0 = int(gt >> 6) * 64;
            const uint32_t chunk = uint32_t(c) >> 3, within = (uint32_t(c) & 7u) * 2u;
            float accs = 0.f;
#pragma unroll 8
            for (int r = r0; r < r0 + 64; ++r) {
              if (m0 + r < p.M)
                accs += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(buf + r * 128 + ((chunk ^ (uint32_t(r) & 7u)) << 4) + within));
            }
            if (gcol0 + c < p.N) atomicAdd(p.colsum + gcol0 + c, accs);
            // aux mode: the issuer TMA-loads the aux slab of item i+2 into THIS buffer right after the next TMEM read, with
            // no group barrier in between -> every thread must be done reading it first
            if (has_aux) named_bar_sync(bar_id, 128);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into its shared memory
  if (warp_idx == 1) {
    tc_fence_after();
    if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D row-major tensor [outer, inner] with `pitch` elements between rows; box = [box_outer, box_inner]; 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                 uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * uint64_t(elem_bytes)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : int(r);
}

template <bool A_MN, bool B_MN, bool OUT_F32, int CL, int FP8 = 0, int EPIT = -1>
static cudaError_t launch(const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = sm100_gemm_kernel<A_MN, B_MN, OUT_F32, CL, FP8, EPIT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  static const bool pdl = getenv("DTB200_NO_PDL") == nullptr;  // A/B switch
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int CL>
static cudaError_t dispatch(const GemmParams& p, int grid, bool a_mn, bool b_mn, bool out_f32, cudaStream_t stream) {
  if (p.fp8 == 2) return launch<false, false, false, CL, 2>(p, grid, stream);
  if (p.fp8) return launch<false, false, false, CL, 1>(p, grid, stream);
  if (out_f32) {
    if (a_mn && b_mn) return launch<true, true, true, CL>(p, grid, stream);
    if (!a_mn && b_mn) return launch<false, true, true, CL>(p, grid, stream);
    if (!a_mn && !b_mn) return launch<false, false, true, CL>(p, grid, stream);
    return launch<true, false, true, CL>(p, grid, stream);
  }
  static const bool no_spec = getenv("DTB200_GEMM_NO_EPI_SPEC") != nullptr;  // A/B switch
  if (CL == 2 && !a_mn && !p.colsum && !p.dual_b && !p.persist_b && !no_spec) {  // the training step's hot fused-epilogue shapes
    if (!b_mn && p.epi == EPI_BIAS) return launch<false, false, false, CL, false, EPI_BIAS>(p, grid, stream);
    if (!b_mn && p.epi == EPI_BIAS_GELU) return launch<false, false, false, CL, 0, EPI_BIAS_GELU>(p, grid, stream);
    if (!b_mn && p.epi == EPI_BIAS_RESID) return launch<false, false, false, CL, 0, EPI_BIAS_RESID>(p, grid, stream);
    if (b_mn && p.epi == EPI_DGELU) return launch<false, true, false, CL, 0, EPI_DGELU>(p, grid, stream);
    if (b_mn && p.epi == EPI_NONE) return launch<false, true, false, CL, 0, EPI_NONE>(p, grid, stream);
  }
  if (a_mn && b_mn) return launch<true, true, false, CL>(p, grid, stream);
  if (!a_mn && b_mn) return launch<false, true, false, CL>(p, grid, stream);
  if (!a_mn && !b_mn) return launch<false, false, false, CL>(p, grid, stream);
  return launch<true, false, false, CL>(p, grid, stream);
}

}  // namespace dtb

// C ABI.  A: K-major => [M, K] pitch lda; MN-major => [K, M] pitch lda.  B likewise with N.  C: [M, N] pitch ldc.
// out_f32 => C is fp32 and the tile is reduce-ADDED into it (caller zeroes C); splits > 1 requires out_f32.
static int g_fp8_a_e5m2 = 0;                       // one-shot: the next fp8 GEMM reads its A operand as e5m2 (dgrad)
static const uint32_t* g_ready_flags = nullptr;   // one-shot, see dtb_gemm_set_ready
static const uint32_t* g_ready_target = nullptr;
static int g_ready_hi = 0;

static int gemm_impl(int esz, const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                     int b_mn, int out_f32, int epi, const void* bias, const void* aux, int ldaux, void* c2, int ldc2,
                     float alpha, int splits, int num_sms, cudaStream_t stream, const void* b2, int ldb2, void* b_persist,
                     int ldbp, const float* scale_a, const float* scale_b, const void* rng, int drop_stream, float drop_p,
                     float* colsum) {
  using namespace dtb;
  const int KBLK = 128 / esz;  // K elements per 128-byte swizzle atom
  if (esz == 1 && (a_mn || b_mn || b2 || b_persist)) return 2002;  // fp8 path: K-major operands only
  GemmParams p;
  int rc = 0;
  const int tiles_m_ = (M + BLOCK_M - 1) / BLOCK_M, tiles_n_ = (N + BLOCK_N - 1) / BLOCK_N;
  static const int force_cl = getenv("DTB200_GEMM_CLUSTER") ? atoi(getenv("DTB200_GEMM_CLUSTER")) : 0;
  int CL = (!b2 && !b_persist && tiles_m_ >= 2 && tiles_m_ * tiles_n_ >= 32) ? 2 : 1;
  if (force_cl == 1 || b2 || b_persist) CL = 1;
  if (a_mn) rc |= make_tmap_2d(&p.tmap_a, a, 2, M, K, lda, 64, BLOCK_K);
  else      rc |= make_tmap_2d(&p.tmap_a, a, esz, K, M, lda, KBLK, BLOCK_M);
  if (b_mn) rc |= make_tmap_2d(&p.tmap_b, b, 2, N, K, ldb, 64, BLOCK_K);
  else      rc |= make_tmap_2d(&p.tmap_b, b, esz, K, N, ldb, KBLK, BLOCK_N / CL);
  if (out_f32) rc |= make_tmap_2d(&p.tmap_c, c, 4, N, M, ldc, 32, BLOCK_M);
  else         rc |= make_tmap_2d(&p.tmap_c, c, 2, N, M, ldc, 64, BLOCK_M);
  if (c2) rc |= make_tmap_2d(&p.tmap_c2, c2, 2, N, M, ldc2, 64, BLOCK_M);
  else    p.tmap_c2 = p.tmap_c;
  if (aux) rc |= make_tmap_2d(&p.tmap_aux, aux, 2, N, M, ldaux, 64, BLOCK_M);
  else     p.tmap_aux = p.tmap_c;
  p.dual_b = b2 != nullptr;
  p.persist_b = b_persist != nullptr;
  p.tmap_b2 = p.tmap_b;
  p.tmap_bp = p.tmap_b;
  if (b2) {
    if (b_mn) rc |= make_tmap_2d(&p.tmap_b2, b2, 2, N, K, ldb2, 64, BLOCK_K);
    else      rc |= make_tmap_2d(&p.tmap_b2, b2, 2, K, N, ldb2, BLOCK_K, BLOCK_N);
  }
  if (b_persist) {
    if (b_mn) rc |= make_tmap_2d(&p.tmap_bp, b_persist, 2, N, K, ldbp, 64, BLOCK_K);
    else      rc |= make_tmap_2d(&p.tmap_bp, b_persist, 2, K, N, ldbp, BLOCK_K, BLOCK_N);
  }
  if (p.dual_b && p.persist_b) return 2001;  // one or the other
  if (rc) return 1000 + rc;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  p.num_kb = (K + KBLK - 1) / KBLK;
  p.kblk = KBLK;
  p.fp8 = esz == 1 ? (g_fp8_a_e5m2 ? 2 : 1) : 0;
  g_fp8_a_e5m2 = 0;
  p.c2 = reinterpret_cast<__nv_bfloat16*>(c2);
  p.ldc2 = ldc2;
  p.scale_a = scale_a;
  p.scale_b = scale_b;
  if (splits < 1) splits = 1;
  if (!out_f32) splits = 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  p.splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.epi = epi;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.aux = reinterpret_cast<const __nv_bfloat16*>(aux);
  p.ldaux = ldaux;
  p.alpha = alpha;
  p.drop.rng = reinterpret_cast<const uint32_t*>(rng);
  p.drop.stream = uint32_t(drop_stream);
  p.drop.thr = (rng && drop_p > 0.f) ? uint32_t(drop_p * 65536.f + 0.5f) : 0u;
  p.drop.scale = 1.f / (1.f - drop_p);
  if (p.drop.thr && ((N & 1) || !(epi == EPI_BIAS_RESID || epi == EPI_RESID))) return 2003;
  p.colsum = colsum;
  p.ready_flags = g_ready_flags; p.ready_target = g_ready_target; p.ready_hi = g_ready_hi;
  g_ready_flags = nullptr; g_ready_target = nullptr;
  if (colsum && (out_f32 || epi == EPI_BIAS_GELU)) return 2004;
  const int tiles_mg = (p.tiles_m + CL - 1) / CL;
  int total = tiles_mg * p.tiles_n * p.splits;   // work items per cluster
  int max_clusters = num_sms / CL;
  int nclusters = total < max_clusters ? total : max_clusters;
  if (nclusters < 1) return 0;
  int grid = nclusters * CL;
  cudaError_t e = CL == 2 ? dispatch<2>(p, grid, a_mn, b_mn, out_f32, stream) : dispatch<1>(p, grid, a_mn, b_mn, out_f32, stream);
  return e == cudaSuccess ? 0 : int(e);
}

// C ABI.  bf16 operands (all majors / modes) and e4m3 operands (K-major; alpha carries the product of the tensor scales).
// one-shot: the NEXT GEMM launch acquires these flags in-kernel (see GemmParams::ready_flags)
extern "C" int dtb_gemm_set_ready(const void* flags, const void* target, int hi) {
  g_ready_flags = (const uint32_t*)flags; g_ready_target = (const uint32_t*)target; g_ready_hi = hi;
  return 0;
}
extern "C" int dtb_gemm_bf16(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                             int b_mn, int out_f32, int epi, const void* bias, const void* aux, int ldaux, void* c2, int ldc2,
                             float alpha, int splits, int num_sms, cudaStream_t stream, const void* b2, int ldb2,
                             void* b_persist, int ldbp, const void* rng, int drop_stream, float drop_p, float* colsum) {
  return gemm_impl(2, a, b, c, M, N, K, lda, ldb, ldc, a_mn, b_mn, out_f32, epi, bias, aux, ldaux, c2, ldc2, alpha, splits, num_sms,
                   stream, b2, ldb2, b_persist, ldbp, nullptr, nullptr, rng, drop_stream, drop_p, colsum);
}
extern "C" int dtb_gemm_fp8_a_e5m2() {  // call right before dtb_gemm_fp8: A (an activation GRADIENT) is e5m2, B (a weight) e4m3
  g_fp8_a_e5m2 = 1;
  return 0;
}
extern "C" int dtb_gemm_fp8(const void* a, const void* b, void* c, int M, int N, int K, int lda, int ldb, int ldc, int epi,
                            const void* bias, const void* aux, int ldaux, void* c2, int ldc2, float alpha, int num_sms,
                            cudaStream_t stream, const float* scale_a, const float* scale_b, const void* rng, int drop_stream,
                            float drop_p) {
  return gemm_impl(1, a, b, c, M, N, K, lda, ldb, ldc, 0, 0, 0, epi, bias, aux, ldaux, c2, ldc2, alpha, 1, num_sms, stream, nullptr, 0,
                   nullptr, 0, scale_a, scale_b, rng, drop_stream, drop_p, nullptr);
}
