// Counter-based dropout masks shared by every kernel that applies (or re-applies in backward) a dropout site.
//
// No mask tensor is ever stored: forward and backward regenerate the same keep/drop decision from
//   (seed, step counter)  -- a 4-word device state, bumped once per training step so a captured CUDA graph replays
//                            with fresh masks,
//   stream id             -- one per dropout site (embedding, attention probs of layer l, the two residual branches),
//   element coordinates   -- a 32-bit PAIR index: one 32-bit hash word decides two neighbouring elements (16 bits each).
// keep(element) = bits16 >= thr16,  thr16 = round(p * 65536); kept values are scaled by 1 / (1 - p).
// ops/reference.py implements the identical arithmetic in torch (the test oracle and the CPU path).
//
// Parity: HF GPT-2 trains with embd/attn/resid dropout 0.1 and the reference miner puts the model in train mode
// (reference hivetrain/training_manager.py:46).
#pragma once
#include <cstdint>

namespace dtb {

struct DropArgs {
  const uint32_t* rng;  // device state {seed, counter, -, -}; nullptr / thr == 0 disables the site
  uint32_t stream;
  uint32_t thr;         // drop if bits16 < thr
  float scale;          // 1 / (1 - p)
};

__device__ __forceinline__ uint32_t mix32s(uint32_t x) {  // two multiply/xorshift rounds (input already decorrelated)
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  return mix32s(x);
}
// per-site key, computed once per thread
__device__ __forceinline__ uint32_t drop_key(const uint32_t* __restrict__ rng, uint32_t stream) {
  const uint32_t seed = rng[0], counter = rng[1];
  uint32_t k = mix32(seed + counter * 0x9E3779B9u);
  return mix32(k ^ (stream * 0x85EBCA6Bu + 0xC2B2AE35u));
}
// hash word for pair index `pair`: low half decides the even element, high half the odd one
__device__ __forceinline__ uint32_t drop_word(uint32_t key, uint32_t pair) { return mix32s(pair ^ key); }
__device__ __forceinline__ float drop_mul_lo(uint32_t word, uint32_t thr, float scale) {
  return (word & 0xFFFFu) >= thr ? scale : 0.f;
}
__device__ __forceinline__ float drop_mul_hi(uint32_t word, uint32_t thr, float scale) {
  return (word >> 16) >= thr ? scale : 0.f;
}
// attention sites: the pair index is local to one (head, query row); its key folds the row in
__device__ __forceinline__ uint32_t drop_row_key(uint32_t key, uint32_t head_row) { return mix32(key ^ head_row); }

}  // namespace dtb
