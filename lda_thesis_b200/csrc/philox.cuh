// Philox4x32-10 (Salmon et al., SC'11; Random123 philox.h) for sm_100a device code.
//
// Draw addressing shared by every sampler in this repository (DESIGN.md "RNG"): one 32-bit word per
// (stream, sweep, global document id d, position p of the draw inside d)
//   ctr = (p >> 2, d, sweep, stream), key = (lo32(seed), hi32(seed)),  word = philox4x32_10(ctr, key)[p & 3]
// It replaces the single multinom_draw(1, prob) call per pair of the reference
// (LabeledLDA.py:119, CascadeLDA.py:415, HSLDA.py:261), whose legacy MT19937 stream consumes a
// data-dependent number of uniforms per draw and cannot be addressed by counter.  Addressing by document
// makes the stream independent of the sharding over GPUs, and a thread walking one document needs one
// Philox block per four draws.
#pragma once
#include <stdint.h>

#define GIBBS_STREAM_SWEEP 0u
#define GIBBS_STREAM_INIT  1u
#define GIBBS_STREAM_TEST  2u
#define GIBBS_STREAM_TEST_INIT 4u

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// All four words of the block that holds positions 4*blk .. 4*blk+3 of document `doc`.
__device__ __forceinline__ uint4 philox_block(uint32_t blk, uint32_t doc, uint32_t sweep, uint32_t stream, uint2 key) {
    return philox4x32_10(make_uint4(blk, doc, sweep, stream), key);
}

__device__ __forceinline__ uint32_t select_word(const uint4 &w, uint32_t sel) {
    return sel == 0 ? w.x : (sel == 1 ? w.y : (sel == 2 ? w.z : w.w));
}

// One word for position `pos` of document `doc` (scalar use: exact mode, init).
__device__ __forceinline__ uint32_t philox_word(uint32_t doc, uint32_t pos, uint32_t sweep, uint32_t stream, uint2 key) {
    const uint4 w = philox_block(pos >> 2, doc, sweep, stream, key);
    return select_word(w, pos & 3u);
}

// fp32 uniform in [0,1) with 24-bit resolution: (word >> 8) * 2^-24 (both steps exact).
__device__ __forceinline__ float u01_f32(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
// fp64 uniform in (0,1) with 32-bit resolution: (word + 0.5) * 2^-32 (exact).
__device__ __forceinline__ double u01_f64(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }
