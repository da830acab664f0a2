// Device kernels for the Labeled-LDA family sweep (LabeledLDA.py:101-125, CascadeLDA.py:397-421).
// sm_100a only.  See DESIGN.md for the data layout and the per-kernel byte model.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"

// ----------------------------------------------------------------------------------------------
// Draw record: 8 bytes per draw.  x = word id, y = (f << 16) | j  where j is the index of the
// draw's current topic inside its document's label list (z = lab_idx[lab0 + j]).
// ----------------------------------------------------------------------------------------------
#define REC_F(y)  ((int)((unsigned)(y) >> 16))
#define REC_J(y)  ((int)((unsigned)(y) & 0xffffu))
#define REC_PACK(f, j) ((int)(((unsigned)(f) << 16) | (unsigned)(j)))

// Label lists of at most GIBBS_SERIAL_MAX topics add their weights left to right (one thread can own the
// document); longer lists add them in 32-lane Kogge-Stone chunks.  Same constant as ORACLE_SERIAL_MAX.
#define GIBBS_SERIAL_MAX 8

// One document as the sampling kernels see it (32 bytes, one per work-list entry).
// Record i of the document is R[rbase + i * stride]: stride 1 for group-per-document kernels, stride 32 for the
// thread-per-document kernel (32 documents of a warp interleaved, so a warp reads one 256-byte line per step).
struct __align__(16) DocDesc {
    long long rbase;
    long long lab0;     // first entry of the label list in lab_idx / n_dk_act
    int len;            // draws
    int A;              // label-list length
    int stride;
    int doc;            // document id inside the shard
};

struct SweepParams {
    const DocDesc   *work;        // work list of this launch
    long long        n_work;
    const int       *lab_idx;     // [n_lab]
    int             *n_dk_act;    // [n_lab]
    int2            *R;           // draw records
    const int       *n_wk;        // [V][ldk] frozen for the refresh block
    int             *delta_wk;    // [V][ldk] +-f land here
    const int       *n_k;         // [K] frozen
    const int       *seg;         // [2*D] or nullptr
    unsigned long long *counter;  // work counter (zeroed before launch)
    unsigned long long *changed;  // draws whose topic changed
    int              ldk;
    int              row_ints;    // ints per ring slot (max segment length, multiple of 4)
    float            alpha, beta, vbeta;
    unsigned         seed_lo, seed_hi, sweep;
    long long        doc_base;    // global id of the shard's document 0 (RNG addressing)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Weight of one active topic, one IEEE operation per statement (mirrored by oracle/gibbs_oracle.c:snapshot_doc):
//   (n_dk + alpha) * (n_wk + beta) / (n_k + V*beta)   with the draw's own f removed from all three counts.
__device__ __forceinline__ float topic_weight(int nd, int nw, int nk, float alpha, float beta, float vbeta) {
    const float a = __fadd_rn((float)nd, alpha);
    const float b = __fadd_rn((float)nw, beta);
    const float cc = __fadd_rn((float)nk, vbeta);
    return __fdiv_rn(__fmul_rn(a, b), cc);
}

// ----------------------------------------------------------------------------------------------
// Snapshot sweep, thread-per-document (label lists of at most GIBBS_SERIAL_MAX topics).
//
// A warp takes a slice of 32 documents (sorted by label-list length, then by length, so the lanes of a slice
// are alike); lane l owns document l for its whole length and keeps the label ids, n_dk and the n_k snapshot of
// its <= AMAX topics in registers.  The records of the 32 documents are interleaved (stride 32), so each step
// is ONE coalesced 256-byte record load per warp; the counts n_wk[v][label] are 4-byte gathers issued two
// steps ahead into registers (they touch |label list| 32-byte sectors of the row, not the whole row).  No
// shuffles, scans or ballots: the weights are added left to right by the owning thread, one Philox block
// serves four consecutive steps of all 32 lanes.  +-f go to the delta table with RED.ADD.
// ----------------------------------------------------------------------------------------------
template <int AMAX>
__device__ __forceinline__ void lane_slice(const SweepParams &p, const DocDesc dd, const int steps, unsigned &n_changed) {
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const int ldk = p.ldk;
    const int *__restrict__ n_wk = p.n_wk;
    const int A = dd.A, len = dd.len;
    const unsigned docg = (unsigned)(p.doc_base + dd.doc);

    int lab[AMAX], ndk[AMAX], nkb[AMAX];
#pragma unroll
    for (int j = 0; j < AMAX; ++j) {
        lab[j] = 0; ndk[j] = 0;
        if (j < A) { lab[j] = p.lab_idx[dd.lab0 + j]; ndk[j] = p.n_dk_act[dd.lab0 + j]; }
    }
#pragma unroll
    for (int j = 0; j < AMAX; ++j) nkb[j] = (j < A) ? p.n_k[lab[j]] - ndk[j] : 0;

    int2 *rp = p.R + dd.rbase;                       // record i at rp[i * 32]
    int2 c[4], nx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        c[q] = (q < len) ? rp[(size_t)q * 32] : make_int2(0, 0);
        nx[q] = (4 + q < len) ? rp[(size_t)(4 + q) * 32] : make_int2(0, 0);
    }
    int g[2][AMAX];                                  // counts of steps s (even slot) and s + 1 (odd slot)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < AMAX; ++j)
            g[q][j] = (j < A && q < len) ? __ldg(n_wk + (size_t)c[q].x * ldk + lab[j]) : 0;

    for (int s0 = 0; s0 < steps; s0 += 4) {
        const uint4 rw = philox_block((unsigned)(s0 >> 2), docg, p.sweep, GIBBS_STREAM_SWEEP, key);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int s = s0 + q;
            const bool live = s < len;
            const int v = c[q].x, y = c[q].y;
            const int f = REC_F(y), jo = REC_J(y);
            int nw[AMAX];
#pragma unroll
            for (int j = 0; j < AMAX; ++j) nw[j] = g[q & 1][j];
            {   // counts of step s + 2
                const int v2 = (q < 2) ? c[q + 2].x : nx[q - 2].x;
                const bool live2 = s + 2 < len;
#pragma unroll
                for (int j = 0; j < AMAX; ++j)
                    g[q & 1][j] = (j < A && live2) ? __ldg(n_wk + (size_t)v2 * ldk + lab[j]) : 0;
            }
            float cum[AMAX];
            float run = 0.0f;
#pragma unroll
            for (int j = 0; j < AMAX; ++j) {
                const int self = (j == jo) ? f : 0;
                const int nd = ndk[j] - self;
                const float w = (j < A) ? topic_weight(nd, nw[j] - self, nkb[j] + nd, alpha, beta, vbeta) : 0.0f;
                run = __fadd_rn(run, w);
                cum[j] = run;
            }
            const unsigned xw = (q == 0) ? rw.x : (q == 1) ? rw.y : (q == 2) ? rw.z : rw.w;
            const float thr = __fmul_rn(u01_f32(xw), run);
            int jn = A - 1;
#pragma unroll
            for (int j = AMAX - 1; j >= 0; --j)
                if (cum[j] > thr) jn = j;            // smallest j with cum[j] > thr (entries past A-1 repeat the total)
            if (live && jn != jo) {
                int lo = 0, ln = 0;
#pragma unroll
                for (int j = 0; j < AMAX; ++j) {
                    if (j == jo) { ndk[j] -= f; lo = lab[j]; }
                    if (j == jn) { ndk[j] += f; ln = lab[j]; }
                }
                int *drow = p.delta_wk + (size_t)v * ldk;
                atomicAdd(drow + lo, -f);
                atomicAdd(drow + ln, f);
                rp[(size_t)s * 32].y = REC_PACK(f, jn);
                ++n_changed;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c[q] = nx[q];
            nx[q] = (s0 + 8 + q < len) ? rp[(size_t)(s0 + 8 + q) * 32] : make_int2(0, 0);
        }
    }
#pragma unroll
    for (int j = 0; j < AMAX; ++j)
        if (j < A) p.n_dk_act[dd.lab0 + j] = ndk[j];
}

__global__ void __launch_bounds__(128) llda_lane_kernel(const SweepParams p) {
    const int lane = threadIdx.x & 31;
    unsigned n_changed = 0;
    while (true) {
        unsigned long long sl = 0;
        if (lane == 0) sl = atomicAdd(p.counter, 1ull);
        sl = __shfl_sync(0xffffffffu, sl, 0);
        const long long idx = (long long)sl * 32 + lane;
        if ((long long)sl * 32 >= p.n_work) break;
        DocDesc dd;
        dd.rbase = 0; dd.lab0 = 0; dd.len = 0; dd.A = 0; dd.stride = 32; dd.doc = 0;
        if (idx < p.n_work) dd = p.work[idx];
        const int amax = __reduce_max_sync(0xffffffffu, dd.A);
        const int steps = __reduce_max_sync(0xffffffffu, dd.len);
        if (amax <= 2) lane_slice<2>(p, dd, steps, n_changed);
        else if (amax <= 4) lane_slice<4>(p, dd, steps, n_changed);
        else if (amax <= 6) lane_slice<6>(p, dd, steps, n_changed);
        else lane_slice<8>(p, dd, steps, n_changed);
    }
    n_changed = __reduce_add_sync(0xffffffffu, n_changed);
    if (lane == 0 && n_changed) atomicAdd(p.changed, (unsigned long long)n_changed);
}

// ----------------------------------------------------------------------------------------------
// Inclusive prefix sum of x over the G lanes of a group.
//   SERIAL: left to right (x0, x0+x1, (x0+x1)+x2, ...) over the first GIBBS_SERIAL_MAX lanes -- the order of the
//           thread-per-document kernel, so both kernels give the same bits for short label lists;
//   else  : Kogge-Stone (offsets 1, 2, 4, ...), whose first G lanes equal those of the 32-lane scan.
// ----------------------------------------------------------------------------------------------
template <int G, bool SERIAL>
__device__ __forceinline__ float group_scan(float x, const int gl, const unsigned gmask) {
    if constexpr (SERIAL) {
        float cum = x;
#pragma unroll
        for (int j = 1; j < GIBBS_SERIAL_MAX; ++j) {
            const float t = __shfl_sync(gmask, cum, j - 1, G);
            if (gl == j) cum = __fadd_rn(t, x);
        }
        return cum;
    } else {
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            const float y = __shfl_up_sync(gmask, x, off, G);
            if (gl >= off) x = __fadd_rn(x, y);
        }
        return x;
    }
}

// ----------------------------------------------------------------------------------------------
// Snapshot sweep, dense row fetch.  A group of G lanes owns one document at a time; lane gl (+32*c for chunk c
// when G == 32) owns one entry of the document's label list.  Per draw:
//   cp.async the draw's n_wk row (segment) into a shared-memory ring R-1 draws ahead,
//   gather the active entries from shared memory, weight as above, prefix sum over the group, Philox uniform,
//   first cum > u*total, +-f to the delta table with RED, n_dk stays in registers until the document ends.
// ----------------------------------------------------------------------------------------------
template <int G, int NCH, int R, bool SERIAL>
__global__ void __launch_bounds__(256) llda_dense_kernel(const SweepParams p) {
    static_assert(G == 32 || NCH == 1, "multi-chunk label lists need a full warp");
    static_assert(!SERIAL || (NCH == 1 && G >= GIBBS_SERIAL_MAX), "serial sums are for short label lists");
    static_assert((R & (R - 1)) == 0 && R >= 2, "ring depth must be a power of two");
    constexpr int P = 2 * R - 1;   // record prefetch distance
    constexpr int M = 2 * R;       // record ring slots
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const int grp = (threadIdx.x >> 5) * GPW + (lane / G);
    const int row_ints = p.row_ints;
    const size_t grp_bytes = (size_t)M * 8 + (size_t)R * row_ints * 4;
    unsigned char *gsm = smem_raw + (size_t)grp * grp_bytes;
    int2 *meta = reinterpret_cast<int2 *>(gsm);
    int *ring = reinterpret_cast<int *>(gsm + M * 8);

    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const int ldk = p.ldk;

    // group state
    DocDesc dd;
    dd.rbase = 0; dd.lab0 = 0; dd.len = 0; dd.A = 0; dd.stride = 1; dd.doc = 0;
    int i = 0, seg_lo = 0, seg_n16 = 0;
    unsigned docg = 0;
    int lab[NCH], ndk[NCH], nkb[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { lab[c] = 0; ndk[c] = 0; nkb[c] = 0; }
    uint4 rw = make_uint4(0, 0, 0, 0);
    int rB0 = -(1 << 30);
    bool have = false, done = false;
    unsigned n_changed = 0;

    auto issue_row = [&](int slot, int v) {
        const int *src = p.n_wk + (size_t)v * ldk + seg_lo;
        int *dst = ring + (size_t)slot * row_ints;
        for (int c16 = gl; c16 < seg_n16; c16 += G) cp_async16(dst + c16 * 4, src + c16 * 4);
    };
    auto rec_at = [&](int q) { return p.R + dd.rbase + (long long)q * dd.stride; };

    while (true) {
        if (!have && !done) {
            unsigned long long di = 0;
            if (gl == 0) di = atomicAdd(p.counter, 1ull);
            di = __shfl_sync(gmask, di, gbase);
            if (di >= (unsigned long long)p.n_work) {
                done = true;
            } else {
                dd = p.work[di];
                docg = (unsigned)(p.doc_base + dd.doc);
                if (p.seg) { seg_lo = p.seg[2 * dd.doc]; seg_n16 = (p.seg[2 * dd.doc + 1] - seg_lo) >> 2; }
                else       { seg_lo = 0; seg_n16 = ldk >> 2; }
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j < dd.A) {
                        lab[c] = p.lab_idx[dd.lab0 + j];
                        ndk[c] = p.n_dk_act[dd.lab0 + j];
                        nkb[c] = p.n_k[lab[c]] - ndk[c];
                    } else { lab[c] = seg_lo; ndk[c] = 0; nkb[c] = 0; }
                }
                if (dd.len > 0) {
                    have = true;
                    i = 0;
                    rB0 = -(1 << 30);
                    for (int q = gl; q < P; q += G)
                        if (q < dd.len) cp_async8(&meta[q & (M - 1)], rec_at(q));
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncwarp(gmask);
#pragma unroll
                    for (int r = 0; r < R - 1; ++r) {
                        if (r < dd.len) issue_row(r, meta[r].x);
                        cp_async_commit();
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (have) {
            const int A = dd.A;
            // -- prefetch: record of draw i+P, row of draw i+R-1
            if (gl == 0 && i + P < dd.len) cp_async8(&meta[(i + P) & (M - 1)], rec_at(i + P));
            if (i + (R - 1) < dd.len) issue_row((i + R - 1) & (R - 1), meta[(i + R - 1) & (M - 1)].x);
            cp_async_commit();
            cp_async_wait<R - 1>();
            __syncwarp(gmask);

            const int2 mt = meta[i & (M - 1)];
            const int v = mt.x, f = REC_F(mt.y), jo = REC_J(mt.y);
            const int *row = ring + (size_t)(i & (R - 1)) * row_ints - seg_lo;

            // -- weights + prefix sums (chunk by chunk; carry is the running total)
            float cum[NCH];
            float carry = 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int j = c * G + gl;
                float x = 0.0f;
                if (j < A) {
                    const int self = (j == jo) ? f : 0;
                    const int nd = ndk[c] - self;
                    x = topic_weight(nd, row[lab[c]] - self, nkb[c] + nd, alpha, beta, vbeta);
                }
                x = group_scan<G, SERIAL>(x, gl, gmask);
                cum[c] = __fadd_rn(carry, x);
                carry = __fadd_rn(carry, __shfl_sync(gmask, x, G - 1, G));
            }
            // total = cum[A-1] (NOT the last lane: a Kogge-Stone lane past A-1 sums the same terms in another order)
            float cl = cum[0];
#pragma unroll
            for (int c = 1; c < NCH; ++c) if (((A - 1) / G) == c) cl = cum[c];
            const float total = __shfl_sync(gmask, cl, (A - 1) & (G - 1), G);

            // -- uniform: lane gl of the group holds the Philox block of positions 4*(rB0+gl) ..
            const int blk = i >> 2;
            if (blk < rB0 || blk >= rB0 + G) {
                rB0 = blk;
                rw = philox_block((unsigned)(blk + gl), docg, p.sweep, GIBBS_STREAM_SWEEP, key);
            }
            const unsigned mine = select_word(rw, (unsigned)(i & 3));
            const unsigned xw = __shfl_sync(gmask, mine, blk - rB0, G);
            const float thr = __fmul_rn(u01_f32(xw), total);

            // -- first index with cum > thr, else A-1
            int jn = A - 1;
            bool found = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int j = c * G + gl;
                unsigned bal = __ballot_sync(gmask, (j < A) && (cum[c] > thr));
                bal >>= gbase;
                if (!found && bal) { jn = c * G + __ffs(bal) - 1; found = true; }
            }

            if (jn != jo) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j == jo) { ndk[c] -= f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab[c]], -f); }
                    else if (j == jn) { ndk[c] += f; atomicAdd(&p.delta_wk[(size_t)v * ldk + lab[c]], f); }
                }
                if (gl == 0) { rec_at(i)->y = REC_PACK(f, jn); ++n_changed; }
            }
            ++i;
            if (i == dd.len) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j = c * G + gl;
                    if (j < A) p.n_dk_act[dd.lab0 + j] = ndk[c];
                }
                have = false;
            }
        }
    }
    if (gl == 0 && n_changed) atomicAdd(p.changed, (unsigned long long)n_changed);
}

// ----------------------------------------------------------------------------------------------
// Snapshot sweep, masked-gather row fetch for label lists of 9..32 topics (G = 16 or 32 lanes per document).
//
// Lane j owns entry j of the label list and keeps n_dk[j] in a register.  The document is walked in chunks of G
// draws: lane i of the group holds the record of draw n0+i (one coalesced 8-byte load per lane per chunk, the
// next chunk's records are requested one chunk ahead) and computes that draw's Philox word; per draw, lane j
// needs ONE count, n_wk[v][lab_j] -- a 4-byte load running R draws ahead in a register ring.  Same arithmetic
// as the dense kernel (Kogge-Stone sums), no shared memory.
// ----------------------------------------------------------------------------------------------
template <int G, int R>
__global__ void __launch_bounds__(256, 3) llda_gather_kernel(const SweepParams p) {
    static_assert(G >= 16 && G <= 32 && (G & (G - 1)) == 0, "group width");
    static_assert(R >= 1 && R <= G && (G % R) == 0, "ring depth must divide the chunk");
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    const int ldk = p.ldk;
    const int *__restrict__ n_wk = p.n_wk;
    unsigned n_changed = 0;

    while (true) {
        // ---- next document of this group
        unsigned long long di = 0;
        if (gl == 0) di = atomicAdd(p.counter, 1ull);
        di = __shfl_sync(gmask, di, 0, G);
        if (di >= (unsigned long long)p.n_work) break;
        const DocDesc dd = p.work[di];
        const int A = dd.A, len = dd.len;
        if (len <= 0) continue;
        const unsigned docg = (unsigned)(p.doc_base + dd.doc);
        int2 *rp = p.R + dd.rbase;
        const long long st = dd.stride;
        int lab = 0, ndk = 0, nkb = 0;
        if (gl < A) {
            lab = p.lab_idx[dd.lab0 + gl];
            ndk = p.n_dk_act[dd.lab0 + gl];
            nkb = p.n_k[lab] - ndk;
        }
        int n0 = 0;
        int2 cur = make_int2(0, 0), nxt = make_int2(0, 0);
        if (gl < len) cur = rp[gl * st];
        if (G + gl < len) nxt = rp[(G + gl) * st];
        int nwq[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = __shfl_sync(gmask, cur.x, r, G);
            nwq[r] = (gl < A && r < len) ? __ldg(n_wk + (size_t)v * ldk + lab) : 0;
        }

        // ---- chunks of G draws
        while (true) {
            const unsigned word = philox_word(docg, (unsigned)(n0 + gl), p.sweep, GIBBS_STREAM_SWEEP, key);
            int newy = cur.y;
            for (int i0 = 0; i0 < G; i0 += R) {
                if (n0 + i0 >= len) break;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = i0 + r;
                    const bool live = n0 + i < len;
                    const int v = __shfl_sync(gmask, cur.x, i, G);
                    const int y = __shfl_sync(gmask, cur.y, i, G);
                    const unsigned xw = __shfl_sync(gmask, word, i, G);
                    const int f = REC_F(y), jo = REC_J(y);
                    const int nw_raw = nwq[r];
                    {   // refill the ring slot with draw i + R (this chunk or the next one)
                        const int ii = i + R;
                        const int srcx = (ii < G) ? cur.x : nxt.x;
                        const int v2 = __shfl_sync(gmask, srcx, ii & (G - 1), G);
                        nwq[r] = (gl < A && n0 + ii < len) ? __ldg(n_wk + (size_t)v2 * ldk + lab) : 0;
                    }
                    float x = 0.0f;
                    const int self = (gl == jo) ? f : 0;
                    const int nd = ndk - self;
                    if (gl < A) x = topic_weight(nd, nw_raw - self, nkb + nd, alpha, beta, vbeta);
                    x = group_scan<G, false>(x, gl, gmask);
                    const float total = __shfl_sync(gmask, x, A - 1, G);
                    const float thr = __fmul_rn(u01_f32(xw), total);
                    unsigned bal = __ballot_sync(gmask, (gl < A) && (x > thr));
                    bal = (bal >> gbase) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
                    const int jn = bal ? (__ffs(bal) - 1) : (A - 1);
                    if (live && jn != jo) {
                        if (gl == jo || gl == jn) {
                            const int df = (gl == jo) ? -f : f;
                            ndk += df;
                            atomicAdd(&p.delta_wk[(size_t)v * ldk + lab], df);
                        }
                        if (gl == i) { newy = REC_PACK(f, jn); ++n_changed; }
                    }
                }
            }
            if (newy != cur.y) rp[(n0 + gl) * st].y = newy;
            n0 += G;
            if (n0 >= len) break;
            cur = nxt;
            nxt = make_int2(0, 0);
            if (n0 + G + gl < len) nxt = rp[(n0 + G + gl) * st];
        }
        if (gl < A) p.n_dk_act[dd.lab0 + gl] = ndk;
    }
    n_changed = __reduce_add_sync(0xffffffffu, n_changed);
    if (lane == 0 && n_changed) atomicAdd(p.changed, (unsigned long long)n_changed);
}

// ----------------------------------------------------------------------------------------------
// Exact sweep: one warp walks the corpus in order with live counts in fp64, the operation order of
// LabeledLDA.py:109-125 (restated by oracle/gibbs_oracle.c:oracle_llda_exact_sweep).
// desc[d] is indexed by document id (corpus order).
// ----------------------------------------------------------------------------------------------
struct ExactParams {
    const DocDesc *desc;
    const int *lab_idx;
    int *n_dk_act;
    int2 *R;
    int *n_wk;
    int *n_k;
    long long d_begin, d_end;
    int ldk;
    double alpha, beta, vbeta;
    unsigned seed_lo, seed_hi, sweep;
    long long doc_base;
    unsigned long long *changed;
};

__global__ void __launch_bounds__(32) llda_exact_kernel(const ExactParams p) {
    const int lane = threadIdx.x;
    volatile int *nwk = p.n_wk;
    volatile int *nk = p.n_k;
    volatile int *ndk = p.n_dk_act;
    const uint2 key = make_uint2(p.seed_lo, p.seed_hi);
    unsigned long long n_changed = 0;
    for (long long d = p.d_begin; d < p.d_end; ++d) {
        const DocDesc dd = p.desc[d];
        const long long lab0 = dd.lab0;
        const int A = dd.A;
        const int nch = (A + 31) >> 5;
        for (int i = 0; i < dd.len; ++i) {
            int2 *rec = p.R + dd.rbase + (long long)i * dd.stride;
            const int2 mt = *rec;
            const int v = mt.x, f = REC_F(mt.y), jo = REC_J(mt.y);
            const int zo = p.lab_idx[lab0 + jo];
            if (lane == 0) {
                nwk[(size_t)v * p.ldk + zo] -= f;   // LabeledLDA.py:109
                ndk[lab0 + jo] -= f;                // :110
                nk[zo] -= f;                        // :111
            }
            __syncwarp();
            const uint32_t xw = philox_word((unsigned)(p.doc_base + d), (unsigned)i, p.sweep, GIBBS_STREAM_SWEEP, key);
            const double u = u01_f64(xw);
            // pass 1: total; pass 2: pick.  Both passes run the same serial sum, so cum values are identical.
            double run = 0.0;
            for (int c = 0; c < nch; ++c) {
                const int j = c * 32 + lane;
                double w = 0.0;
                if (j < A) {
                    const int k = p.lab_idx[lab0 + j];
                    const double a = __dadd_rn((double)ndk[lab0 + j], p.alpha);                  // :113
                    const double num = __dadd_rn((double)nwk[(size_t)v * p.ldk + k], p.beta);    // :114
                    const double den = __dadd_rn((double)nk[k], p.vbeta);                        // :115
                    w = __dmul_rn(a, __ddiv_rn(num, den));                                       // :117
                }
                const int cnt = min(32, A - c * 32);
                for (int l = 0; l < cnt; ++l) run = __dadd_rn(run, __shfl_sync(0xffffffffu, w, l));
            }
            const double thr = __dmul_rn(u, run);
            int jn = A - 1;
            bool found = false;
            double run2 = 0.0;
            for (int c = 0; c < nch && !found; ++c) {
                const int j = c * 32 + lane;
                double w = 0.0;
                if (j < A) {
                    const int k = p.lab_idx[lab0 + j];
                    const double a = __dadd_rn((double)ndk[lab0 + j], p.alpha);
                    const double num = __dadd_rn((double)nwk[(size_t)v * p.ldk + k], p.beta);
                    const double den = __dadd_rn((double)nk[k], p.vbeta);
                    w = __dmul_rn(a, __ddiv_rn(num, den));
                }
                const int cnt = min(32, A - c * 32);
                double cum = 0.0;
                for (int l = 0; l < cnt; ++l) {
                    run2 = __dadd_rn(run2, __shfl_sync(0xffffffffu, w, l));
                    if (lane == l) cum = run2;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, (j < A) && (cum > thr));
                if (bal) { jn = c * 32 + __ffs(bal) - 1; found = true; }
            }
            const int zn = p.lab_idx[lab0 + jn];
            if (lane == 0) {
                if (jn != jo) { rec->y = REC_PACK(f, jn); ++n_changed; }   // :121
                nwk[(size_t)v * p.ldk + zn] += f;    // :123
                ndk[lab0 + jn] += f;                 // :124
                nk[zn] += f;                         // :125
            }
            __syncwarp();
        }
    }
    if (lane == 0 && n_changed) atomicAdd(p.changed, n_changed);
}

// ----------------------------------------------------------------------------------------------
// Corpus preparation, histogram, merge, export.  One warp per work-list entry (= per document); the CSR side
// (word / freq / z in corpus order) is indexed through doc_ptr[dd.doc], the record side through (rbase, stride).
// ----------------------------------------------------------------------------------------------

// Pack records; z from z_init (global topic ids -> label-list index) or Uniform(label list) from Philox stream 1.
// err[0] is set when a z_init value is not in the list.
__global__ void prepare_records_kernel(long long n_work, const DocDesc *work, const long long *doc_ptr, const int *word,
                                       const int *freq, const int *z_init, const int *lab_idx, int2 *R, int V, uint2 key,
                                       long long doc_base, int *err) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const DocDesc dd = work[w];
    const long long src = doc_ptr[dd.doc];
    for (int i = lane; i < dd.len; i += 32) {
        const int f = freq ? freq[src + i] : 1;
        int j = -1;
        if (z_init) {
            const int zg = z_init[src + i];
            for (int q = 0; q < dd.A; ++q)
                if (lab_idx[dd.lab0 + q] == zg) { j = q; break; }
        } else if (dd.A > 0) {
            const uint32_t x = philox_word((unsigned)(doc_base + dd.doc), (unsigned)i, 0u, GIBBS_STREAM_INIT, key);
            j = (int)(((uint64_t)x * (uint64_t)dd.A) >> 32);
        }
        int v = word[src + i];
        if (j < 0 || f < 0 || f > 0xffff || v < 0 || v >= V) { atomicExch(err, 1); j = 0; v = 0; }
        R[dd.rbase + (long long)i * dd.stride] = make_int2(v, REC_PACK(f, j));
    }
}

// Histogram (LabeledLDA.py:89-92).  n_k is NOT touched here: it is the column sum of n_wk (column_sums_kernel).
__global__ void counts_build_kernel(long long n_work, const DocDesc *work, const int *lab_idx, const int2 *R, int ldk,
                                    int *n_wk, int *n_dk_act) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const DocDesc dd = work[w];
    for (int i = lane; i < dd.len; i += 32) {
        const int2 r = R[dd.rbase + (long long)i * dd.stride];
        const int f = REC_F(r.y), j = REC_J(r.y);
        const int k = lab_idx[dd.lab0 + j];
        atomicAdd(&n_wk[(size_t)r.x * ldk + k], f);
        atomicAdd(&n_dk_act[dd.lab0 + j], f);
    }
}

// n_wk[word][topic] += count for a COO list (duplicates accumulate).
__global__ void add_counts_kernel(long long n, const int *word, const int *topic, const int *count, int ldk, int *n_wk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&n_wk[(size_t)word[i] * ldk + topic[i]], count[i]);
}

// Replace z: global topic ids -> label-list index inside the existing records.
__global__ void set_z_kernel(long long n_work, const DocDesc *work, const long long *doc_ptr, const int *lab_idx,
                             const int *z, int2 *R, int *err) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const DocDesc dd = work[w];
    const long long src = doc_ptr[dd.doc];
    for (int i = lane; i < dd.len; i += 32) {
        int j = -1;
        for (int q = 0; q < dd.A; ++q)
            if (lab_idx[dd.lab0 + q] == z[src + i]) { j = q; break; }
        if (j < 0) { atomicExch(err, 1); j = 0; }
        int2 *r = R + dd.rbase + (long long)i * dd.stride;
        r->y = REC_PACK(REC_F(r->y), j);
    }
}

__global__ void export_z_kernel(long long n_work, const DocDesc *work, const long long *doc_ptr, const int *lab_idx,
                                const int2 *R, int *z_out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const DocDesc dd = work[w];
    const long long src = doc_ptr[dd.doc];
    for (int i = lane; i < dd.len; i += 32)
        z_out[src + i] = lab_idx[dd.lab0 + REC_J(R[dd.rbase + (long long)i * dd.stride].y)];
}

// n_wk += delta; delta = 0; n_k += column sums of delta.  blockDim = (ldk/4 capped to 256, rows per block).
// Thread (x, y) owns int4 column chunks x, x + blockDim.x, ... so its topic columns are fixed; four rows are in
// flight per thread (both tables are read unconditionally) so the pass runs at memory speed; untouched 16-byte chunks
// are not written.
// The column sums of a block are first combined in shared memory (blockDim.y rows per column), so n_k sees one RED
// per column per block instead of one per thread: K addresses sit in K/8 sectors and same-sector REDs serialise
// in L2 (ncu r02a: 1.2 M REDs on 16 sectors were 90 % of this kernel's time).
__global__ void __launch_bounds__(256) merge_delta_kernel(int4 *__restrict__ n_wk, int4 *__restrict__ delta,
                                                          int *__restrict__ n_k, long long V, int ldk4, int K) {
    extern __shared__ int col_acc[];                 // [ldk4 * 4]
    const long long stride = (long long)gridDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int q = tid; q < ldk4 * 4; q += nthr) col_acc[q] = 0;
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < ldk4; c4 += blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (long long v0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; v0 < V; v0 += 4 * stride) {
            int4 dl[4], t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {      // eight independent 16-byte loads in flight per thread
                const long long v = v0 + u * stride;
                const size_t idx = (size_t)v * ldk4 + c4;
                dl[u] = (v < V) ? delta[idx] : make_int4(0, 0, 0, 0);
                t[u] = (v < V) ? n_wk[idx] : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (dl[u].x | dl[u].y | dl[u].z | dl[u].w) {
                    const size_t idx = (size_t)(v0 + u * stride) * ldk4 + c4;
                    t[u].x += dl[u].x; t[u].y += dl[u].y; t[u].z += dl[u].z; t[u].w += dl[u].w;
                    n_wk[idx] = t[u];
                    delta[idx] = make_int4(0, 0, 0, 0);
                    acc.x += dl[u].x; acc.y += dl[u].y; acc.z += dl[u].z; acc.w += dl[u].w;
                }
            }
        }
        const int k = c4 * 4;
        if (acc.x) atomicAdd(&col_acc[k], acc.x);
        if (acc.y) atomicAdd(&col_acc[k + 1], acc.y);
        if (acc.z) atomicAdd(&col_acc[k + 2], acc.z);
        if (acc.w) atomicAdd(&col_acc[k + 3], acc.w);
    }
    __syncthreads();
    for (int k = tid; k < K; k += nthr) {
        const int a = col_acc[k];
        if (a) atomicAdd(&n_k[k], a);
    }
}

// Column sums of the word-major table: out[k] += sum_v n_wk[v][k]  (= row sums of the reference's n_k_v).
// Same thread layout as merge_delta_kernel; out must be zeroed by the caller.
__global__ void column_sums_kernel(const int4 *__restrict__ n_wk, int *__restrict__ out, long long V, int ldk4, int K) {
    extern __shared__ int col_acc[];                 // [ldk4 * 4]
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int q = tid; q < ldk4 * 4; q += nthr) col_acc[q] = 0;
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < ldk4; c4 += blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < V;
             v += (long long)gridDim.x * blockDim.y) {
            const int4 t = n_wk[(size_t)v * ldk4 + c4];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const int k = c4 * 4;
        if (acc.x) atomicAdd(&col_acc[k], acc.x);
        if (acc.y) atomicAdd(&col_acc[k + 1], acc.y);
        if (acc.z) atomicAdd(&col_acc[k + 2], acc.z);
        if (acc.w) atomicAdd(&col_acc[k + 3], acc.w);
    }
    __syncthreads();
    for (int k = tid; k < K; k += nthr) {
        const int a = col_acc[k];
        if (a) atomicAdd(&out[k], a);
    }
}

// phi[k][v] from n_wk[v][k]: 32x32 tile transpose through shared memory.
// smoothed: (n + beta) / (den[k] + V*beta) with den = n_k (LabeledLDA.py:231-234);
// otherwise n / den[k] with den = column sums of the table (CascadeLDA.py:394-395; 0/0 -> NaN as in NumPy).
// accumulate: phi = c_old * phi + c_new * current -- the thinning mean of LabeledLDA.py:143-144 / CascadeLDA.py:431-432,
// one rounding per operation as NumPy evaluates it.
__global__ void emit_phi_kernel(const int *__restrict__ n_wk, const int *__restrict__ den_k, double *__restrict__ phi,
                                int V, int K, int ldk, double beta, double vbeta, int smoothed,
                                int accumulate, double c_old, double c_new) {
    __shared__ int tile[32][33];
    const int v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int v = v0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (v < V && k < ldk) ? n_wk[(size_t)v * ldk + k] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, v = v0 + threadIdx.x;
        if (k < K && v < V) {
            const int c = tile[threadIdx.x][r];
            const double den = smoothed ? __dadd_rn((double)den_k[k], vbeta) : (double)den_k[k];
            const double num = smoothed ? __dadd_rn((double)c, beta) : (double)c;
            double out = __ddiv_rn(num, den);
            if (accumulate) out = __dadd_rn(__dmul_rn(c_old, phi[(size_t)k * V + v]), __dmul_rn(c_new, out));
            phi[(size_t)k * V + v] = out;
        }
    }
}

// theta[d][:] dense.  One warp per document (lab_ptr order = document order); the row is zero-filled, then the
// active entries written.
__global__ void emit_theta_kernel(long long D, const long long *lab_ptr, const int *lab_idx, const int *n_dk_act,
                                  double *theta, int K, double alpha, int smoothed) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    double *row = theta + (size_t)d * K;
    for (int k = lane; k < K; k += 32) row[k] = 0.0;
    double den = 0.0;
    for (int j = 0; j < A; ++j) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        den = __dadd_rn(den, x);
    }
    __syncwarp();
    for (int j = lane; j < A; j += 32) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        row[lab_idx[lab0 + j]] = __ddiv_rn(x, den);
    }
}

// theta over the label lists only (the non-zero entries of LabeledLDA.py:236-239), aligned with lab_idx.
// accumulate: thinning mean as in emit_phi_kernel (entries outside the label list stay 0 in the reference too).
__global__ void emit_theta_csr_kernel(long long D, const long long *lab_ptr, const int *n_dk_act, double *theta_act,
                                      double alpha, int smoothed, int accumulate, double c_old, double c_new) {
    const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (d >= D) return;
    const long long lab0 = lab_ptr[d];
    const int A = (int)(lab_ptr[d + 1] - lab0);
    double den = 0.0;
    for (int j = 0; j < A; ++j) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        den = __dadd_rn(den, x);
    }
    for (int j = lane; j < A; j += 32) {
        const double x = smoothed ? __dadd_rn((double)n_dk_act[lab0 + j], alpha) : (double)n_dk_act[lab0 + j];
        double out = __ddiv_rn(x, den);
        if (accumulate) out = __dadd_rn(__dmul_rn(c_old, theta_act[lab0 + j]), __dmul_rn(c_new, out));
        theta_act[lab0 + j] = out;
    }
}

// Training perplexity, LabeledLDA.py:256-265: per document the sum over its (unique) word ids of
// -log(phi[:, w] . theta_d) with the smoothed phi / theta of get_phi / get_theta; theta_d is zero outside the label
// list, so the inner product runs over the list.  One warp per work-list entry; out[doc] is written, never added to.
__global__ void perplexity_doc_kernel(long long n_work, const DocDesc *work, const int *lab_idx, const int *n_dk_act,
                                      const int2 *R, const int *n_wk, const int *n_k, int ldk, double alpha, double beta,
                                      double vbeta, double *out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const DocDesc dd = work[w];
    double den = 0.0;
    for (int j = 0; j < dd.A; ++j) den = __dadd_rn(den, __dadd_rn((double)n_dk_act[dd.lab0 + j], alpha));
    double acc = 0.0;
    for (int i = lane; i < dd.len; i += 32) {
        const int v = R[dd.rbase + (long long)i * dd.stride].x;
        double dot = 0.0;
        for (int j = 0; j < dd.A; ++j) {
            const int k = lab_idx[dd.lab0 + j];
            const double th = __ddiv_rn(__dadd_rn((double)n_dk_act[dd.lab0 + j], alpha), den);
            const double ph = __ddiv_rn(__dadd_rn((double)n_wk[(size_t)v * ldk + k], beta), __dadd_rn((double)n_k[k], vbeta));
            dot = __dadd_rn(dot, __dmul_rn(ph, th));
        }
        acc -= log(dot);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[dd.doc] = acc;
}

// Fixed-order sum of n doubles into out[0] (one block; thread t adds elements t, t + 1024, ... then a shared-memory tree).
__global__ void __launch_bounds__(1024) sum_f64_kernel(const double *x, long long n, double *out) {
    __shared__ double part[1024];
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) a += x[i];
    part[threadIdx.x] = a;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = part[0];
}

__global__ void philox_kat_kernel(int n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr4[4 * i], ctr4[4 * i + 1], ctr4[4 * i + 2], ctr4[4 * i + 3]),
                                  make_uint2(key2[2 * i], key2[2 * i + 1]));
    out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
}
