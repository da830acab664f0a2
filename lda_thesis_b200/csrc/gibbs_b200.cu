// libgibbs_b200.so -- C-ABI implementation (include/gibbs_b200.h).  sm_100a only, no CPU path.
#include "../../include/gibbs_b200.h"
#include "llda_kernels.cuh"
#include "hslda_kernels.cuh"
#include "test_kernels.cuh"
#include "nccl_dl.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return fail(e == cudaErrorMemoryAllocation ? GIBBS_E_NOMEM : GIBBS_E_CUDA, buf);
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call, __FILE__, __LINE__); } while (0)
#define TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)
#define NCK(call) do { int r_ = (call); if (r_ != 0) { char b_[256]; snprintf(b_, sizeof b_, "%s failed: %s", #call, nccl_dl::error_string(r_)); return fail(GIBBS_E_CUDA, b_); } } while (0)

extern "C" const char *gibbs_last_error(void) { return g_err.c_str(); }
extern "C" const char *gibbs_version(void) { return "gibbs_b200 0.2 (sm_100a)"; }
extern "C" int gibbs_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------------- handle
struct DocList { long long off = 0, len = 0; };   // range inside the work list
static const int N_BINS = 7;                       // label-list length <= 8, 16, 32, 64, 128, 256, 512
static const int BIN_CAP[N_BINS] = {GIBBS_SERIAL_MAX, 16, 32, 64, 128, 256, 512};
static const int GATHER_BINS = 3;                  // bins 0..2 (<= 32 labels) have a masked-gather kernel

// Device buffer that only ever grows: gibbs_load on a live handle re-uses the previous allocation.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;     // elements
};

struct gibbs_handle {
    gibbs_desc desc{};
    int ldk = 0, sm_count = 0, max_active = 0, row_ints = 0;
    long long N = 0, n_lab = 0;
    bool loaded = false, has_seg = false;
    uint32_t sweep = 0;
    cudaStream_t stream = nullptr;
    // corpus
    DevBuf<long long> doc_ptr, lab_ptr;
    DevBuf<int> lab_idx, seg;
    DevBuf<DocDesc> work;                  // one entry per document, [refresh block][bin] segments
    DevBuf<int2> rec;                      // draw records R (see DocDesc)
    // counts
    DevBuf<int> n_wk, delta_wk, n_k, n_dk_act, colsum;
    DevBuf<unsigned long long> counters;   // [0] work counter, [1] changed
    DevBuf<int> err_flag;
    DevBuf<unsigned char> scratch;         // uploads, z export, phi/theta staging
    DevBuf<double> hat_phi, hat_theta, doc_sum;   // thinning means (gibbs_thin_accumulate), perplexity partials
    bool hat_phi_set = false, hat_theta_set = false;
    std::vector<DocList> lists;            // [block * N_BINS + bin]
    long long bin_draws[N_BINS] = {0};
    // multi-GPU
    void *comm = nullptr;
    int nranks = 1, rank = 0;
    // hslda
    HsldaState hs{};
    // stats
    gibbs_stats_t st{};
    std::vector<cudaEvent_t> ev;
    size_t dev_bytes = 0;
};

static bool bin_uses_lane(const gibbs_handle *h);
static bool bin_uses_gather(const gibbs_handle *h, int bin);

template <typename T>
static int reserve(gibbs_handle *h, DevBuf<T> &b, size_t n) {
    if (n == 0) n = 1;
    if (b.cap >= n) return 0;
    if (b.p) { cudaFree(b.p); h->dev_bytes -= b.cap * sizeof(T); b.p = nullptr; b.cap = 0; }
    CK(cudaMalloc((void **)&b.p, n * sizeof(T)));
    b.cap = n;
    h->dev_bytes += n * sizeof(T);
    return 0;
}
template <typename T>
static void release(gibbs_handle *h, DevBuf<T> &b) {
    if (b.p) { cudaFree(b.p); h->dev_bytes -= b.cap * sizeof(T); }
    b.p = nullptr; b.cap = 0;
}

static cudaEvent_t get_event(gibbs_handle *h, size_t i) {
    while (h->ev.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        h->ev.push_back(e);
    }
    return h->ev[i];
}

extern "C" int gibbs_create(gibbs_t **out, const gibbs_desc *desc) {
    if (!out || !desc) return fail(GIBBS_E_ARG, "gibbs_create: null argument");
    *out = nullptr;
    if (desc->D < 0 || desc->V <= 0 || desc->K <= 0) return fail(GIBBS_E_ARG, "gibbs_create: D, V, K must be positive");
    if (desc->D > 0x7fffffffLL) return fail(GIBBS_E_ARG, "gibbs_create: more than 2^31-1 documents in one shard");
    if (desc->kind != GIBBS_KIND_LLDA && desc->kind != GIBBS_KIND_HSLDA) return fail(GIBBS_E_ARG, "gibbs_create: unknown kind");
    if (desc->mode != GIBBS_MODE_EXACT && desc->mode != GIBBS_MODE_SNAPSHOT) return fail(GIBBS_E_ARG, "gibbs_create: unknown mode");
    if (desc->row_fetch < GIBBS_FETCH_AUTO || desc->row_fetch > GIBBS_FETCH_GATHER) return fail(GIBBS_E_ARG, "gibbs_create: unknown row_fetch");
    if (desc->n_refresh > 31) return fail(GIBBS_E_ARG, "gibbs_create: n_refresh too large (max 31)");
    if (desc->kind == GIBBS_KIND_HSLDA) return fail(GIBBS_E_STATE, "gibbs_create: GIBBS_KIND_HSLDA is not implemented in this build");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(GIBBS_E_CUDA, std::string("gibbs_create: no CUDA device (") + cudaGetErrorString(e) +
                                      "); this library has no CPU path");
    if (desc->device < 0 || desc->device >= ndev) return fail(GIBBS_E_ARG, "gibbs_create: bad device ordinal");
    CK(cudaSetDevice(desc->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, desc->device));
    if (prop.major < 10) return fail(GIBBS_E_CUDA, "gibbs_create: device is not sm_100 class; kernels are built for sm_100a only");
    cudaStream_t stream = nullptr;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    gibbs_handle *h = new gibbs_handle();
    h->stream = stream;
    h->desc = *desc;
    if (h->desc.n_refresh < 1) h->desc.n_refresh = 1;
    if (h->desc.mode == GIBBS_MODE_EXACT) h->desc.n_refresh = 1;
    if (h->desc.tile_docs <= 0) h->desc.tile_docs = 256;
    h->ldk = (desc->K + 31) / 32 * 32;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return 0;
}

extern "C" void gibbs_destroy(gibbs_t *h) {
    if (!h) return;
    cudaSetDevice(h->desc.device);
    cudaStreamSynchronize(h->stream);
    if (h->comm) nccl_dl::comm_destroy(h->comm);
    release(h, h->doc_ptr); release(h, h->lab_ptr); release(h, h->lab_idx); release(h, h->seg);
    release(h, h->work); release(h, h->rec); release(h, h->n_wk); release(h, h->delta_wk);
    release(h, h->n_k); release(h, h->n_dk_act); release(h, h->colsum); release(h, h->counters);
    release(h, h->err_flag); release(h, h->scratch);
    release(h, h->hat_phi); release(h, h->hat_theta); release(h, h->doc_sum);
    hslda_free(&h->hs);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    delete h;
}

// ---------------------------------------------------------------------------------------------- multi-GPU
extern "C" int gibbs_comm_unique_id(char *id) {
    if (!id) return fail(GIBBS_E_ARG, "gibbs_comm_unique_id: null argument");
    std::string why;
    if (!nccl_dl::load(&why)) return fail(GIBBS_E_CUDA, "gibbs_comm_unique_id: " + why);
    NCK(nccl_dl::get_unique_id(id));
    return 0;
}

extern "C" int gibbs_comm_init(gibbs_t *h, int32_t nranks, int32_t rank, const char *id) {
    if (!h || !id) return fail(GIBBS_E_ARG, "gibbs_comm_init: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GIBBS_E_ARG, "gibbs_comm_init: bad rank / nranks");
    if (h->loaded) return fail(GIBBS_E_STATE, "gibbs_comm_init: must be called before gibbs_load (the initial counts are all-reduced there)");
    if (h->comm) return fail(GIBBS_E_STATE, "gibbs_comm_init: communicator already initialised");
    if (h->desc.mode != GIBBS_MODE_SNAPSHOT) return fail(GIBBS_E_STATE, "gibbs_comm_init: only the snapshot schedule shards across GPUs");
    std::string why;
    if (!nccl_dl::load(&why)) return fail(GIBBS_E_CUDA, "gibbs_comm_init: " + why);
    CK(cudaSetDevice(h->desc.device));
    NCK(nccl_dl::comm_init_rank(&h->comm, nranks, id, rank));
    h->nranks = nranks;
    h->rank = rank;
    return 0;
}

// ---------------------------------------------------------------------------------------------- load
static int column_sums(gibbs_handle *h, int *out) {
    const int ldk4 = h->ldk / 4;
    int bx = std::min(256, (ldk4 + 31) / 32 * 32);
    int by = std::max(1, 256 / bx);
    unsigned grid = (unsigned)std::min<long long>((h->desc.V + by - 1) / by, (long long)h->sm_count * 8);
    CK(cudaMemsetAsync(out, 0, sizeof(int) * (size_t)h->desc.K, h->stream));
    column_sums_kernel<<<grid, dim3(bx, by), sizeof(int) * (size_t)h->ldk, h->stream>>>(reinterpret_cast<const int4 *>(h->n_wk.p), out, h->desc.V, ldk4, h->desc.K);
    CK(cudaGetLastError());
    return 0;
}

static int rebuild_counts(gibbs_handle *h) {
    const size_t tab = (size_t)h->desc.V * h->ldk;
    CK(cudaMemsetAsync(h->n_wk.p, 0, sizeof(int) * tab, h->stream));
    if (h->delta_wk.p) CK(cudaMemsetAsync(h->delta_wk.p, 0, sizeof(int) * tab, h->stream));
    CK(cudaMemsetAsync(h->n_dk_act.p, 0, sizeof(int) * (size_t)std::max<long long>(h->n_lab, 1), h->stream));
    if (h->desc.D > 0) {
        const long long blocks = (h->desc.D * 32 + 255) / 256;
        counts_build_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->desc.D, h->work.p, h->lab_idx.p, h->rec.p, h->ldk,
                                                                     h->n_wk.p, h->n_dk_act.p);
        CK(cudaGetLastError());
    }
    // every rank holds the full word-topic table: sum of all shards' histograms
    if (h->comm) NCK(nccl_dl::all_reduce_i32(h->n_wk.p, tab, h->comm, h->stream));
    TRY(column_sums(h, h->n_k.p));     // n_zk = row sums of n_k_v (LabeledLDA.py:90,92 add the same f to both)
    return 0;
}

extern "C" int gibbs_load(gibbs_t *h, const int64_t *doc_ptr, const int32_t *word, const int32_t *freq,
                          const int32_t *z_init, const int64_t *lab_ptr, const int32_t *lab_idx, const int32_t *seg) {
    if (!h || !doc_ptr || !lab_ptr) return fail(GIBBS_E_ARG, "gibbs_load: null argument");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    if (doc_ptr[D] > 0 && !word) return fail(GIBBS_E_ARG, "gibbs_load: null word array");
    if (lab_ptr[D] > 0 && !lab_idx) return fail(GIBBS_E_ARG, "gibbs_load: null lab_idx array");
    h->loaded = false;
    // ---- host validation
    if (doc_ptr[0] != 0 || lab_ptr[0] != 0) return fail(GIBBS_E_ARG, "gibbs_load: CSR offsets must start at 0");
    int max_a = 0;
    for (long long d = 0; d < D; ++d) {
        if (doc_ptr[d + 1] < doc_ptr[d] || lab_ptr[d + 1] < lab_ptr[d]) return fail(GIBBS_E_ARG, "gibbs_load: CSR offsets must be non-decreasing");
        const long long a = lab_ptr[d + 1] - lab_ptr[d];
        if (a < 1 || a > BIN_CAP[N_BINS - 1]) {
            char b[160]; snprintf(b, sizeof b, "gibbs_load: document %lld has %lld active topics (supported: 1..%d)", d, a, BIN_CAP[N_BINS - 1]);
            return fail(GIBBS_E_ARG, b);
        }
        max_a = std::max<int>(max_a, (int)a);
        for (long long q = lab_ptr[d]; q < lab_ptr[d + 1]; ++q) {
            if (lab_idx[q] < 0 || lab_idx[q] >= h->desc.K) return fail(GIBBS_E_ARG, "gibbs_load: lab_idx out of range");
            if (q > lab_ptr[d] && lab_idx[q] <= lab_idx[q - 1]) return fail(GIBBS_E_ARG, "gibbs_load: lab_idx must be strictly ascending inside a document");
        }
    }
    if (h->desc.doc_base < 0 || h->desc.doc_base + D > 0xffffffffLL) return fail(GIBBS_E_ARG, "gibbs_load: global document ids must fit 32 bits");
    h->N = doc_ptr[D];
    h->n_lab = lab_ptr[D];
    h->max_active = max_a;
    h->row_ints = h->ldk;
    h->has_seg = seg != nullptr;
    if (seg) {
        int mx = 0;
        for (long long d = 0; d < D; ++d) {
            const int lo = seg[2 * d], hi = seg[2 * d + 1];
            if (lo < 0 || hi > h->ldk || lo >= hi || (lo & 3) || (hi & 3)) return fail(GIBBS_E_ARG, "gibbs_load: bad topic segment (need 0 <= lo < hi <= ldk, multiples of 4)");
            for (long long q = lab_ptr[d]; q < lab_ptr[d + 1]; ++q)
                if (lab_idx[q] < lo || lab_idx[q] >= hi) return fail(GIBBS_E_ARG, "gibbs_load: label outside the document's topic segment");
            mx = std::max(mx, hi - lo);
        }
        h->row_ints = std::max(mx, 4);
    }

    // ---- work list: one DocDesc per document in [refresh block][bin] segments; inside a segment longest first
    // (thread-per-document segments: by label-list length, then length, so the 32 lanes of a slice are alike)
    const int nb = h->desc.n_refresh;
    const bool exact = h->desc.mode == GIBBS_MODE_EXACT;
    std::vector<DocDesc> work((size_t)std::max<long long>(D, 1));
    long long r_total = 0;
    {
        for (int b = 0; b < N_BINS; ++b) h->bin_draws[b] = 0;
        h->lists.assign((size_t)nb * N_BINS, DocList());
        if (exact) {      // corpus order, one segment
            for (long long d = 0; d < D; ++d) {
                DocDesc &w = work[(size_t)d];
                w.rbase = doc_ptr[d]; w.lab0 = lab_ptr[d]; w.len = (int)(doc_ptr[d + 1] - doc_ptr[d]);
                w.A = (int)(lab_ptr[d + 1] - lab_ptr[d]); w.stride = 1; w.doc = (int)d;
            }
            h->lists[0].off = 0; h->lists[0].len = D;
            r_total = h->N;
        } else {
            long long max_len = 0;
            for (long long d = 0; d < D; ++d) max_len = std::max<long long>(max_len, doc_ptr[d + 1] - doc_ptr[d]);
            if (max_len > 0x7fffffffLL) return fail(GIBBS_E_ARG, "gibbs_load: a document has more than 2^31-1 draws");
            std::vector<unsigned long long> keyed((size_t)D);      // (segment, sort key, doc) packed for one sort
            const unsigned long long LB = (unsigned long long)max_len + 1;
            if ((unsigned long long)nb * N_BINS * (GIBBS_SERIAL_MAX + 1) * LB >= (1ull << 32))
                return fail(GIBBS_E_ARG, "gibbs_load: documents too long for the work-list sort key");
            std::vector<long long> cnt((size_t)nb * N_BINS, 0);
            for (long long d = 0; d < D; ++d) {
                const long long tile = (h->desc.doc_base + d) / h->desc.tile_docs;
                const int b = (int)(tile % nb);
                const int a = (int)(lab_ptr[d + 1] - lab_ptr[d]);
                int bin = 0;
                while (BIN_CAP[bin] < a) ++bin;
                const long long len = doc_ptr[d + 1] - doc_ptr[d];
                const int seg_id = b * N_BINS + bin;
                const unsigned long long arank = bin == 0 ? (unsigned long long)(GIBBS_SERIAL_MAX - a) : 0ull;
                const unsigned long long key = ((unsigned long long)seg_id * (GIBBS_SERIAL_MAX + 1) + arank) * LB + (unsigned long long)(max_len - len);
                keyed[(size_t)d] = (key << 32) | (unsigned long long)d;
                cnt[(size_t)seg_id]++;
                h->bin_draws[bin] += len;
            }
            std::sort(keyed.begin(), keyed.end());
            long long off = 0;
            for (size_t q = 0; q < cnt.size(); ++q) { h->lists[q].off = off; h->lists[q].len = cnt[q]; off += cnt[q]; }
            for (size_t q = 0; q < cnt.size(); ++q) {
                const int bin = (int)(q % N_BINS);
                const bool sliced = bin == 0 && bin_uses_lane(h);
                const long long o = h->lists[q].off, n = h->lists[q].len;
                for (long long s0 = 0; s0 < n; s0 += 32) {
                    const long long cntl = std::min<long long>(32, n - s0);
                    long long mx = 0;
                    for (long long l = 0; l < cntl; ++l) {
                        const long long d = (long long)(keyed[(size_t)(o + s0 + l)] & 0xffffffffull);
                        DocDesc &w = work[(size_t)(o + s0 + l)];
                        w.lab0 = lab_ptr[d]; w.len = (int)(doc_ptr[d + 1] - doc_ptr[d]);
                        w.A = (int)(lab_ptr[d + 1] - lab_ptr[d]); w.doc = (int)d;
                        if (sliced) { w.rbase = r_total + l; w.stride = 32; mx = std::max<long long>(mx, w.len); }
                        else        { w.rbase = r_total; w.stride = 1; r_total += w.len; }
                    }
                    if (sliced) r_total += 32 * mx;
                }
            }
        }
    }

    // ---- device buffers (re-used when the handle is loaded again)
    const size_t tab = (size_t)h->desc.V * h->ldk;
    const size_t Nn = (size_t)std::max<long long>(h->N, 1);
    TRY(reserve(h, h->doc_ptr, (size_t)D + 1));
    TRY(reserve(h, h->lab_ptr, (size_t)D + 1));
    TRY(reserve(h, h->lab_idx, (size_t)h->n_lab));
    TRY(reserve(h, h->rec, (size_t)std::max<long long>(r_total, 1)));
    TRY(reserve(h, h->n_wk, tab));
    if (h->desc.mode == GIBBS_MODE_SNAPSHOT) TRY(reserve(h, h->delta_wk, tab));
    TRY(reserve(h, h->n_k, (size_t)h->desc.K));
    TRY(reserve(h, h->n_dk_act, (size_t)h->n_lab));
    TRY(reserve(h, h->colsum, (size_t)h->ldk));
    TRY(reserve(h, h->counters, 4));
    TRY(reserve(h, h->err_flag, 1));
    TRY(reserve(h, h->work, (size_t)std::max<long long>(D, 1)));
    TRY(reserve(h, h->scratch, 3 * Nn * sizeof(int)));
    CK(cudaMemcpyAsync(h->work.p, work.data(), sizeof(DocDesc) * (size_t)D, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->doc_ptr.p, doc_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->lab_ptr.p, lab_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice, h->stream));
    if (h->n_lab) CK(cudaMemcpyAsync(h->lab_idx.p, lab_idx, sizeof(int) * h->n_lab, cudaMemcpyHostToDevice, h->stream));
    if (seg) {
        TRY(reserve(h, h->seg, (size_t)2 * D));
        CK(cudaMemcpyAsync(h->seg.p, seg, sizeof(int) * 2 * D, cudaMemcpyHostToDevice, h->stream));
    }
    int *t_word = reinterpret_cast<int *>(h->scratch.p), *t_freq = nullptr, *t_z = nullptr;
    if (h->N) CK(cudaMemcpyAsync(t_word, word, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    if (freq && h->N) {
        t_freq = t_word + Nn;
        CK(cudaMemcpyAsync(t_freq, freq, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    }
    if (z_init && h->N) {
        t_z = t_word + 2 * Nn;
        CK(cudaMemcpyAsync(t_z, z_init, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
    const uint2 key = make_uint2((uint32_t)h->desc.seed, (uint32_t)(h->desc.seed >> 32));
    if (D > 0) {
        const long long blocks = (D * 32 + 255) / 256;
        prepare_records_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->work.p, h->doc_ptr.p, t_word, t_freq, t_z,
                                                                        h->lab_idx.p, h->rec.p, h->desc.V, key,
                                                                        h->desc.doc_base, h->err_flag.p);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));   // `work` (host) goes out of scope
    int err = 0;
    CK(cudaMemcpy(&err, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(GIBBS_E_ARG, "gibbs_load: inconsistent corpus (word id out of range, f outside 0..65535, or z not in the document's label list)");

    TRY(rebuild_counts(h));
    CK(cudaStreamSynchronize(h->stream));
    h->loaded = true;
    h->sweep = 0;
    h->hat_phi_set = h->hat_theta_set = false;
    h->st.ldk = h->ldk;
    h->st.max_active = max_a;
    // byte models: record 8 B read + 4 B z write-back, two 4-byte RMWs on the delta table, the row fetch,
    // and per-document terms (doc_list 4, doc_ptr/lab_ptr 16, per label: id 4 + n_dk in 4 + n_dk out 4 + n_k 4)
    {
        const double nd = D > 0 ? std::max((double)h->N / (double)D, 1.0) : 1.0;
        const double abar = D > 0 ? (double)h->n_lab / (double)D : 1.0;
        const double per_doc = (20.0 + 16.0 * abar) / nd;
        h->st.bytes_per_draw_dense = 12.0 + 4.0 * h->row_ints + 16.0 + per_doc;
        h->st.bytes_per_draw_gather = 12.0 + 32.0 * abar + 16.0 + per_doc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- sweeps
// Short label lists (bin 0): thread-per-document kernel unless the dense row fetch is forced.
static bool bin_uses_lane(const gibbs_handle *h) {
    return h->desc.mode == GIBBS_MODE_SNAPSHOT && h->desc.row_fetch != GIBBS_FETCH_DENSE;
}
static bool bin_uses_gather(const gibbs_handle *h, int bin) {
    if (bin >= GATHER_BINS || h->desc.mode != GIBBS_MODE_SNAPSHOT) return false;
    if (h->desc.row_fetch == GIBBS_FETCH_DENSE) return false;
    if (bin == 0 || h->desc.row_fetch == GIBBS_FETCH_GATHER) return true;
    // auto: the gather touches <= BIN_CAP sectors of 32 B per draw; the dense path streams row_ints * 4 B through
    // shared memory
    return 32 * BIN_CAP[bin] <= 4 * h->row_ints;
}

template <int G, int NCH, int R, bool SERIAL>
static int launch_dense(gibbs_handle *h, const SweepParams &p) {
    auto kern = llda_dense_kernel<G, NCH, R, SERIAL>;
    const size_t grp_bytes = (size_t)(2 * R) * 8 + (size_t)R * p.row_ints * 4;
    const int gpw = 32 / G;
    const size_t warp_bytes = grp_bytes * gpw;
    int wpc = (int)std::min<size_t>(8, std::max<size_t>(1, (72 * 1024) / warp_bytes));
    const size_t smem = warp_bytes * wpc;
    if (smem > 227 * 1024) return -100;   // caller retries with a shallower ring / wider group
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem));
    if (occ < 1) return -100;
    const long long ctas_needed = (p.n_work + (long long)wpc * gpw - 1) / ((long long)wpc * gpw);
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, (long long)occ * h->sm_count);
    kern<<<grid, wpc * 32, smem, h->stream>>>(p);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

template <int G, int NCH, bool SERIAL>
static int launch_dense_r(gibbs_handle *h, const SweepParams &p) {
    int r = launch_dense<G, NCH, 4, SERIAL>(h, p);
    if (r == -100) r = launch_dense<G, NCH, 2, SERIAL>(h, p);
    return r;
}

static int launch_lane(gibbs_handle *h, const SweepParams &p) {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, llda_lane_kernel, 128, 0));
    if (occ < 1) return fail(GIBBS_E_CUDA, "gibbs_sweep: lane kernel does not fit on an SM");
    const long long slices = (p.n_work + 31) / 32;
    const long long ctas_needed = (slices + 3) / 4;
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, (long long)occ * h->sm_count);
    llda_lane_kernel<<<grid, 128, 0, h->stream>>>(p);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

template <int G, int R>
static int launch_gather(gibbs_handle *h, const SweepParams &p) {
    auto kern = llda_gather_kernel<G, R>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) return fail(GIBBS_E_CUDA, "gibbs_sweep: gather kernel does not fit on an SM");
    const long long groups_per_cta = 256 / G;
    const long long ctas_needed = (p.n_work + groups_per_cta - 1) / groups_per_cta;
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, (long long)occ * h->sm_count);
    kern<<<grid, 256, 0, h->stream>>>(p);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

static int sample_block(gibbs_handle *h, int block) {
    if (h->desc.mode == GIBBS_MODE_EXACT) {
        ExactParams p{};
        p.desc = h->work.p; p.lab_idx = h->lab_idx.p; p.n_dk_act = h->n_dk_act.p;
        p.R = h->rec.p; p.n_wk = h->n_wk.p; p.n_k = h->n_k.p; p.d_begin = 0; p.d_end = h->desc.D; p.ldk = h->ldk;
        p.alpha = h->desc.alpha; p.beta = h->desc.beta; p.vbeta = (double)h->desc.V * h->desc.beta;
        p.seed_lo = (uint32_t)h->desc.seed; p.seed_hi = (uint32_t)(h->desc.seed >> 32); p.sweep = h->sweep;
        p.doc_base = h->desc.doc_base; p.changed = h->counters.p + 1;
        llda_exact_kernel<<<1, 32, 0, h->stream>>>(p);
        CK(cudaGetLastError());
        h->st.last_launches++;
        return 0;
    }
    for (int bin = 0; bin < N_BINS; ++bin) {
        const DocList &dl = h->lists[(size_t)block * N_BINS + bin];
        if (!dl.len) continue;
        SweepParams p{};
        p.work = h->work.p + dl.off; p.n_work = dl.len;
        p.lab_idx = h->lab_idx.p; p.n_dk_act = h->n_dk_act.p;
        p.R = h->rec.p; p.n_wk = h->n_wk.p; p.delta_wk = h->delta_wk.p; p.n_k = h->n_k.p;
        p.seg = h->has_seg ? h->seg.p : nullptr;
        p.counter = h->counters.p; p.changed = h->counters.p + 1;
        p.ldk = h->ldk; p.row_ints = h->row_ints;
        p.alpha = (float)h->desc.alpha; p.beta = (float)h->desc.beta; p.vbeta = (float)((double)h->desc.V * h->desc.beta);
        p.seed_lo = (uint32_t)h->desc.seed; p.seed_hi = (uint32_t)(h->desc.seed >> 32); p.sweep = h->sweep;
        p.doc_base = h->desc.doc_base;
        CK(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long), h->stream));
        if (bin == 0 && bin_uses_lane(h)) { TRY(launch_lane(h, p)); continue; }
        if (bin_uses_gather(h, bin)) {
            if (bin == 1) TRY((launch_gather<16, 4>(h, p)));
            else          TRY((launch_gather<32, 4>(h, p)));
            continue;
        }
        // dense-row path: the ring is per group, so fall back to wider groups (fewer rings per warp) for long rows
        int r = -100;
        if (bin == 0) {
            r = launch_dense_r<8, 1, true>(h, p);
            if (r == -100) r = launch_dense_r<16, 1, true>(h, p);
            if (r == -100) r = launch_dense_r<32, 1, true>(h, p);
        } else if (bin == 1) {
            r = launch_dense_r<16, 1, false>(h, p);
            if (r == -100) r = launch_dense_r<32, 1, false>(h, p);
        } else if (bin == 2) r = launch_dense_r<32, 1, false>(h, p);
        else if (bin == 3) r = launch_dense_r<32, 2, false>(h, p);
        else if (bin == 4) r = launch_dense_r<32, 4, false>(h, p);
        else if (bin == 5) r = launch_dense_r<32, 8, false>(h, p);
        else r = launch_dense_r<32, 16, false>(h, p);
        if (r == -100) return fail(GIBBS_E_ARG, "gibbs_sweep: n_wk row segment too long for the shared-memory ring (K > ~28000)");
        TRY(r);
    }
    return 0;
}

static int merge_block(gibbs_handle *h) {
    if (h->desc.mode != GIBBS_MODE_SNAPSHOT) return 0;
    const size_t tab = (size_t)h->desc.V * h->ldk;
    if (h->comm) NCK(nccl_dl::all_reduce_i32(h->delta_wk.p, tab, h->comm, h->stream));
    const int ldk4 = h->ldk / 4;
    int bx = std::min(256, (ldk4 + 31) / 32 * 32);
    int by = std::max(1, 256 / bx);
    dim3 block(bx, by);
    unsigned grid = (unsigned)std::min<long long>((h->desc.V + by * 4 - 1) / (by * 4), (long long)h->sm_count * 4);
    merge_delta_kernel<<<grid, block, sizeof(int) * (size_t)h->ldk, h->stream>>>(reinterpret_cast<int4 *>(h->n_wk.p), reinterpret_cast<int4 *>(h->delta_wk.p),
                                                      h->n_k.p, h->desc.V, ldk4, h->desc.K);
    CK(cudaGetLastError());
    h->st.last_launches++;
    return 0;
}

// n_sweeps sweeps on the handle's stream, no host synchronisation in between.  Events: [0] call start, [1] call end,
// then (start, mid, end) per refresh block of the LAST sweep only.
extern "C" int gibbs_sweep(gibbs_t *h, int32_t n_sweeps) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_sweep: corpus not loaded");
    if (n_sweeps < 0) return fail(GIBBS_E_ARG, "gibbs_sweep: n_sweeps < 0");
    if (n_sweeps == 0) return 0;
    CK(cudaSetDevice(h->desc.device));
    const int nb = h->desc.n_refresh;
    if (!get_event(h, 2 + 3 * (size_t)nb)) return fail(GIBBS_E_CUDA, "gibbs_sweep: cudaEventCreate failed");
    h->st.last_launches = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    for (int s = 0; s < n_sweeps; ++s) {
        const bool last = s == n_sweeps - 1;
        CK(cudaMemsetAsync(h->counters.p + 1, 0, sizeof(unsigned long long), h->stream));
        for (int b = 0; b < nb; ++b) {
            if (last) CK(cudaEventRecord(h->ev[2 + 3 * b], h->stream));
            TRY(sample_block(h, b));
            if (last) CK(cudaEventRecord(h->ev[3 + 3 * b], h->stream));
            TRY(merge_block(h));
            if (last) CK(cudaEventRecord(h->ev[4 + 3 * b], h->stream));
        }
        h->sweep++;
    }
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaMemcpyAsync(&h->st.changed, h->counters.p + 1, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->st.last_call_ms = ms;
    double samp = 0, mrg = 0;
    for (int b = 0; b < nb; ++b) {
        float a = 0, m = 0;
        CK(cudaEventElapsedTime(&a, h->ev[2 + 3 * b], h->ev[3 + 3 * b]));
        CK(cudaEventElapsedTime(&m, h->ev[3 + 3 * b], h->ev[4 + 3 * b]));
        samp += a; mrg += m;
    }
    h->st.last_sweep_ms = samp;
    h->st.last_merge_ms = mrg;
    h->st.last_launches /= n_sweeps;
    h->st.sweeps += n_sweeps;
    h->st.draws += (long long)n_sweeps * h->N;
    return 0;
}

extern "C" int gibbs_stream(gibbs_t *h, void **stream) {
    if (!h || !stream) return fail(GIBBS_E_ARG, "gibbs_stream: null argument");
    *stream = (void *)h->stream;
    return 0;
}

extern "C" int gibbs_set_sweep_counter(gibbs_t *h, uint32_t sweep) {
    if (!h) return fail(GIBBS_E_ARG, "gibbs_set_sweep_counter: null handle");
    h->sweep = sweep;
    return 0;
}

// ---------------------------------------------------------------------------------------------- state out / in
extern "C" int gibbs_get_state(gibbs_t *h, int32_t *z, int32_t *n_wk, int32_t *n_dk_act, int32_t *n_k) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_get_state: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    if (z && h->N > 0) {
        TRY(reserve(h, h->scratch, (size_t)h->N * sizeof(int)));
        int *t_z = reinterpret_cast<int *>(h->scratch.p);
        const long long blocks = (D * 32 + 255) / 256;
        export_z_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->work.p, h->doc_ptr.p, h->lab_idx.p, h->rec.p, t_z);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(z, t_z, sizeof(int) * h->N, cudaMemcpyDeviceToHost, h->stream));
    }
    if (n_wk)
        CK(cudaMemcpy2DAsync(n_wk, sizeof(int) * h->desc.K, h->n_wk.p, sizeof(int) * h->ldk, sizeof(int) * h->desc.K,
                             (size_t)h->desc.V, cudaMemcpyDeviceToHost, h->stream));
    if (n_dk_act && h->n_lab) CK(cudaMemcpyAsync(n_dk_act, h->n_dk_act.p, sizeof(int) * h->n_lab, cudaMemcpyDeviceToHost, h->stream));
    if (n_k) CK(cudaMemcpyAsync(n_k, h->n_k.p, sizeof(int) * h->desc.K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_set_z(gibbs_t *h, const int32_t *z) {
    if (!h || !h->loaded || !z) return fail(GIBBS_E_STATE, "gibbs_set_z: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    TRY(reserve(h, h->scratch, (size_t)std::max<long long>(h->N, 1) * sizeof(int)));
    int *t_z = reinterpret_cast<int *>(h->scratch.p);
    CK(cudaMemcpyAsync(t_z, z, sizeof(int) * h->N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), h->stream));
    if (D > 0) {
        const long long blocks = (D * 32 + 255) / 256;
        set_z_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->work.p, h->doc_ptr.p, h->lab_idx.p, t_z, h->rec.p, h->err_flag.p);
        CK(cudaGetLastError());
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (err) { h->loaded = false; return fail(GIBBS_E_ARG, "gibbs_set_z: z not in the document's label list (the handle must be loaded again)"); }
    TRY(rebuild_counts(h));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_add_counts(gibbs_t *h, int64_t n, const int32_t *word, const int32_t *topic, const int32_t *count) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_add_counts: corpus not loaded");
    if (n < 0 || (n > 0 && (!word || !topic || !count))) return fail(GIBBS_E_ARG, "gibbs_add_counts: null argument");
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (word[i] < 0 || word[i] >= h->desc.V || topic[i] < 0 || topic[i] >= h->desc.K)
            return fail(GIBBS_E_ARG, "gibbs_add_counts: word or topic id out of range");
    CK(cudaSetDevice(h->desc.device));
    TRY(reserve(h, h->scratch, 3 * (size_t)n * sizeof(int)));
    int *t = reinterpret_cast<int *>(h->scratch.p);
    CK(cudaMemcpyAsync(t, word, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(t + n, topic, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(t + 2 * n, count, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    add_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, t, t + n, t + 2 * n, h->ldk, h->n_wk.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------- outputs
extern "C" int gibbs_emit_phi(gibbs_t *h, double *phi_KV, int32_t smoothed) {
    if (!h || !h->loaded || !phi_KV) return fail(GIBBS_E_STATE, "gibbs_emit_phi: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const int K = h->desc.K, V = h->desc.V;
    TRY(reserve(h, h->scratch, sizeof(double) * (size_t)K * V));
    double *d_phi = reinterpret_cast<double *>(h->scratch.p);
    const int *den = h->n_k.p;
    if (!smoothed) {
        // CascadeLDA.py:394-395 / HSLDA.py:151-152 divide by the row sums of n_k_v itself (they differ from n_zk when
        // the table carries SubLDA's spurious initial counts), so sum the table's columns here
        TRY(column_sums(h, h->colsum.p));
        den = h->colsum.p;
    }
    dim3 grid((V + 31) / 32, (K + 31) / 32), block(32, 8);
    emit_phi_kernel<<<grid, block, 0, h->stream>>>(h->n_wk.p, den, d_phi, V, K, h->ldk, h->desc.beta,
                                                   (double)V * h->desc.beta, smoothed, 0, 0.0, 0.0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(phi_KV, d_phi, sizeof(double) * (size_t)K * V, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_emit_theta(gibbs_t *h, double *theta_DK, int32_t smoothed) {
    if (!h || !h->loaded || !theta_DK) return fail(GIBBS_E_STATE, "gibbs_emit_theta: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    const int K = h->desc.K;
    if (D == 0) return 0;
    TRY(reserve(h, h->scratch, sizeof(double) * (size_t)D * K));
    double *d_th = reinterpret_cast<double *>(h->scratch.p);
    const long long blocks = (D * 32 + 255) / 256;
    emit_theta_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->lab_ptr.p, h->lab_idx.p, h->n_dk_act.p, d_th, K, h->desc.alpha, smoothed);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(theta_DK, d_th, sizeof(double) * (size_t)D * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_emit_theta_csr(gibbs_t *h, double *theta_act, int32_t smoothed) {
    if (!h || !h->loaded || !theta_act) return fail(GIBBS_E_STATE, "gibbs_emit_theta_csr: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    if (D == 0 || h->n_lab == 0) return 0;
    TRY(reserve(h, h->scratch, sizeof(double) * (size_t)h->n_lab));
    double *d_th = reinterpret_cast<double *>(h->scratch.p);
    const long long blocks = (D * 32 + 255) / 256;
    emit_theta_csr_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->lab_ptr.p, h->n_dk_act.p, d_th, h->desc.alpha, smoothed,
                                                                   0, 0.0, 0.0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(theta_act, d_th, sizeof(double) * (size_t)h->n_lab, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------- thinning mean, perplexity
extern "C" int gibbs_thin_accumulate(gibbs_t *h, double c_old, double c_new, int32_t smoothed, int32_t what) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_thin_accumulate: corpus not loaded");
    if (!(what & 3)) return fail(GIBBS_E_ARG, "gibbs_thin_accumulate: `what` selects neither phi (1) nor theta (2)");
    CK(cudaSetDevice(h->desc.device));
    const int K = h->desc.K, V = h->desc.V;
    const long long D = h->desc.D;
    if (what & 1) {
        TRY(reserve(h, h->hat_phi, (size_t)K * V));
        const int *den = h->n_k.p;
        if (!smoothed) { TRY(column_sums(h, h->colsum.p)); den = h->colsum.p; }
        dim3 grid((V + 31) / 32, (K + 31) / 32), block(32, 8);
        emit_phi_kernel<<<grid, block, 0, h->stream>>>(h->n_wk.p, den, h->hat_phi.p, V, K, h->ldk, h->desc.beta,
                                                       (double)V * h->desc.beta, smoothed, h->hat_phi_set ? 1 : 0, c_old, c_new);
        CK(cudaGetLastError());
        h->hat_phi_set = true;
    }
    if ((what & 2) && D > 0 && h->n_lab > 0) {
        TRY(reserve(h, h->hat_theta, (size_t)h->n_lab));
        const long long blocks = (D * 32 + 255) / 256;
        emit_theta_csr_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->lab_ptr.p, h->n_dk_act.p, h->hat_theta.p, h->desc.alpha,
                                                                       smoothed, h->hat_theta_set ? 1 : 0, c_old, c_new);
        CK(cudaGetLastError());
        h->hat_theta_set = true;
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_thin_get(gibbs_t *h, double *ph_hat_KV, double *th_hat_act) {
    if (!h || !h->loaded) return fail(GIBBS_E_STATE, "gibbs_thin_get: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    if (ph_hat_KV) {
        if (!h->hat_phi_set) return fail(GIBBS_E_STATE, "gibbs_thin_get: no phi sample accumulated yet");
        CK(cudaMemcpyAsync(ph_hat_KV, h->hat_phi.p, sizeof(double) * (size_t)h->desc.K * h->desc.V, cudaMemcpyDeviceToHost, h->stream));
    }
    if (th_hat_act && h->n_lab > 0) {
        if (!h->hat_theta_set) return fail(GIBBS_E_STATE, "gibbs_thin_get: no theta sample accumulated yet");
        CK(cudaMemcpyAsync(th_hat_act, h->hat_theta.p, sizeof(double) * (size_t)h->n_lab, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_perplexity(gibbs_t *h, double *neg_log_sum, int64_t *n_pairs) {
    if (!h || !h->loaded || !neg_log_sum) return fail(GIBBS_E_STATE, "gibbs_perplexity: corpus not loaded");
    CK(cudaSetDevice(h->desc.device));
    const long long D = h->desc.D;
    *neg_log_sum = 0.0;
    if (n_pairs) *n_pairs = h->N;
    if (D == 0) return 0;
    TRY(reserve(h, h->doc_sum, (size_t)D + 1));
    const long long blocks = (D * 32 + 255) / 256;
    perplexity_doc_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(D, h->work.p, h->lab_idx.p, h->n_dk_act.p, h->rec.p, h->n_wk.p,
                                                                   h->n_k.p, h->ldk, h->desc.alpha, h->desc.beta,
                                                                   (double)h->desc.V * h->desc.beta, h->doc_sum.p);
    CK(cudaGetLastError());
    sum_f64_kernel<<<1, 1024, 0, h->stream>>>(h->doc_sum.p, D, h->doc_sum.p + D);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(neg_log_sum, h->doc_sum.p + D, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int gibbs_stats(gibbs_t *h, gibbs_stats_t *out) {
    if (!h || !out) return fail(GIBBS_E_ARG, "gibbs_stats: null argument");
    h->st.device_bytes = (int64_t)h->dev_bytes;
    // the row fetch that carries most draws decides which byte model describes the handle
    long long gd = 0, dd = 0;
    for (int b = 0; b < N_BINS; ++b) (bin_uses_gather(h, b) ? gd : dd) += h->bin_draws[b];
    h->st.row_fetch = (h->desc.mode == GIBBS_MODE_SNAPSHOT && gd >= dd) ? GIBBS_FETCH_GATHER : GIBBS_FETCH_DENSE;
    h->st.bytes_per_draw = h->st.row_fetch == GIBBS_FETCH_GATHER ? h->st.bytes_per_draw_gather : h->st.bytes_per_draw_dense;
    *out = h->st;
    return 0;
}

extern "C" int gibbs_trim(gibbs_t *h) {
    if (!h) return fail(GIBBS_E_ARG, "gibbs_trim: null handle");
    CK(cudaSetDevice(h->desc.device));
    CK(cudaStreamSynchronize(h->stream));
    release(h, h->scratch);
    return 0;
}

// ---------------------------------------------------------------------------------------------- HSLDA state
extern "C" int gibbs_hslda_set(gibbs_t *, int32_t, const double *, const double *, const double *, const double *) {
    return fail(GIBBS_E_STATE, "gibbs_hslda_set: not implemented in this build");
}

// ---------------------------------------------------------------------------------------------- test chains
struct gibbs_test_handle {
    int device = 0, K = 0, V = 0, ldk = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    double *phiT = nullptr;
    unsigned char *buf = nullptr;     // per-run staging, grows on demand
    size_t cap = 0;
};

extern "C" int gibbs_test_create(gibbs_test_t **out, int32_t device, int32_t K, int32_t V, const double *phi_KV) {
    if (!out || !phi_KV) return fail(GIBBS_E_ARG, "gibbs_test_create: null argument");
    *out = nullptr;
    if (K <= 0 || V <= 0) return fail(GIBBS_E_ARG, "gibbs_test_create: K and V must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GIBBS_E_CUDA, "gibbs_test_create: no CUDA device; this library has no CPU path");
    if (device < 0 || device >= ndev) return fail(GIBBS_E_ARG, "gibbs_test_create: bad device ordinal");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    gibbs_test_handle *t = new gibbs_test_handle();
    t->device = device; t->K = K; t->V = V; t->ldk = (K + 3) / 4 * 4; t->sm_count = prop.multiProcessorCount;
    double *tmp = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&t->phiT, sizeof(double) * (size_t)V * t->ldk);
    if (e == cudaSuccess) e = cudaMalloc((void **)&tmp, sizeof(double) * (size_t)K * V);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, phi_KV, sizeof(double) * (size_t)K * V, cudaMemcpyHostToDevice, t->stream);
    if (e == cudaSuccess) {
        dim3 grid((V + 31) / 32, (t->ldk + 31) / 32), block(32, 8);
        transpose_phi_kernel<<<grid, block, 0, t->stream>>>(tmp, t->phiT, K, V, t->ldk);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
    if (tmp) cudaFree(tmp);
    if (e != cudaSuccess) {
        if (t->phiT) cudaFree(t->phiT);
        if (t->stream) cudaStreamDestroy(t->stream);
        delete t;
        return cuda_fail(e, "gibbs_test_create", __FILE__, __LINE__);
    }
    *out = t;
    return 0;
}

extern "C" void gibbs_test_destroy(gibbs_test_t *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaStreamSynchronize(t->stream);
    if (t->phiT) cudaFree(t->phiT);
    if (t->buf) cudaFree(t->buf);
    cudaStreamDestroy(t->stream);
    delete t;
}

extern "C" int gibbs_test_run(gibbs_test_t *t, double alpha, double beta_fb, int64_t n_chains, const int64_t *doc_ptr,
                              const int32_t *word, const int32_t *freq, const int64_t *lab_ptr, const int32_t *lab_idx,
                              int32_t *z, int32_t init_mode, int32_t it, int32_t thinning, uint64_t seed,
                              int64_t chain_base, double *th_hat) {
    if (!t || !doc_ptr || !th_hat) return fail(GIBBS_E_ARG, "gibbs_test_run: null argument");
    if (n_chains < 0 || it < 0 || thinning <= 0) return fail(GIBBS_E_ARG, "gibbs_test_run: bad sizes");
    if (init_mode < GIBBS_TEST_INIT_GIVEN || init_mode > GIBBS_TEST_INIT_CASCADE) return fail(GIBBS_E_ARG, "gibbs_test_run: unknown init_mode");
    if (init_mode == GIBBS_TEST_INIT_GIVEN && !z) return fail(GIBBS_E_ARG, "gibbs_test_run: init_mode GIVEN needs z");
    if ((lab_ptr == nullptr) != (lab_idx == nullptr)) return fail(GIBBS_E_ARG, "gibbs_test_run: lab_ptr and lab_idx go together");
    if (n_chains == 0) return 0;
    if (doc_ptr[0] != 0 || (lab_ptr && lab_ptr[0] != 0)) return fail(GIBBS_E_ARG, "gibbs_test_run: CSR offsets must start at 0");
    const long long N = doc_ptr[n_chains];
    if (N > 0 && !word) return fail(GIBBS_E_ARG, "gibbs_test_run: null word array");
    int amax = lab_ptr ? 0 : t->K;
    for (int64_t c = 0; c < n_chains; ++c) {
        if (doc_ptr[c + 1] < doc_ptr[c]) return fail(GIBBS_E_ARG, "gibbs_test_run: CSR offsets must be non-decreasing");
        if (doc_ptr[c + 1] - doc_ptr[c] > 0x7fffffffLL) return fail(GIBBS_E_ARG, "gibbs_test_run: chain too long");
        if (lab_ptr) {
            const long long a = lab_ptr[c + 1] - lab_ptr[c];
            if (a < 1 || a > t->K) return fail(GIBBS_E_ARG, "gibbs_test_run: a chain's topic list must have 1..K entries");
            amax = std::max<int>(amax, (int)a);
        }
    }
    for (long long i = 0; i < N; ++i)
        if (word[i] < 0 || word[i] >= t->V) return fail(GIBBS_E_ARG, "gibbs_test_run: word id out of range");
    const long long n_lab = lab_ptr ? lab_ptr[n_chains] : n_chains * (long long)t->K;
    if (lab_ptr)
        for (long long i = 0; i < n_lab; ++i)
            if (lab_idx[i] < 0 || lab_idx[i] >= t->K) return fail(GIBBS_E_ARG, "gibbs_test_run: topic id out of range");
    CK(cudaSetDevice(t->device));

    // staging layout: doc_ptr | lab_ptr | th_out | word | freq | z | lab_idx | err
    auto up8 = [](size_t x) { return (x + 7) / 8 * 8; };
    size_t off = 0;
    const size_t o_doc = off; off += sizeof(long long) * (size_t)(n_chains + 1);
    const size_t o_labp = off; off += lab_ptr ? sizeof(long long) * (size_t)(n_chains + 1) : 0;
    const size_t o_th = off; off += sizeof(double) * (size_t)std::max<long long>(n_lab, 1);
    const size_t o_word = off; off += up8(sizeof(int) * (size_t)std::max<long long>(N, 1));
    const size_t o_freq = off; off += freq ? up8(sizeof(int) * (size_t)std::max<long long>(N, 1)) : 0;
    const size_t o_z = off; off += up8(sizeof(int) * (size_t)std::max<long long>(N, 1));
    const size_t o_labi = off; off += lab_ptr ? up8(sizeof(int) * (size_t)n_lab) : 0;
    const size_t o_err = off; off += 8;
    if (off > t->cap) {
        if (t->buf) { cudaFree(t->buf); t->buf = nullptr; t->cap = 0; }
        CK(cudaMalloc((void **)&t->buf, off));
        t->cap = off;
    }
    unsigned char *b = t->buf;
    CK(cudaMemcpyAsync(b + o_doc, doc_ptr, sizeof(long long) * (size_t)(n_chains + 1), cudaMemcpyHostToDevice, t->stream));
    if (lab_ptr) {
        CK(cudaMemcpyAsync(b + o_labp, lab_ptr, sizeof(long long) * (size_t)(n_chains + 1), cudaMemcpyHostToDevice, t->stream));
        CK(cudaMemcpyAsync(b + o_labi, lab_idx, sizeof(int) * (size_t)n_lab, cudaMemcpyHostToDevice, t->stream));
    }
    if (N) CK(cudaMemcpyAsync(b + o_word, word, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, t->stream));
    if (N && freq) CK(cudaMemcpyAsync(b + o_freq, freq, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, t->stream));
    if (N && init_mode == GIBBS_TEST_INIT_GIVEN) CK(cudaMemcpyAsync(b + o_z, z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, t->stream));
    CK(cudaMemsetAsync(b + o_th, 0, sizeof(double) * (size_t)std::max<long long>(n_lab, 1), t->stream));
    CK(cudaMemsetAsync(b + o_err, 0, 8, t->stream));

    TestParams p{};
    p.phiT = t->phiT; p.ldk = t->ldk; p.K = t->K; p.n_chains = n_chains;
    p.doc_ptr = reinterpret_cast<const long long *>(b + o_doc);
    p.word = reinterpret_cast<const int *>(b + o_word);
    p.freq = freq ? reinterpret_cast<const int *>(b + o_freq) : nullptr;
    p.lab_ptr = lab_ptr ? reinterpret_cast<const long long *>(b + o_labp) : nullptr;
    p.lab_idx = lab_ptr ? reinterpret_cast<const int *>(b + o_labi) : nullptr;
    p.z = reinterpret_cast<int *>(b + o_z);
    p.th_out = reinterpret_cast<double *>(b + o_th);
    p.alpha = alpha; p.beta_fb = beta_fb; p.it = it; p.thinning = thinning; p.init_mode = init_mode;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.chain_base = chain_base;
    p.a_cap = (amax + 31) / 32 * 32;
    p.err = reinterpret_cast<int *>(b + o_err);
    const size_t per_warp = (size_t)p.a_cap * 16;
    int wpc = (int)std::min<size_t>(8, (96 * 1024) / per_warp);
    if (wpc < 1) {
        wpc = 1;
        if (per_warp > 227 * 1024) return fail(GIBBS_E_ARG, "gibbs_test_run: topic list too long for the shared-memory state (max ~14000 topics)");
    }
    const size_t smem = per_warp * wpc;
    CK(cudaFuncSetAttribute(test_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas = (n_chains + wpc - 1) / wpc;
    const unsigned grid = (unsigned)std::min<long long>(ctas, (long long)t->sm_count * 8);
    test_chain_kernel<<<grid, wpc * 32, smem, t->stream>>>(p);
    CK(cudaGetLastError());
    int err = 0;
    CK(cudaMemcpyAsync(&err, b + o_err, sizeof(int), cudaMemcpyDeviceToHost, t->stream));
    CK(cudaMemcpyAsync(th_hat, b + o_th, sizeof(double) * (size_t)n_lab, cudaMemcpyDeviceToHost, t->stream));
    if (z && N) CK(cudaMemcpyAsync(z, b + o_z, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, t->stream));
    CK(cudaStreamSynchronize(t->stream));
    if (err) return fail(GIBBS_E_ARG, "gibbs_test_run: a z value is not in its chain's topic list");
    return 0;
}

// ---------------------------------------------------------------------------------------------- KAT
extern "C" int gibbs_philox_kat(int32_t device, int32_t n, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4) {
    if (n <= 0 || !ctr4 || !key2 || !out4) return fail(GIBBS_E_ARG, "gibbs_philox_kat: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GIBBS_E_CUDA, "gibbs_philox_kat: no CUDA device");
    CK(cudaSetDevice(device));
    uint32_t *d = nullptr;
    CK(cudaMalloc((void **)&d, sizeof(uint32_t) * 10 * (size_t)n));
    uint32_t *d_c = d, *d_k = d + 4 * (size_t)n, *d_o = d + 6 * (size_t)n;
    cudaError_t e = cudaMemcpy(d_c, ctr4, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_k, key2, sizeof(uint32_t) * 2 * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        philox_kat_kernel<<<(n + 127) / 128, 128>>>(n, d_c, d_k, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out4, d_o, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "gibbs_philox_kat", __FILE__, __LINE__);
    return 0;
}
