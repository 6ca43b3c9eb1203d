// Host-side companion of the bit-packed detection output (rvmc_biascorr_fused_packed): expands the
// packed words into the bool array get_detected_binaries returns (RVMC_bias_corr.py:451-476 hands
// back `detected`, dtype bool, shape (n_stars, n_samples)) with a small persistent thread pool.
//
// Why it exists: one detection flag per binary is 1 byte in the reference's result but 1 BIT of
// information.  Shipping bytes over PCIe makes the device-to-host copy the slowest stage of the
// end-to-end pass (100 MB per 10^8 binaries against a 3 ms kernel); shipping bits and widening them
// where the result is consumed moves 8x less over the link.  No GPU work happens here and nothing of
// the Monte-Carlo arithmetic either: this is the tail end of a copy.
//
// Bit order: bit j of word w of a row is column 32 w + j (what __ballot_sync produces for 32
// consecutive samples).  On a little-endian host the words are a byte stream in which bit j of byte
// k is column 8 k + j, i.e. np.unpackbits(..., bitorder="little").
#include <cuda_runtime.h>
#include <emmintrin.h>   // SSE2: baseline of every x86-64 CPU (streaming stores)
#include <immintrin.h>   // AVX-512 path, selected at run time
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/rvmc_b200.h"

namespace rvmc {
void set_error(const char* fmt, ...);
}

namespace {

// ---- a minimal persistent fork-join pool ----------------------------------------------------------
// parallel_for(n_tasks, n_threads, fn): tasks are claimed from an atomic counter by up to
// n_threads - 1 pool workers plus the calling thread.  One job at a time (callers serialise on
// run_mu); workers sleep on a condition variable between jobs.
class Pool {
  public:
    typedef void (*TaskFn)(void* ctx, int64_t task);

    void run(int64_t n_tasks, int n_threads, TaskFn fn, void* ctx) {
        if (n_tasks <= 0) return;
        if (n_threads > (int)n_tasks) n_threads = (int)n_tasks;
        if (n_threads <= 1) {
            for (int64_t t = 0; t < n_tasks; ++t) fn(ctx, t);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_);
        grow(n_threads - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = fn;
            ctx_ = ctx;
            n_tasks_ = n_tasks;
            next_.store(0, std::memory_order_relaxed);
            wanted_ = n_threads - 1;
            running_ = 0;
            ++generation_;
        }
        cv_job_.notify_all();
        drain(fn, ctx, n_tasks);
        std::unique_lock<std::mutex> lk(mu_);
        wanted_ = 0;                                        // late wakers find nothing to join
        cv_done_.wait(lk, [this] { return running_ == 0; });
    }

    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++generation_;
        }
        cv_job_.notify_all();
        for (auto& t : workers_)
            if (t.joinable()) t.join();
    }

  private:
    void drain(TaskFn fn, void* ctx, int64_t n_tasks) {
        for (;;) {
            const int64_t t = next_.fetch_add(1, std::memory_order_relaxed);
            if (t >= n_tasks) break;
            fn(ctx, t);
        }
    }

    void grow(int n) {
        while ((int)workers_.size() < n) workers_.emplace_back([this] { worker(); });
    }

    void worker() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_job_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            if (stop_) return;
            if (wanted_ <= 0) continue;
            --wanted_;
            ++running_;
            TaskFn fn = fn_;
            void* ctx = ctx_;
            const int64_t n = n_tasks_;
            lk.unlock();
            drain(fn, ctx, n);
            lk.lock();
            if (--running_ == 0) cv_done_.notify_all();
        }
    }

    std::mutex run_mu_, mu_;
    std::condition_variable cv_job_, cv_done_;
    std::vector<std::thread> workers_;
    TaskFn fn_ = nullptr;
    void* ctx_ = nullptr;
    int64_t n_tasks_ = 0;
    std::atomic<int64_t> next_{0};
    int wanted_ = 0, running_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

std::atomic<Pool*> g_pool{nullptr};
std::atomic<bool> g_atfork_registered{false};

void pool_child_after_fork() {
    // the child gets a fresh pool on first use; the parent's object (its mutexes possibly held by
    // threads that do not exist here) is leaked on purpose
    g_pool.store(nullptr, std::memory_order_release);
}

// The pool lives until the process ends (never destroyed: its sleeping workers die with the process).
Pool& pool() {
    Pool* p = g_pool.load(std::memory_order_acquire);
    if (!p) {
        Pool* fresh = new Pool();
        if (g_pool.compare_exchange_strong(p, fresh, std::memory_order_acq_rel)) {
            p = fresh;
            if (!g_atfork_registered.exchange(true)) pthread_atfork(nullptr, nullptr, pool_child_after_fork);
        } else {
            delete fresh;       // another thread won the race; `fresh` never started a worker
        }
    }
    return *p;
}

// ---- bit -> byte expansion --------------------------------------------------------------------------
uint64_t g_lut[256];
std::once_flag g_lut_once;

void build_lut();

// n_cols flags of one row: src = packed bytes (column 0 is bit 0 of src[0]), dst = n_cols bytes of 0/1.
void expand_row(const uint8_t* src, uint8_t* dst, int64_t n_cols) {
    int64_t c = 0;
    const int64_t full = n_cols & ~(int64_t)7;
    // head: up to the first 16-byte boundary of dst, in whole source bytes
    while (c < full && ((uintptr_t)(dst + c) & 15)) {
        const uint64_t v = g_lut[src[c >> 3]];
        memcpy(dst + c, &v, 8);
        c += 8;
    }
    // body: two source bytes -> one 16-byte streaming store (the result is written once and read later
    // by somebody else: no reason to pull the destination lines into this core's cache first)
    for (; c + 16 <= full; c += 16) {
        const __m128i v = _mm_set_epi64x((long long)g_lut[src[(c >> 3) + 1]], (long long)g_lut[src[c >> 3]]);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + c), v);
    }
    for (; c < full; c += 8) {
        const uint64_t v = g_lut[src[c >> 3]];
        memcpy(dst + c, &v, 8);
    }
    if (c < n_cols) {
        const uint64_t v = g_lut[src[c >> 3]];
        memcpy(dst + c, &v, (size_t)(n_cols - c));
    }
}

// The same with AVX-512BW when the CPU has it (checked once at run time): 64 flags per step -- the 64-bit
// mask IS the packed input, one masked byte broadcast widens it, one full-cache-line streaming store writes
// it.  Whole lines leave the core in one piece, which the memory system likes better than four 16-byte
// write-combining fragments when every core of the host is streaming at once.
__attribute__((target("avx512f,avx512bw")))
void expand_row_avx512(const uint8_t* src, uint8_t* dst, int64_t n_cols) {
    if ((uintptr_t)dst & 7) {           // rows that do not start on an 8-byte boundary: the portable path
        expand_row(src, dst, n_cols);
        return;
    }
    int64_t c = 0;
    const int64_t full = n_cols & ~(int64_t)7;
    while (c < full && ((uintptr_t)(dst + c) & 63)) {
        const uint64_t v = g_lut[src[c >> 3]];
        memcpy(dst + c, &v, 8);
        c += 8;
    }
    const __m512i one = _mm512_set1_epi8(1);
    for (; c + 64 <= full; c += 64) {
        uint64_t m;
        memcpy(&m, src + (c >> 3), 8);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + c), _mm512_maskz_mov_epi8((__mmask64)m, one));
    }
    if (c < n_cols) expand_row(src + (c >> 3), dst + c, n_cols - c);
}

typedef void (*ExpandFn)(const uint8_t*, uint8_t*, int64_t);
ExpandFn pick_expand() {
    if (getenv("RVMC_HOST_NO_AVX512")) return expand_row;
    __builtin_cpu_init();
    return (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) ? expand_row_avx512 : expand_row;
}
ExpandFn g_expand = nullptr;

struct UnpackJob {
    const uint8_t* bits;
    int64_t bits_stride_bytes;
    uint8_t* out;
    int64_t out_stride;
    int64_t n_cols;
    int64_t segs_per_row;
    int64_t seg_cols;          // multiple of 64
};

void unpack_task(void* ctx, int64_t task) {
    const UnpackJob& j = *static_cast<const UnpackJob*>(ctx);
    const int64_t row = task / j.segs_per_row;
    const int64_t c0 = (task - row * j.segs_per_row) * j.seg_cols;
    const int64_t n = (j.n_cols - c0 < j.seg_cols) ? j.n_cols - c0 : j.seg_cols;
    g_expand(j.bits + row * j.bits_stride_bytes + (c0 >> 3), j.out + row * j.out_stride + c0, n);
    _mm_sfence();
}

void build_lut() {
    for (int b = 0; b < 256; ++b) {
        uint64_t v = 0;
        for (int j = 0; j < 8; ++j)
            if (b & (1 << j)) v |= (uint64_t)1 << (8 * j);
        g_lut[b] = v;
    }
    g_expand = pick_expand();
}

struct CountJob {
    const uint32_t* bits;
    int64_t bits_stride;
    int64_t n_cols;
    int64_t* counts;
};

void count_task(void* ctx, int64_t row) {
    const CountJob& j = *static_cast<const CountJob*>(ctx);
    const uint32_t* w = j.bits + row * j.bits_stride;
    const int64_t full = j.n_cols >> 5;
    int64_t n = 0;
    for (int64_t k = 0; k < full; ++k) n += __builtin_popcount(w[k]);
    const int rem = (int)(j.n_cols & 31);
    if (rem) n += __builtin_popcount(w[full] & ((1u << rem) - 1u));
    j.counts[row] = n;
}

// ---- grid fit: one pass over the statistics table -------------------------------------------------
// The host side of a sharded grid fit ends with a table of 80-byte rvmc_point_stats records in host
// memory (5 MB for the 65 536-point grid).  Picking out one 8-byte field of every record with NumPy costs a
// strided gather per field (the best-fit rule reads `median`, the sanity checks `reserved` and
// `n_nonconverged`: three gathers, ~0.2 ms each -- 4 % of an 8-GPU grid); this does all of it in one pass.
struct FitPart {
    int64_t idx;
    double dist;
    uint64_t reserved, nonconv, nonconv_points, guard;
};

struct FitJob {
    const rvmc_point_stats* stats;
    int64_t n, seg;
    int use_mode, have_observed;
    double observed;
    FitPart* parts;
};

void fit_task(void* ctx, int64_t task) {
    const FitJob& j = *static_cast<const FitJob*>(ctx);
    const int64_t lo = task * j.seg;
    const int64_t hi = (lo + j.seg < j.n) ? lo + j.seg : j.n;
    FitPart p = {-1, 100.0, 0, 0, 0, 0};
    for (int64_t i = lo; i < hi; ++i) {
        const rvmc_point_stats& s = j.stats[i];
        p.reserved += s.reserved;
        p.nonconv += s.n_nonconverged;
        p.nonconv_points += s.n_nonconverged ? 1 : 0;
        p.guard += s.n_guard;
        if (j.have_observed) {
            // findBestDistributions.py:138-158: strict `<` against the best so far, which starts at 100;
            // a NaN never compares below it
            const double c = j.use_mode ? s.mode : s.median;
            double d = c - j.observed;
            d = d < 0.0 ? -d : d;
            if (d < p.dist) {
                p.dist = d;
                p.idx = i;
            }
        }
    }
    j.parts[task] = p;
}

struct CopyJob {
    uint8_t* dst;
    const uint8_t* src;
    int64_t nbytes, seg;
};

void copy_task(void* ctx, int64_t task) {
    const CopyJob& j = *static_cast<const CopyJob*>(ctx);
    const int64_t lo = task * j.seg;
    const int64_t n = (j.nbytes - lo < j.seg) ? j.nbytes - lo : j.seg;
    memcpy(j.dst + lo, j.src + lo, (size_t)n);
}

}  // namespace

extern "C" {

int rvmc_copy_host(void* dst_host, const void* src_host, int64_t nbytes, int n_threads) {
    if (nbytes < 0 || (nbytes > 0 && (!dst_host || !src_host))) {
        rvmc::set_error("rvmc_copy_host: bad arguments");
        return RVMC_ERR_INVALID;
    }
    if (nbytes == 0) return RVMC_OK;
    CopyJob j;
    j.dst = static_cast<uint8_t*>(dst_host);
    j.src = static_cast<const uint8_t*>(src_host);
    j.nbytes = nbytes;
    j.seg = 1 << 20;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    pool().run((nbytes + j.seg - 1) / j.seg, n_threads, copy_task, &j);
    return RVMC_OK;
}

int rvmc_unpack_bits_host(const uint32_t* bits_host, int64_t bits_stride_words, int64_t n_rows, int64_t n_cols,
                          uint8_t* out_host, int64_t out_stride, int n_threads) {
    if (n_rows < 0 || n_cols < 0) {
        rvmc::set_error("rvmc_unpack_bits_host: negative shape");
        return RVMC_ERR_INVALID;
    }
    if (n_rows == 0 || n_cols == 0) return RVMC_OK;
    if (!bits_host || !out_host) {
        rvmc::set_error("rvmc_unpack_bits_host: bits_host and out_host are required");
        return RVMC_ERR_INVALID;
    }
    if (bits_stride_words < (n_cols + 31) / 32 || out_stride < n_cols) {
        rvmc::set_error("rvmc_unpack_bits_host: a row stride is shorter than the row");
        return RVMC_ERR_INVALID;
    }
    std::call_once(g_lut_once, build_lut);
    UnpackJob j;
    j.bits = reinterpret_cast<const uint8_t*>(bits_host);
    j.bits_stride_bytes = bits_stride_words * 4;
    j.out = out_host;
    j.out_stride = out_stride;
    j.n_cols = n_cols;
    // tasks of at most 256 Ki flags (32 KB in, 256 KB out): fine-grained enough to balance a few
    // threads over one long row, coarse enough that claiming a task costs nothing
    j.seg_cols = 262144;
    j.segs_per_row = (n_cols + j.seg_cols - 1) / j.seg_cols;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    pool().run(n_rows * j.segs_per_row, n_threads, unpack_task, &j);
    return RVMC_OK;
}

// The whole device-to-host leg of the packed flags in ONE call, without a host-language loop in the way:
// for every block (rows x columns of the [n_rows][bits_stride] device array, in the order given) the copy
// stream waits for the block's `ready` event (recorded behind the kernel that produced it), copies its words
// into the pinned mirror `packed_host`, and as copies land the calling thread and the helper pool widen them
// into `out_host` -- consecutive blocks that have landed together go out as one widening call.  Returns when
// every flag is in place.  The GPU keeps computing later blocks meanwhile; nothing here touches the SMs.
int rvmc_fetch_flags_host(const uint32_t* bits_dev, int64_t bits_stride_words, int64_t n_cols,
                          const rvmc_flag_block* blocks, int n_blocks, void* copy_stream, uint32_t* packed_host,
                          uint8_t* out_host, int64_t out_stride, int n_threads) {
    if (n_blocks < 0 || n_cols < 0 || (n_blocks > 0 && (!bits_dev || !blocks || !packed_host || !out_host))) {
        rvmc::set_error("rvmc_fetch_flags_host: bad arguments");
        return RVMC_ERR_INVALID;
    }
    if (bits_stride_words < (n_cols + 31) / 32 || out_stride < n_cols) {
        rvmc::set_error("rvmc_fetch_flags_host: a row stride is shorter than the row");
        return RVMC_ERR_INVALID;
    }
    for (int k = 0; k < n_blocks; ++k) {
        const rvmc_flag_block& b = blocks[k];
        if (b.row0 < 0 || b.row1 < b.row0 || b.col0 < 0 || b.col1 < b.col0 || b.col1 > n_cols || (b.col0 & 31)) {
            rvmc::set_error("rvmc_fetch_flags_host: block %d is malformed (columns must start on a multiple of 32)", k);
            return RVMC_ERR_INVALID;
        }
    }
    cudaStream_t cs = static_cast<cudaStream_t>(copy_stream);
    std::vector<cudaEvent_t> done((size_t)n_blocks, nullptr);
    cudaError_t err = cudaSuccess;
    auto fail = [&](const char* what) {
        rvmc::set_error("rvmc_fetch_flags_host: %s: %s", what, cudaGetErrorString(err));
        cudaStreamSynchronize(cs);                    // nothing of ours may still be in flight into packed_host
        for (cudaEvent_t e : done)
            if (e) cudaEventDestroy(e);
        cudaGetLastError();
        return (int)err;
    };
    for (int k = 0; k < n_blocks; ++k) {
        const rvmc_flag_block& b = blocks[k];
        if (b.ready_event && (err = cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(b.ready_event), 0)) != cudaSuccess)
            return fail("cudaStreamWaitEvent");
        const int64_t w0 = b.col0 >> 5, w1 = (b.col1 + 31) >> 5, rows = b.row1 - b.row0;
        if (rows > 0 && w1 > w0) {
            const size_t off = (size_t)b.row0 * bits_stride_words + w0;
            if (w0 == 0 && w1 == bits_stride_words)
                err = cudaMemcpyAsync(packed_host + off, bits_dev + off, (size_t)rows * bits_stride_words * 4,
                                      cudaMemcpyDeviceToHost, cs);
            else
                err = cudaMemcpy2DAsync(packed_host + off, (size_t)bits_stride_words * 4, bits_dev + off,
                                        (size_t)bits_stride_words * 4, (size_t)(w1 - w0) * 4, (size_t)rows,
                                        cudaMemcpyDeviceToHost, cs);
            if (err != cudaSuccess) return fail("cudaMemcpyAsync");
        }
        if ((err = cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate");
        if ((err = cudaEventRecord(done[k], cs)) != cudaSuccess) return fail("cudaEventRecord");
    }
    int rc = RVMC_OK;
    for (int k = 0; k < n_blocks && rc == RVMC_OK;) {
        if ((err = cudaEventSynchronize(done[k])) != cudaSuccess) return fail("cudaEventSynchronize");
        int64_t r0 = blocks[k].row0, r1 = blocks[k].row1, c0 = blocks[k].col0, c1 = blocks[k].col1;
        int m = k + 1;
        // the GPU usually runs ahead of the widening: take every following block that has also landed and
        // continues this one (next rows of the same columns, or next columns of the same rows)
        while (m < n_blocks && cudaEventQuery(done[m]) == cudaSuccess) {
            const rvmc_flag_block& q = blocks[m];
            if (q.col0 == c0 && q.col1 == c1 && q.row0 == r1) r1 = q.row1;
            else if (q.row0 == r0 && q.row1 == r1 && q.col0 == c1) c1 = q.col1;
            else break;
            ++m;
        }
        cudaGetLastError();                              // cudaEventQuery leaves cudaErrorNotReady behind
        if (r1 > r0 && c1 > c0)
            rc = rvmc_unpack_bits_host(packed_host + (size_t)r0 * bits_stride_words + (c0 >> 5), bits_stride_words,
                                       r1 - r0, c1 - c0, out_host + (size_t)r0 * out_stride + c0, out_stride, n_threads);
        k = m;
    }
    for (cudaEvent_t e : done)
        if (e) cudaEventDestroy(e);
    return rc;
}

int rvmc_count_bits_host(const uint32_t* bits_host, int64_t bits_stride_words, int64_t n_rows, int64_t n_cols,
                         int64_t* counts_host, int n_threads) {
    if (n_rows < 0 || n_cols < 0) {
        rvmc::set_error("rvmc_count_bits_host: negative shape");
        return RVMC_ERR_INVALID;
    }
    if (n_rows == 0) return RVMC_OK;
    if (!bits_host || !counts_host) {
        rvmc::set_error("rvmc_count_bits_host: bits_host and counts_host are required");
        return RVMC_ERR_INVALID;
    }
    if (bits_stride_words < (n_cols + 31) / 32) {
        rvmc::set_error("rvmc_count_bits_host: the row stride is shorter than the row");
        return RVMC_ERR_INVALID;
    }
    CountJob j;
    j.bits = bits_host;
    j.bits_stride = bits_stride_words;
    j.n_cols = n_cols;
    j.counts = counts_host;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    pool().run(n_rows, n_threads, count_task, &j);
    return RVMC_OK;
}

int rvmc_grid_best_fit_host(const rvmc_point_stats* stats_host, int64_t n_points, int use_mode, int have_observed,
                            double observed, int64_t* best_index, double* best_distance, uint64_t* totals,
                            int n_threads) {
    if (n_points < 0 || (n_points > 0 && !stats_host)) {
        rvmc::set_error("rvmc_grid_best_fit_host: bad arguments");
        return RVMC_ERR_INVALID;
    }
    FitJob j;
    j.stats = stats_host;
    j.n = n_points;
    j.seg = 4096;
    j.use_mode = use_mode;
    j.have_observed = have_observed;
    j.observed = observed;
    const int64_t n_tasks = (n_points + j.seg - 1) / j.seg;
    std::vector<FitPart> parts((size_t)(n_tasks > 0 ? n_tasks : 1));
    j.parts = parts.data();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    pool().run(n_tasks, n_threads, fit_task, &j);
    FitPart all = {-1, 100.0, 0, 0, 0, 0};
    for (int64_t t = 0; t < n_tasks; ++t) {        // in table order: the FIRST strict minimum wins
        const FitPart& p = parts[(size_t)t];
        all.reserved += p.reserved;
        all.nonconv += p.nonconv;
        all.nonconv_points += p.nonconv_points;
        all.guard += p.guard;
        if (p.idx >= 0 && p.dist < all.dist) {
            all.dist = p.dist;
            all.idx = p.idx;
        }
    }
    if (best_index) *best_index = all.idx < 0 ? 0 : all.idx;          // nothing below 100: index 0, distance 100
    if (best_distance) *best_distance = all.dist;
    if (totals) {
        totals[0] = all.reserved;
        totals[1] = all.nonconv;
        totals[2] = all.nonconv_points;
        totals[3] = all.guard;
    }
    return RVMC_OK;
}

}  // extern "C"
