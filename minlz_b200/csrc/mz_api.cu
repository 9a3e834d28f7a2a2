// mz_api.cu -- C ABI of libminlz_cuda.so (see include/minlz_cuda.h).
//
// Host side of the drop-in boundary: argument checking, block-header handling
// (reference encode.go:74-139, decode.go:50-171), device workspace pooling,
// launches.  All compression / decompression work happens in the CUDA kernels;
// there is no CPU codec in this library.
#include "../../include/minlz_cuda.h"

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "mz_decode_pc.cuh"
#include "mz_encode_l1.cuh"
#include "mz_encode_l2.cuh"
#include "mz_pack.cuh"
#include "mz_crc32c.cuh"
#include "mz_validate.cuh"

#include <condition_variable>
#include <map>
#include <thread>
#include <climits>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace {

// Encoder flavour (see minlz_cuda.h): process-wide, like the build tag it stands for.
std::atomic<int> g_flavor{MZCU_FLAVOR_GO};

// Validate mode (mzcu_set_validate / MZCU_VALIDATE=1): decode-after-encode on the device,
// the reference's debugValidateBlocks (minlz.go:52, encode.go:108-133).  -1 = read the env once.
std::atomic<int> g_validate{-1};
bool validate_on() {
    int v = g_validate.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("MZCU_VALIDATE");
        v = (e && *e && *e != '0') ? 1 : 0;
        g_validate.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

thread_local char g_err[512] = "";
thread_local float g_last_kernel_ms = 0.f;

// Launch gate of the asynchronous calls (mzcu_submit_*): the FIRST kernel launch of every submitted
// job happens in submission order, so "decode of batch k, then encode of batch k+1" reaches the
// device in that order whatever the job threads' start-up timing (see DeviceState::dec_done).
std::mutex g_gate_mu;
std::condition_variable g_gate_cv;
int64_t g_gate_next = 1;             // ticket whose turn it is
thread_local int64_t t_ticket = 0;   // this thread's job ticket; 0 = not a submitted job, or already through
void gate_enter() {
    if (!t_ticket) return;
    std::unique_lock<std::mutex> lk(g_gate_mu);
    g_gate_cv.wait(lk, [] { return g_gate_next == t_ticket; });
}
void gate_leave() {
    if (!t_ticket) return;
    {
        std::lock_guard<std::mutex> lk(g_gate_mu);
        g_gate_next = t_ticket + 1;
    }
    t_ticket = 0;
    g_gate_cv.notify_all();
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) return fail(MZCU_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// ---- per-device state ------------------------------------------------------
constexpr int kMaxDevices = 64;
constexpr int kCounterSlots = 1024;
constexpr int kDefaultL2Fetch = 0;  // bytes; 0 = leave the device default

// Scratch for one encode launch: the per-warp match tables.  A slice may be
// reused once the launch that used it has finished (event query).
struct TableWs {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
};

struct DeviceState {
    std::once_flag once;
    cudaError_t init_err = cudaSuccess;
    int num_sms = 0;
    int enc_l1_ctas_per_sm = 1;
    int enc_l2_ctas_per_sm = 1;
    int *counters = nullptr;  // kCounterSlots ints
    mz::CrcTables *crc_tabs = nullptr;
    std::atomic<unsigned> next_counter{0};
    std::mutex ws_mu;
    std::vector<TableWs> table_ws;
    // Launch order across concurrent host calls: an encode launch is a persistent kernel that holds
    // every SM for ~100 ms, a decode launch is short but its download is long.  A decode kernel that
    // queues behind an encode kernel delays that download by the whole encode; so an encode launch
    // waits for the decode kernels already submitted on this device (<= one decode's duration) and
    // the download then overlaps the encode.
    cudaEvent_t dec_done = nullptr;
    std::mutex order_mu;
};
DeviceState g_dev[kMaxDevices];

int resolve_device(int device) {
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return -1;
    }
    return (device >= 0 && device < kMaxDevices) ? device : -1;
}

// CRC-32C (Castagnoli, reflected) byte table and the "advance by 2^k zero
// bytes" matrices used to combine per-lane partial checksums.
void build_crc_tables(mz::CrcTables *t) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
        t->byte_table[i] = c;
    }
    // one zero byte: c -> table[c & 0xff] ^ (c >> 8), applied to each unit vector
    for (int i = 0; i < 32; i++) {
        uint32_t c = 1u << i;
        t->zeros[0][i] = t->byte_table[c & 0xff] ^ (c >> 8);
    }
    for (int k = 1; k < 32; k++)  // square the operator
        for (int i = 0; i < 32; i++) {
            uint32_t v = t->zeros[k - 1][i], r = 0;
            for (int b = 0; b < 32; b++)
                if ((v >> b) & 1) r ^= t->zeros[k - 1][b];
            t->zeros[k][i] = r;
        }
}

int init_device(int device) {
    DeviceState &st = g_dev[device];
    std::call_once(st.once, [&] {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) {
            // The encoders probe 32-byte table slots at random: ask L2 to fetch 32 B per miss instead of
            // promoting every miss to a wider DRAM read (a hint; MINLZ_CUDA_L2_FETCH=0 leaves it alone).
            const char *g = getenv("MINLZ_CUDA_L2_FETCH");
            const int gran = g ? atoi(g) : kDefaultL2Fetch;
            if (gran > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran) != cudaSuccess)
                cudaGetLastError();
        }
        if (e == cudaSuccess) {
            // experiment knob: L2 set-aside for persisting (evict_last) lines, in MiB
            const char *g = getenv("MINLZ_CUDA_L2_PERSIST_MB");
            if (g && atoi(g) > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)atoi(g) << 20) != cudaSuccess)
                cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st.dec_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&st.counters, kCounterSlots * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&st.crc_tabs, sizeof(mz::CrcTables));
        if (e == cudaSuccess) {
            mz::CrcTables *h = new mz::CrcTables();
            build_crc_tables(h);
            e = cudaMemcpy(st.crc_tabs, h, sizeof(mz::CrcTables), cudaMemcpyHostToDevice);
            delete h;
        }
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st.enc_l1_ctas_per_sm, mz::encode_l1_kernel<false>,
                                                              mz::kEncL1Warps * 32, 0);
        if (st.enc_l1_ctas_per_sm < 1) st.enc_l1_ctas_per_sm = 1;
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st.enc_l2_ctas_per_sm, mz::encode_l2_kernel,
                                                              mz::kEncL2Warps * 32, 0);
        if (st.enc_l2_ctas_per_sm < 1) st.enc_l2_ctas_per_sm = 1;
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mz::decode_pc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)mz::kDecSmemBytes);
        st.init_err = e;
    });
    if (st.init_err != cudaSuccess)
        return fail(MZCU_ERR_CUDA, "device %d init: %s", device, cudaGetErrorString(st.init_err));
    CU_TRY(cudaSetDevice(device));
    return MZCU_OK;
}

// ---- launches --------------------------------------------------------------
int launch_decode(int device, int nblk, const uint8_t *src, const uint64_t *sbeg, const uint64_t *send, uint8_t *dst,
                  const uint64_t *dbeg, const uint64_t *dend, int32_t *status, cudaStream_t stream, int batch_blocks = 0,
                  bool record_done = true) {
    if (nblk == 0) return MZCU_OK;
    // up to 32 blocks per CTA (one parser lane each), spread over all SMs.  When this
    // launch is one chunk of a larger batch whose chunks run concurrently, size the
    // CTAs for the whole batch so that the chunks share the SMs instead of each
    // spreading thinly over all of them.
    const int sms = g_dev[device].num_sms;
    const int whole = batch_blocks > nblk ? batch_blocks : nblk;
    int slots = (whole + sms - 1) / sms;
    if (slots > mz::kDecSlots) slots = mz::kDecSlots;
    if (slots < 1) slots = 1;
    const int grid = (nblk + slots - 1) / slots;
    mz::decode_pc_kernel<<<grid, mz::kDecThreads, mz::kDecSmemBytes, stream>>>(nblk, slots, src, sbeg, send, dst, dbeg,
                                                                               dend, status);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess && record_done) {
        std::lock_guard<std::mutex> lk(g_dev[device].order_mu);
        le = cudaEventRecord(g_dev[device].dec_done, stream);
    }
    if (le != cudaSuccess) return fail(MZCU_ERR_CUDA, "decode launch: %s", cudaGetErrorString(le));
    return MZCU_OK;
}

int acquire_tables(DeviceState &st, size_t bytes, TableWs *out) {
    {
        std::lock_guard<std::mutex> lk(st.ws_mu);
        for (size_t i = 0; i < st.table_ws.size(); i++) {
            TableWs &w = st.table_ws[i];
            if (w.bytes >= bytes && cudaEventQuery(w.done) == cudaSuccess) {
                *out = w;
                st.table_ws.erase(st.table_ws.begin() + i);
                return MZCU_OK;
            }
        }
        cudaGetLastError();  // clear cudaErrorNotReady from the queries
    }
    TableWs w;
    CU_TRY(cudaMalloc(&w.ptr, bytes));
    w.bytes = bytes;
    CU_TRY(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    *out = w;
    return MZCU_OK;
}

void release_tables(DeviceState &st, const TableWs &w) {
    std::lock_guard<std::mutex> lk(st.ws_mu);
    st.table_ws.push_back(w);
}

int launch_encode(int device, int level, int nblk, const uint8_t *src, const uint64_t *sbeg, const uint64_t *send,
                  uint8_t *dst, const uint64_t *dbeg, uint32_t *out_len, cudaStream_t stream, const int *gate = nullptr,
                  int slice = 0) {
    if (nblk == 0) return MZCU_OK;
    DeviceState &st = g_dev[device];
    int *counter = st.counters + (st.next_counter.fetch_add(1) % kCounterSlots);
    CU_TRY(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    gate_enter();
    {
        std::lock_guard<std::mutex> lk(st.order_mu);
        cudaError_t we = cudaStreamWaitEvent(stream, st.dec_done, 0);  // short decode kernels first (see DeviceState)
        if (we != cudaSuccess) {
            gate_leave();
            return fail(MZCU_ERR_CUDA, "encode launch order: %s", cudaGetErrorString(we));
        }
    }
    struct GateGuard {
        ~GateGuard() { gate_leave(); }
    } gate_guard;  // released after the launch below (or on any early return)
    if (level == MZCU_LEVEL_FASTEST || level == MZCU_LEVEL_SUPERFAST) {
        // one block per warp, all resident: grid = min(blocks, what fits on the chip)
        int grid = (nblk + mz::kEncL1Warps - 1) / mz::kEncL1Warps;
        int resident = st.num_sms * st.enc_l1_ctas_per_sm;
        if (grid > resident) grid = resident;
        TableWs ws;
        int rc = acquire_tables(st, (size_t)grid * mz::kEncL1Warps * mz::kEncL1WsBytesPerWarp, &ws);
        if (rc) return rc;
        const bool amd64 = g_flavor.load(std::memory_order_relaxed) == MZCU_FLAVOR_AMD64;
        if (level == MZCU_LEVEL_FASTEST && !amd64)
            mz::encode_l1_kernel<false><<<grid, mz::kEncL1Warps * 32, 0, stream>>>(
                nblk, src, sbeg, send, dst, dbeg, out_len, counter, static_cast<mz::Slot *>(ws.ptr), gate, slice);
        else if (level == MZCU_LEVEL_FASTEST)
            mz::encode_l1_asm_kernel<false><<<grid, mz::kEncL1Warps * 32, 0, stream>>>(
                nblk, src, sbeg, send, dst, dbeg, out_len, counter, static_cast<mz::Slot *>(ws.ptr), gate, slice);
        else if (!amd64)
            mz::encode_l1_kernel<true><<<grid, mz::kEncL1Warps * 32, 0, stream>>>(
                nblk, src, sbeg, send, dst, dbeg, out_len, counter, static_cast<mz::Slot *>(ws.ptr), gate, slice);
        else
            mz::encode_l1_asm_kernel<true><<<grid, mz::kEncL1Warps * 32, 0, stream>>>(
                nblk, src, sbeg, send, dst, dbeg, out_len, counter, static_cast<mz::Slot *>(ws.ptr), gate, slice);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ws.done, stream);
        release_tables(st, ws);
        if (e != cudaSuccess) return fail(MZCU_ERR_CUDA, "encode_l1 launch: %s", cudaGetErrorString(e));
    } else if (level == MZCU_LEVEL_BALANCED) {
        if (gate) return fail(MZCU_ERR_INVALID_ARG, "encode_l2 has no arrival gate");
        int grid = (nblk + mz::kEncL2Warps - 1) / mz::kEncL2Warps;
        int resident = st.num_sms * st.enc_l2_ctas_per_sm;
        if (grid > resident) grid = resident;
        TableWs ws;
        int rc = acquire_tables(st, (size_t)grid * mz::kEncL2Warps * mz::kEncL2WsBytesPerWarp, &ws);
        if (rc) return rc;
        if (g_flavor.load(std::memory_order_relaxed) == MZCU_FLAVOR_AMD64)
            mz::encode_l2_asm_kernel<<<grid, mz::kEncL2Warps * 32, 0, stream>>>(nblk, src, sbeg, send, dst, dbeg, out_len,
                                                                               counter, static_cast<uint32_t *>(ws.ptr));
        else
            mz::encode_l2_kernel<<<grid, mz::kEncL2Warps * 32, 0, stream>>>(nblk, src, sbeg, send, dst, dbeg, out_len, counter,
                                                                           static_cast<uint32_t *>(ws.ptr));
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ws.done, stream);
        release_tables(st, ws);
        if (e != cudaSuccess) return fail(MZCU_ERR_CUDA, "encode_l2 launch: %s", cudaGetErrorString(e));
    } else {
        return fail(MZCU_ERR_INVALID_LEVEL, "invalid level %d", level);
    }
    CU_TRY(cudaGetLastError());
    return MZCU_OK;
}

int launch_crc(int device, int nblk, const uint8_t *src, const uint64_t *sbeg, const uint64_t *send, uint32_t *crc,
               cudaStream_t stream) {
    if (nblk == 0) return MZCU_OK;
    const int grid = (nblk + mz::kCrcWarps - 1) / mz::kCrcWarps;
    mz::crc32c_blocks_kernel<<<grid, mz::kCrcWarps * 32, 0, stream>>>(nblk, src, sbeg, send, crc, g_dev[device].crc_tabs);
    CU_TRY(cudaGetLastError());
    return MZCU_OK;
}

int launch_pack(int device, int nblk, const uint8_t *src, const uint64_t *sbeg, const uint32_t *len, uint8_t *dst,
                uint64_t *off, cudaStream_t stream) {
    if (nblk == 0) return MZCU_OK;
    mz::scan_lengths_kernel<<<1, 1024, 0, stream>>>(nblk, len, off);
    CU_TRY(cudaGetLastError());
    int grid = g_dev[device].num_sms * 8;
    if (grid > nblk) grid = nblk;
    mz::pack_blocks_kernel<<<grid, 256, 0, stream>>>(nblk, src, sbeg, len, dst, off);
    CU_TRY(cudaGetLastError());
    return MZCU_OK;
}

// Validate mode: decodes what the encoder wrote (token streams in their slots: dst + dbeg[i],
// out_len[i] bytes) with the product decode kernel into a scratch image of the source layout and
// compares.  Synchronises the stream; a debugging aid, like debugValidateBlocks in the reference.
int validate_encoded(int device, int nblk, const uint8_t *d_src, const uint64_t *d_sbeg, const uint64_t *d_send,
                     const uint8_t *d_enc, const uint64_t *d_dbeg, const uint32_t *d_out_len, cudaStream_t stream) {
    if (nblk == 0) return MZCU_OK;
    uint64_t span = 0;
    CU_TRY(cudaMemcpyAsync(&span, d_send + (nblk - 1), sizeof span, cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    uint8_t *d_dec = nullptr;
    uint64_t *d_t = nullptr;
    int32_t *d_status = nullptr;
    int *d_first = nullptr;
    auto cleanup = [&] {
        cudaFree(d_dec);
        cudaFree(d_t);
        cudaFree(d_status);
        cudaFree(d_first);
    };
    cudaError_t e = cudaMalloc(&d_dec, span + 64);
    if (e == cudaSuccess) e = cudaMalloc(&d_t, 2 * (size_t)nblk * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_status, (size_t)nblk * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_first, sizeof(int));
    const int none = INT_MAX;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_first, &none, sizeof(int), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
        cleanup();
        return fail(MZCU_ERR_CUDA, "validate scratch: %s", cudaGetErrorString(e));
    }
    mz::validate_ranges_kernel<<<(nblk + 255) / 256, 256, 0, stream>>>(nblk, d_dbeg, d_out_len, d_t, d_t + nblk);
    int rc = launch_decode(device, nblk, d_enc, d_t, d_t + nblk, d_dec, d_sbeg, d_send, d_status, stream);
    int first = none;
    if (rc == MZCU_OK) {
        int grid = g_dev[device].num_sms * 4;
        if (grid > nblk) grid = nblk;
        mz::validate_compare_kernel<<<grid, 256, 0, stream>>>(nblk, d_src, d_sbeg, d_send, d_dec, d_out_len, d_status, d_first);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&first, d_first, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) rc = fail(MZCU_ERR_CUDA, "validate: %s", cudaGetErrorString(e));
    }
    cleanup();
    if (rc) return rc;
    if (first != none) return fail(MZCU_ERR_VALIDATE, "validate: block %d does not decode back to its source", first);
    return MZCU_OK;
}

constexpr int kMaxChunks = 16;
// chunks per host call; MZCU_CHUNKS overrides (tuning knob)
int max_chunks() {
    static const int v = [] {
        int c = 8;
        if (const char *e = getenv("MZCU_CHUNKS")) c = atoi(e);
        return c < 1 ? 1 : c > kMaxChunks ? kMaxChunks : c;
    }();
    return v;
}
constexpr int kMinChunkBlocks = 256;
// slice of the sliced source copy (0 disables); MZCU_SLICE_KB overrides (tuning knob)
int slice_bytes() {
    static const int v = [] {
        int kb = 32;
        if (const char *e = getenv("MZCU_SLICE_KB")) kb = atoi(e);
        return kb <= 0 ? 0 : kb << 10;
    }();
    return v;
}  // do not split batches below this many blocks per chunk

// ---- pooled host-call workspaces ------------------------------------------
struct Workspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    size_t cap_src = 0, cap_dst = 0;
    uint64_t *d_tab = nullptr;  // 4 offset arrays + out/status
    size_t cap_tab = 0;         // in blocks
    uint64_t *h_tab = nullptr;  // pinned mirror of d_tab
    cudaStream_t cs[kMaxChunks] = {};  // chunk pipelines (copy in / kernels / copy out overlap across chunks)
    cudaEvent_t tab_ready = nullptr;
    cudaEvent_t copied = nullptr;  // sliced upload: the last slice has landed
    cudaEvent_t chunk_ev[kMaxChunks] = {};  // decode: the chunk's kernel is done
    int *d_arrived = nullptr;  // arrival gate of the sliced host->device source copy
    int *h_slice_no = nullptr; // pinned 1, 2, 3, ... (source of the gate writes)
};
constexpr int kMaxSlices = 1024;

std::mutex g_ws_mu;
std::vector<Workspace *> g_ws_free;

int ws_acquire(int device, Workspace **out) {
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        for (size_t i = 0; i < g_ws_free.size(); i++) {
            if (g_ws_free[i]->device == device) {
                *out = g_ws_free[i];
                g_ws_free.erase(g_ws_free.begin() + i);
                return MZCU_OK;
            }
        }
    }
    Workspace *w = new Workspace();
    w->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev1);
    for (int c = 0; c < kMaxChunks && e == cudaSuccess; c++) e = cudaStreamCreateWithFlags(&w->cs[c], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->tab_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->copied, cudaEventDisableTiming);
    for (int c = 0; c < kMaxChunks && e == cudaSuccess; c++) e = cudaEventCreateWithFlags(&w->chunk_ev[c], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_arrived, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_slice_no, kMaxSlices * sizeof(int));
    if (e == cudaSuccess)
        for (int i = 0; i < kMaxSlices; i++) w->h_slice_no[i] = i + 1;
    if (e != cudaSuccess) {
        delete w;
        return fail(MZCU_ERR_CUDA, "workspace: %s", cudaGetErrorString(e));
    }
    *out = w;
    return MZCU_OK;
}

void ws_release(Workspace *w) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    g_ws_free.push_back(w);
}

int ws_reserve(Workspace *w, size_t src_bytes, size_t dst_bytes, size_t nblk) {
    auto grow = [](uint8_t **p, size_t *cap, size_t need) -> cudaError_t {
        if (need <= *cap) return cudaSuccess;
        if (*p) cudaFree(*p);
        *p = nullptr;
        *cap = 0;
        size_t want = need + need / 4 + 256;
        cudaError_t e = cudaMalloc(p, want);
        if (e == cudaSuccess) *cap = want;
        return e;
    };
    CU_TRY(grow(&w->d_src, &w->cap_src, src_bytes + 64));
    CU_TRY(grow(&w->d_dst, &w->cap_dst, dst_bytes + 64));
    if (nblk + 1 > w->cap_tab) {
        if (w->d_tab) cudaFree(w->d_tab);
        if (w->h_tab) cudaFreeHost(w->h_tab);
        w->d_tab = nullptr;
        w->h_tab = nullptr;
        w->cap_tab = 0;
        size_t want = (nblk + 1) * 2;
        CU_TRY(cudaMalloc(&w->d_tab, want * 5 * sizeof(uint64_t)));
        CU_TRY(cudaMallocHost(&w->h_tab, want * 5 * sizeof(uint64_t)));
        w->cap_tab = want;
    }
    return MZCU_OK;
}

struct WsGuard {
    Workspace *w = nullptr;
    ~WsGuard() {
        if (w) ws_release(w);
    }
};

// uvarint helpers (encoding/binary semantics)
int put_uvarint(uint8_t *dst, uint64_t x) {
    int i = 0;
    while (x >= 0x80) {
        dst[i++] = (uint8_t)x | 0x80;
        x >>= 7;
    }
    dst[i] = (uint8_t)x;
    return i + 1;
}

int get_uvarint(const uint8_t *buf, size_t len, uint64_t *out) {
    uint64_t x = 0;
    unsigned sft = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t b = buf[i];
        if (i == 10) return -(int)(i + 1);
        if (b < 0x80) {
            if (i == 9 && b > 1) return -(int)(i + 1);
            *out = x | (uint64_t)b << sft;
            return (int)i + 1;
        }
        x |= (uint64_t)(b & 0x7f) << sft;
        sft += 7;
    }
    *out = 0;
    return 0;
}

// decode.go:120-156 isMinLZ.  hdr = offset of the payload in the block.
int parse_block_header(const uint8_t *src, size_t n, int *is_mlz, int *lits, size_t *hdr, int64_t *size) {
    *is_mlz = 0;
    *lits = 0;
    *hdr = 0;
    *size = 0;
    if (n <= 1) {
        if (n == 0) return MZCU_ERR_CORRUPT;
        if (src[0] == 0) {
            *is_mlz = 1;
            *lits = 1;
            *hdr = 1;
            return MZCU_OK;
        }
    }
    uint64_t v;
    if (src[0] != 0) {
        int k = get_uvarint(src, n, &v);
        if (k <= 0 || v > 0xffffffffull) return MZCU_ERR_CORRUPT;
        *size = (int64_t)v;
        return MZCU_OK;
    }
    int k = get_uvarint(src + 1, n - 1, &v);
    if (k <= 0 || v > 0xffffffffull) return MZCU_ERR_CORRUPT;
    if (v > MZCU_MAX_BLOCK_SIZE) return MZCU_ERR_TOO_LARGE;
    size_t rest = n - 1 - (size_t)k;
    if (rest == 0) return MZCU_ERR_CORRUPT;
    *hdr = 1 + (size_t)k;
    if (v == 0) {
        *is_mlz = 1;
        *lits = 1;
        *size = (int64_t)rest;
        return MZCU_OK;
    }
    *size = (int64_t)v;
    if (v < rest) return MZCU_ERR_CORRUPT;
    *is_mlz = 1;
    return MZCU_OK;
}

// encode.go:223-229 encodeUncompressed
int64_t store_uncompressed(uint8_t *dst, size_t cap, const uint8_t *src, size_t n) {
    if (n == 0) {
        if (cap < 1) return fail(MZCU_ERR_DST_TOO_SMALL, "dst too small");
        dst[0] = 0;
        return 1;
    }
    if (cap < n + 2) return fail(MZCU_ERR_DST_TOO_SMALL, "dst too small");
    dst[0] = 0;
    dst[1] = 0;
    memcpy(dst + 2, src, n);
    return (int64_t)n + 2;
}

// Runs the seam-level encode for host buffers.  Blocks are packed densely on
// the device (16-byte aligned starts); results land in the caller's layout.
int host_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                       const uint64_t *dst_off, uint32_t *out_len) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !out_len))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (level != MZCU_LEVEL_SUPERFAST && level != MZCU_LEVEL_FASTEST && level != MZCU_LEVEL_BALANCED)
        return fail(MZCU_ERR_INVALID_LEVEL, "invalid level %d", level);
    if (nblk == 0) return MZCU_OK;
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    WsGuard g;
    rc = ws_acquire(device, &g.w);
    if (rc) return rc;
    Workspace *w = g.w;

    size_t src_bytes = 0, dst_bytes = 0;
    for (int i = 0; i < nblk; i++) {
        if (src_off[i + 1] < src_off[i] || dst_off[i + 1] < dst_off[i]) return fail(MZCU_ERR_INVALID_ARG, "offsets not monotonic");
        size_t n = src_off[i + 1] - src_off[i];
        if (n > MZCU_MAX_BLOCK_SIZE) return fail(MZCU_ERR_TOO_LARGE, "block %d larger than 8 MiB", i);
        if (n >= 16 && dst_off[i + 1] - dst_off[i] < n + 2) return fail(MZCU_ERR_DST_TOO_SMALL, "block %d: dst capacity < MaxEncodedLen", i);
        dst_bytes += (n + 2 + 15) & ~size_t(15);
    }
    src_bytes = src_off[nblk] - src_off[0];
    rc = ws_reserve(w, src_bytes, dst_bytes, (size_t)nblk);
    if (rc) return rc;
    const size_t T = w->cap_tab;
    uint64_t *h_sbeg = w->h_tab, *h_send = w->h_tab + T, *h_dbeg = w->h_tab + 2 * T;
    uint64_t *d_sbeg = w->d_tab, *d_send = w->d_tab + T, *d_dbeg = w->d_tab + 2 * T;
    uint32_t *d_out = reinterpret_cast<uint32_t *>(w->d_tab + 4 * T);
    uint32_t *h_out = reinterpret_cast<uint32_t *>(w->h_tab + 4 * T);
    // The device copy mirrors the caller's layout (kernels are alignment
    // agnostic), so the whole batch moves with one H2D.
    size_t dof = 0;
    {
        const size_t base = src_off[0];
        const size_t total = src_off[nblk] - base;
        for (int i = 0; i < nblk; i++) {
            h_sbeg[i] = src_off[i] - base;
            h_send[i] = src_off[i + 1] - base;
        }
        if (total) CU_TRY(cudaMemcpyAsync(w->d_src, src + base, total, cudaMemcpyHostToDevice, w->stream));
    }
    for (int i = 0; i < nblk; i++) {
        size_t n = src_off[i + 1] - src_off[i];
        h_dbeg[i] = dof;
        dof += (n + 2 + 15) & ~size_t(15);
    }
    CU_TRY(cudaMemcpyAsync(w->d_tab, w->h_tab, 3 * T * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream));
    CU_TRY(cudaEventRecord(w->ev0, w->stream));
    rc = launch_encode(device, level, nblk, w->d_src, d_sbeg, d_send, w->d_dst, d_dbeg, d_out, w->stream);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(w->ev1, w->stream));
    if (validate_on()) {
        rc = validate_encoded(device, nblk, w->d_src, d_sbeg, d_send, w->d_dst, d_dbeg, d_out, w->stream);
        if (rc) return rc;
    }
    CU_TRY(cudaMemcpyAsync(h_out, d_out, (size_t)nblk * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    CU_TRY(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < nblk; i++) {
        out_len[i] = h_out[i];
        if (h_out[i])
            CU_TRY(cudaMemcpyAsync(dst + dst_off[i], w->d_dst + h_dbeg[i], h_out[i], cudaMemcpyDeviceToHost, w->stream));
    }
    CU_TRY(cudaStreamSynchronize(w->stream));
    cudaEventElapsedTime(&g_last_kernel_ms, w->ev0, w->ev1);
    return MZCU_OK;
}

// Seam-level encode for host buffers with dense output: the token streams are
// packed back to back on the device and leave with one D2H copy.
int host_encode_blocks_packed(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                              size_t dst_cap, uint64_t *dst_off_out, uint32_t *crc_out) {
    if (nblk < 0 || !dst_off_out || (nblk > 0 && (!src_off || !dst))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (level != MZCU_LEVEL_SUPERFAST && level != MZCU_LEVEL_FASTEST && level != MZCU_LEVEL_BALANCED)
        return fail(MZCU_ERR_INVALID_LEVEL, "invalid level %d", level);
    dst_off_out[0] = 0;
    if (nblk == 0) return MZCU_OK;
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    WsGuard g;
    rc = ws_acquire(device, &g.w);
    if (rc) return rc;
    Workspace *w = g.w;
    size_t dst_bytes = 0;
    for (int i = 0; i < nblk; i++) {
        if (src_off[i + 1] < src_off[i]) return fail(MZCU_ERR_INVALID_ARG, "offsets not monotonic");
        size_t n = src_off[i + 1] - src_off[i];
        if (n > MZCU_MAX_BLOCK_SIZE) return fail(MZCU_ERR_TOO_LARGE, "block %d larger than 8 MiB", i);
        dst_bytes += (n + 2 + 15) & ~size_t(15);
    }
    const size_t base = src_off[0];
    const size_t total = src_off[nblk] - base;
    rc = ws_reserve(w, total, dst_bytes, (size_t)nblk);
    if (rc) return rc;
    const size_t T = w->cap_tab;
    uint64_t *h_sbeg = w->h_tab, *h_send = w->h_tab + T, *h_dbeg = w->h_tab + 2 * T, *h_poff = w->h_tab + 3 * T;
    uint64_t *d_sbeg = w->d_tab, *d_send = w->d_tab + T, *d_dbeg = w->d_tab + 2 * T, *d_poff = w->d_tab + 3 * T;
    uint32_t *d_out = reinterpret_cast<uint32_t *>(w->d_tab + 4 * T);
    uint32_t *d_crc = d_out + T;  // second half of the 5th table row
    // checksums travel through the pinned mirror: a D2H into the caller's (possibly pageable) array
    // would block the host until the kernels before it are done and serialise the chunk pipelines
    uint32_t *h_crc = reinterpret_cast<uint32_t *>(w->h_tab + 4 * T) + T;
    size_t dof = 0;
    for (int i = 0; i < nblk; i++) {
        h_sbeg[i] = src_off[i] - base;
        h_send[i] = src_off[i + 1] - base;
        h_dbeg[i] = dof;
        dof += (src_off[i + 1] - src_off[i] + 2 + 15) & ~size_t(15);
    }
    // Sliced pipeline (LevelFastest / LevelSuperFast, equal-sized contiguous blocks): the
    // walk of a block is a serial chain that eats ~10 MB/s, PCIe delivers ~50 GB/s, so
    // instead of waiting for whole blocks the source goes up in SLICES -- bytes
    // [k*S, (k+1)*S) of every block, one strided copy -- each followed by a 4-byte
    // write of k+1 to the arrival gate, while ONE persistent encode launch already
    // runs and chases the arrival front (gate_wait in mz_encode_l1.cuh).  The call
    // then costs max(kernel, copy) instead of copy + kernel.
    {
        const size_t B = src_off[1] - src_off[0];
        bool uniform = level != MZCU_LEVEL_BALANCED && nblk >= 64 && B >= (256u << 10) && B % 128 == 0 && slice_bytes() > 0 &&
                       !validate_on();
        for (int i = 1; uniform && i < nblk; i++) {
            const size_t n = src_off[i + 1] - src_off[i];
            uniform = i + 1 < nblk ? n == B : (n <= B && n > 0);
        }
        if (uniform) {
            const size_t S = (size_t)slice_bytes();
            const int nsl = (int)((B + S - 1) / S);
            const size_t last = src_off[nblk] - src_off[nblk - 1];
            const int full = last == B ? nblk : nblk - 1;  // rows of the strided copies
            if (nsl <= kMaxSlices) {
                cudaStream_t cp = w->cs[0];
                CU_TRY(cudaMemsetAsync(w->d_arrived, 0, sizeof(int), w->stream));
                CU_TRY(cudaMemcpyAsync(w->d_tab, w->h_tab, 3 * T * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream));
                CU_TRY(cudaEventRecord(w->tab_ready, w->stream));
                CU_TRY(cudaEventRecord(w->ev0, w->stream));
                rc = launch_encode(device, level, nblk, w->d_src, d_sbeg, d_send, w->d_dst, d_dbeg, d_out, w->stream,
                                   w->d_arrived, (int)S);
                if (rc) return rc;
                CU_TRY(cudaStreamWaitEvent(cp, w->tab_ready, 0));  // the gate is zeroed before the first write
                cudaError_t ce = cudaSuccess;
                for (int k = 0; k < nsl && ce == cudaSuccess; k++) {
                    const size_t o = (size_t)k * S, wd = o + S <= B ? S : B - o;
                    ce = cudaMemcpy2DAsync(w->d_src + o, B, src + base + o, B, wd, (size_t)full, cudaMemcpyHostToDevice, cp);
                    if (ce == cudaSuccess && full < nblk && o < last)
                        ce = cudaMemcpyAsync(w->d_src + (size_t)full * B + o, src + base + (size_t)full * B + o,
                                             o + wd <= last ? wd : last - o, cudaMemcpyHostToDevice, cp);
                    if (ce == cudaSuccess)
                        ce = cudaMemcpyAsync(w->d_arrived, w->h_slice_no + k, sizeof(int), cudaMemcpyHostToDevice, cp);
                }
                if (ce != cudaSuccess) {
                    // never leave the kernel waiting for slices that will not come
                    cudaMemcpyAsync(w->d_arrived, w->h_slice_no + kMaxSlices - 1, sizeof(int), cudaMemcpyHostToDevice, cp);
                    cudaStreamSynchronize(cp);
                    cudaStreamSynchronize(w->stream);
                    return fail(MZCU_ERR_CUDA, "sliced copy: %s", cudaGetErrorString(ce));
                }
                // The encode kernel only waits for the bytes it reads: a block that bails out early
                // (incompressible) lets the launch finish while later slices are still in flight.
                // The checksum reads, and the pack overwrites, d_src: both must wait for the copy.
                CU_TRY(cudaEventRecord(w->copied, cp));
                CU_TRY(cudaStreamWaitEvent(w->stream, w->copied, 0));
                if (crc_out) {  // checksum of the uncompressed blocks (writer.go:672)
                    rc = launch_crc(device, nblk, w->d_src, d_sbeg, d_send, d_crc, w->stream);
                    if (rc == MZCU_OK)
                        CU_TRY(cudaMemcpyAsync(h_crc, d_crc, (size_t)nblk * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                               w->stream));
                }
                if (rc == MZCU_OK) rc = launch_pack(device, nblk, w->d_dst, d_dbeg, d_out, w->d_src, d_poff, w->stream);
                if (rc == MZCU_OK)
                    CU_TRY(cudaMemcpyAsync(h_poff, d_poff, (size_t)(nblk + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                           w->stream));
                CU_TRY(cudaStreamSynchronize(cp));
                CU_TRY(cudaStreamSynchronize(w->stream));
                if (rc) return rc;
                const size_t packed = h_poff[nblk];
                if (packed > dst_cap) return fail(MZCU_ERR_DST_TOO_SMALL, "packed output exceeds capacity %zu", dst_cap);
                if (packed) CU_TRY(cudaMemcpyAsync(dst, w->d_src, packed, cudaMemcpyDeviceToHost, w->stream));
                for (int i = 0; i <= nblk; i++) dst_off_out[i] = h_poff[i];
                if (crc_out) memcpy(crc_out, h_crc, (size_t)nblk * sizeof(uint32_t));  // landed before the sync above
                CU_TRY(cudaEventRecord(w->ev1, w->stream));
                CU_TRY(cudaStreamSynchronize(w->stream));
                cudaEventElapsedTime(&g_last_kernel_ms, w->ev0, w->ev1);
                return MZCU_OK;
            }
        }
    }

    // Chunk pipeline: the batch is cut into up to kMaxChunks runs of blocks, each on its
    // own stream (H2D -> crc -> encode -> pack -> D2H of the sizes).  The kernels are
    // latency bound per block, so chunks overlap on the device while later chunks are
    // still arriving over PCIe, and packed chunks leave while others still encode.
    int nchunks = nblk / kMinChunkBlocks;
    if (nchunks > max_chunks()) nchunks = max_chunks();
    if (nchunks < 1) nchunks = 1;
    int first[kMaxChunks + 1];
    for (int c = 0; c <= nchunks; c++) first[c] = (int)((int64_t)nblk * c / nchunks);
    CU_TRY(cudaMemcpyAsync(w->d_tab, w->h_tab, 3 * T * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream));
    CU_TRY(cudaEventRecord(w->tab_ready, w->stream));
    CU_TRY(cudaEventRecord(w->ev0, w->stream));
    for (int c = 0; c < nchunks; c++) {
        cudaStream_t cs = w->cs[c];
        const int f = first[c], m = first[c + 1] - first[c];
        const size_t cb = h_sbeg[f], ce = h_send[first[c + 1] - 1];
        CU_TRY(cudaStreamWaitEvent(cs, w->tab_ready, 0));
        if (ce > cb) CU_TRY(cudaMemcpyAsync(w->d_src + cb, src + base + cb, ce - cb, cudaMemcpyHostToDevice, cs));
        if (crc_out) {  // checksum of the uncompressed blocks while they are resident (writer.go:672)
            rc = launch_crc(device, m, w->d_src, d_sbeg + f, d_send + f, d_crc + f, cs);
            if (rc) return rc;
        }
        rc = launch_encode(device, level, m, w->d_src, d_sbeg + f, d_send + f, w->d_dst, d_dbeg + f, d_out + f, cs);
        if (rc) return rc;
        if (validate_on()) {
            rc = validate_encoded(device, m, w->d_src, d_sbeg + f, d_send + f, w->d_dst, d_dbeg + f, d_out + f, cs);
            if (rc) return rc;
        }
        // the chunk's source is dead after its encode: pack into its own source range
        // (sum(len) < chunk bytes); the chunk's offsets start at 0 (own poff slice, stride m+1)
        rc = launch_pack(device, m, w->d_dst, d_dbeg + f, d_out + f, w->d_src + cb, d_poff + f + c, cs);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(h_poff + f + c, d_poff + f + c, (size_t)(m + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, cs));
        if (crc_out) CU_TRY(cudaMemcpyAsync(h_crc + f, d_crc + f, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    }
    size_t host_pos = 0;
    for (int c = 0; c < nchunks; c++) {
        cudaStream_t cs = w->cs[c];
        const int f = first[c], m = first[c + 1] - first[c];
        CU_TRY(cudaStreamSynchronize(cs));
        const uint64_t *po = h_poff + f + c;
        const size_t packed = po[m];
        if (host_pos + packed > dst_cap) {
            for (int k = 0; k < nchunks; k++) cudaStreamSynchronize(w->cs[k]);
            return fail(MZCU_ERR_DST_TOO_SMALL, "packed output exceeds capacity %zu", dst_cap);
        }
        if (packed) CU_TRY(cudaMemcpyAsync(dst + host_pos, w->d_src + h_sbeg[f], packed, cudaMemcpyDeviceToHost, cs));
        for (int i = 0; i <= m; i++) dst_off_out[f + i] = host_pos + po[i];
        host_pos += packed;
    }
    for (int c = 0; c < nchunks; c++) {
        CU_TRY(cudaStreamSynchronize(w->cs[c]));
        CU_TRY(cudaStreamWaitEvent(w->stream, w->tab_ready, 0));
    }
    CU_TRY(cudaEventRecord(w->ev1, w->stream));
    CU_TRY(cudaStreamSynchronize(w->stream));
    if (crc_out) memcpy(crc_out, h_crc, (size_t)nblk * sizeof(uint32_t));
    cudaEventElapsedTime(&g_last_kernel_ms, w->ev0, w->ev1);
    return MZCU_OK;
}

// Seam-level decode for host buffers.  `sbeg/send` index into `src`.
int host_decode_ranges(int device, int nblk, const uint8_t *src, const uint64_t *sbeg, const uint64_t *send, uint8_t *dst,
                       const uint64_t *dbeg, const uint64_t *dend, int32_t *status, uint32_t *crc_out = nullptr) {
    if (nblk == 0) return MZCU_OK;
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    WsGuard g;
    rc = ws_acquire(device, &g.w);
    if (rc) return rc;
    Workspace *w = g.w;
    size_t dst_bytes = 0;
    uint64_t lo = ~0ull, hi = 0;
    bool dense = true;
    for (int i = 0; i < nblk; i++) {
        if (send[i] < sbeg[i] || dend[i] < dbeg[i]) return fail(MZCU_ERR_INVALID_ARG, "offsets not monotonic");
        if (dend[i] - dbeg[i] > MZCU_MAX_BLOCK_SIZE) return fail(MZCU_ERR_TOO_LARGE, "block %d larger than 8 MiB", i);
        if (send[i] > sbeg[i]) {
            if (sbeg[i] < lo) lo = sbeg[i];
            if (send[i] > hi) hi = send[i];
        }
        if (i + 1 < nblk && dend[i] != dbeg[i + 1]) dense = false;
        dst_bytes += (dend[i] - dbeg[i] + 15) & ~size_t(15);
    }
    if (hi < lo) lo = hi = 0;
    rc = ws_reserve(w, hi - lo, dst_bytes, (size_t)nblk);
    if (rc) return rc;
    const size_t T = w->cap_tab;
    uint64_t *h = w->h_tab;
    size_t dof = 0;
    for (int i = 0; i < nblk; i++) {
        size_t m = dend[i] - dbeg[i];
        h[i] = send[i] > sbeg[i] ? sbeg[i] - lo : 0;
        h[T + i] = send[i] > sbeg[i] ? send[i] - lo : 0;
        h[2 * T + i] = dense ? dbeg[i] - dbeg[0] : dof;
        h[3 * T + i] = h[2 * T + i] + m;
        dof += (m + 15) & ~size_t(15);
    }
    int32_t *d_status = reinterpret_cast<int32_t *>(w->d_tab + 4 * T);
    int32_t *h_status = reinterpret_cast<int32_t *>(w->h_tab + 4 * T);
    uint32_t *d_crc = reinterpret_cast<uint32_t *>(d_status) + T;
    uint32_t *h_crc = reinterpret_cast<uint32_t *>(h_status) + T;  // pinned mirror (see host_encode_blocks_packed)
    // Chunk pipeline (see host_encode_blocks_packed): copy in, decode, checksum and copy
    // out run per chunk on their own streams.  Source ranges must be ascending for the
    // per-chunk H2D; otherwise the batch goes as one chunk.
    int nchunks = nblk / kMinChunkBlocks;
    if (nchunks > max_chunks()) nchunks = max_chunks();
    if (nchunks < 1) nchunks = 1;
    bool ascending = true;
    for (int i = 0; i + 1 < nblk; i++)
        if (h[T + i] > h[i + 1] && h[T + i + 1] > h[i + 1]) ascending = false;
    if (!ascending) nchunks = 1;
    int first[kMaxChunks + 1];
    for (int c = 0; c <= nchunks; c++) first[c] = (int)((int64_t)nblk * c / nchunks);
    CU_TRY(cudaMemcpyAsync(w->d_tab, w->h_tab, 4 * T * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream));
    CU_TRY(cudaEventRecord(w->tab_ready, w->stream));
    CU_TRY(cudaEventRecord(w->ev0, w->stream));
    // Asynchronous jobs reach the device in submission order, a whole call at a time: every launch of
    // this call is enqueued before a later job may launch (an encode launched in between would hold all
    // SMs while the remaining chunks of this decode wait behind it).
    gate_enter();
    struct GateGuard {
        ~GateGuard() { gate_leave(); }
    } gate_guard;
    for (int c = 0; c < nchunks; c++) {
        cudaStream_t cs = w->cs[c];
        const int f = first[c], l = first[c + 1], m = l - f;
        CU_TRY(cudaStreamWaitEvent(cs, w->tab_ready, 0));
        // compressed bytes of this chunk: [min begin, max end) over its non-empty ranges
        uint64_t clo = ~0ull, chi = 0;
        for (int i = f; i < l; i++)
            if (h[T + i] > h[i]) {
                if (h[i] < clo) clo = h[i];
                if (h[T + i] > chi) chi = h[T + i];
            }
        if (nchunks == 1) {
            clo = 0;
            chi = hi - lo;
        }
        if (chi > clo && clo != ~0ull) CU_TRY(cudaMemcpyAsync(w->d_src + clo, src + lo + clo, chi - clo, cudaMemcpyHostToDevice, cs));
        rc = launch_decode(device, m, w->d_src, w->d_tab + f, w->d_tab + T + f, w->d_dst, w->d_tab + 2 * T + f,
                           w->d_tab + 3 * T + f, d_status + f, cs, nblk, false);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(w->chunk_ev[c], cs));
        if (crc_out) {  // checksum of the decoded blocks while they are resident (reader.go:341-351)
            rc = launch_crc(device, m, w->d_dst, w->d_tab + 2 * T + f, w->d_tab + 3 * T + f, d_crc + f, cs);
            if (rc) return rc;
            CU_TRY(cudaMemcpyAsync(h_crc + f, d_crc + f, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
        }
        CU_TRY(cudaMemcpyAsync(h_status + f, d_status + f, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
        if (dense) {
            const size_t o0 = dbeg[f] - dbeg[0], o1 = dend[l - 1] - dbeg[0];
            if (o1 > o0) CU_TRY(cudaMemcpyAsync(dst + dbeg[f], w->d_dst + o0, o1 - o0, cudaMemcpyDeviceToHost, cs));
        } else {
            for (int i = f; i < l; i++) {
                size_t mm = dend[i] - dbeg[i];
                if (mm) CU_TRY(cudaMemcpyAsync(dst + dbeg[i], w->d_dst + h[2 * T + i], mm, cudaMemcpyDeviceToHost, cs));
            }
        }
    }
    // "every decode kernel of this call is done" for the encode launches that follow (DeviceState::dec_done)
    for (int c = 0; c < nchunks; c++) CU_TRY(cudaStreamWaitEvent(w->stream, w->chunk_ev[c], 0));
    {
        std::lock_guard<std::mutex> lk(g_dev[device].order_mu);
        CU_TRY(cudaEventRecord(g_dev[device].dec_done, w->stream));
    }
    gate_leave();
    for (int c = 0; c < nchunks; c++) CU_TRY(cudaStreamSynchronize(w->cs[c]));
    CU_TRY(cudaEventRecord(w->ev1, w->stream));
    CU_TRY(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < nblk; i++) status[i] = h_status[i];
    if (crc_out) memcpy(crc_out, h_crc, (size_t)nblk * sizeof(uint32_t));
    cudaEventElapsedTime(&g_last_kernel_ms, w->ev0, w->ev1);
    return MZCU_OK;
}

}  // namespace

// ---- helpers of the multi-device and asynchronous entry points ----
namespace {
struct MultiPart {
    int rc = MZCU_OK;
    char err[512] = "";
};
template <class F>
int run_on_devices(int ndev, const int *devices, int nblk, F &&call) {
    if (ndev <= 0 || !devices) return fail(MZCU_ERR_INVALID_ARG, "empty device list");
    if (ndev > kMaxDevices) return fail(MZCU_ERR_INVALID_ARG, "too many devices");
    std::vector<MultiPart> part(ndev);
    std::vector<std::thread> th;
    for (int k = 0; k < ndev; k++) {
        const int lo = (int)((int64_t)nblk * k / ndev), hi = (int)((int64_t)nblk * (k + 1) / ndev);
        if (hi == lo) continue;
        auto work = [&, k, lo, hi] {
            mzcu_bind_host_to_device(devices[k]);
            part[k].rc = call(devices[k], lo, hi);
            if (part[k].rc) snprintf(part[k].err, sizeof part[k].err, "device %d: %.480s", devices[k], g_err);
        };
        try {
            th.emplace_back(work);
        } catch (...) {  // no thread to be had: this range runs on the caller's thread
            work();
        }
    }
    for (auto &t : th) t.join();
    for (int k = 0; k < ndev; k++)
        if (part[k].rc) return fail(part[k].rc, "%s", part[k].err);
    return MZCU_OK;
}
}  // namespace

namespace {
struct Job {
    std::thread th;
    int rc = MZCU_OK;
    char err[512] = "";
};
std::mutex g_job_mu;
std::map<int64_t, Job *> g_jobs;
int64_t g_next_job = 1;

template <class F>
int64_t submit_job(F &&call) {
    Job *j = new Job();
    int bound = -1;
    if (cudaGetDevice(&bound) != cudaSuccess) {
        cudaGetLastError();
        bound = -1;
    }
    std::lock_guard<std::mutex> lk(g_job_mu);
    const int64_t id = g_next_job;
    try {
        j->th = std::thread([j, call, bound, id] {
            if (bound >= 0) cudaSetDevice(bound);
            t_ticket = id;
            j->rc = call();
            if (j->rc) snprintf(j->err, sizeof j->err, "%s", g_err);
            gate_enter();  // a job that never launched still takes (and passes on) its turn
            gate_leave();
        });
    } catch (...) {
        delete j;
        return fail(MZCU_ERR_INVALID_ARG, "cannot start a job thread");
    }
    g_next_job++;  // the ticket is taken only when its thread exists (the launch gate waits for every ticket)
    g_jobs[id] = j;
    return id;
}
}  // namespace

// ============================ exported C ABI ================================
extern "C" {

int mzcu_abi_version(void) { return MZCU_ABI_VERSION; }

int mzcu_set_encoder_flavor(int flavor) {
    if (flavor != MZCU_FLAVOR_GO && flavor != MZCU_FLAVOR_AMD64)
        return fail(MZCU_ERR_INVALID_ARG, "unknown encoder flavour %d", flavor);
    g_flavor.store(flavor, std::memory_order_relaxed);
    return MZCU_OK;
}

int mzcu_get_encoder_flavor(void) { return g_flavor.load(std::memory_order_relaxed); }

#ifdef MZ_ENC_STATS
// profiling builds only (profiles/enc_stats.py): read and clear the walk counters
extern "C" int mzcu_debug_enc_stats(unsigned long long *out) {
    if (cudaMemcpyFromSymbol(out, mz::g_enc_stats, sizeof(mz::g_enc_stats)) != cudaSuccess) return MZCU_ERR_CUDA;
    unsigned long long zero[16] = {0};
    if (cudaMemcpyToSymbol(mz::g_enc_stats, zero, sizeof zero) != cudaSuccess) return MZCU_ERR_CUDA;
    return MZCU_OK;
}
#endif
const char *mzcu_last_error(void) { return g_err; }
float mzcu_last_kernel_ms(void) { return g_last_kernel_ms; }

int mzcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t mzcu_max_encoded_len(int64_t src_len) {
    if (src_len < 0 || src_len > MZCU_MAX_BLOCK_SIZE) return -1;
    if (src_len == 0) return 1;
    return src_len + 2;
}

int mzcu_is_minlz(const uint8_t *block, size_t n, int *is_minlz, int64_t *size) {
    int mlz, lits;
    size_t hdr;
    int64_t sz;
    int rc = parse_block_header(block, n, &mlz, &lits, &hdr, &sz);
    if (is_minlz) *is_minlz = mlz;
    if (size) *size = sz;
    if (rc) return fail(rc, "invalid block header");
    return MZCU_OK;
}

int64_t mzcu_decoded_len(const uint8_t *block, size_t n) {
    int mlz;
    int64_t sz;
    int rc = mzcu_is_minlz(block, n, &mlz, &sz);
    return rc ? rc : sz;
}

int mzcu_encode_blocks_dev(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                           const uint64_t *dst_off, uint32_t *out_len, void *stream) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_off || !dst || !dst_off || !out_len)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (level != MZCU_LEVEL_SUPERFAST && level != MZCU_LEVEL_FASTEST && level != MZCU_LEVEL_BALANCED)
        return fail(MZCU_ERR_INVALID_LEVEL, "invalid level %d", level);
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    rc = launch_encode(device, level, nblk, src, src_off, src_off + 1, dst, dst_off, out_len, (cudaStream_t)stream);
    if (rc == MZCU_OK && validate_on())
        rc = validate_encoded(device, nblk, src, src_off, src_off + 1, dst, dst_off, out_len, (cudaStream_t)stream);
    return rc;
}

int mzcu_decode_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                           const uint64_t *dst_off, int32_t *status, void *stream) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_off || !dst || !dst_off || !status)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    return launch_decode(device, nblk, src, src_off, src_off + 1, dst, dst_off, dst_off + 1, status, (cudaStream_t)stream);
}

int mzcu_pack_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, const uint32_t *len,
                         uint8_t *dst, uint64_t *dst_off, void *stream) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_off || !len || !dst || !dst_off)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    return launch_pack(device, nblk, src, src_off, len, dst, dst_off, (cudaStream_t)stream);
}

int mzcu_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                       const uint64_t *dst_off, uint32_t *out_len) {
    return host_encode_blocks(device, level, nblk, src, src_off, dst, dst_off, out_len);
}

int mzcu_encode_blocks_packed(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                              uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out) {
    return host_encode_blocks_packed(device, level, nblk, src, src_off, dst, dst_cap, dst_off_out, nullptr);
}

int mzcu_stream_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                              size_t dst_cap, uint64_t *dst_off_out, uint32_t *crc_out) {
    return host_encode_blocks_packed(device, level, nblk, src, src_off, dst, dst_cap, dst_off_out, crc_out);
}

int mzcu_stream_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                              const uint64_t *dst_off, int32_t *status, uint32_t *crc_out) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !status))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    return host_decode_ranges(device, nblk, src, src_off, src_off + 1, dst, dst_off, dst_off + 1, status, crc_out);
}

int mzcu_crc32c_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint32_t *crc, void *stream) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_off || !crc))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    return launch_crc(device, nblk, src, src_off, src_off + 1, crc, (cudaStream_t)stream);
}

int mzcu_crc32c_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint32_t *crc) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !crc))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (nblk == 0) return MZCU_OK;
    for (int i = 0; i < nblk; i++)
        if (src_off[i + 1] < src_off[i] || src_off[i + 1] - src_off[i] > 0xffffffffull)
            return fail(MZCU_ERR_TOO_LARGE, "block %d: the checksum kernel takes blocks below 4 GiB", i);
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    WsGuard g;
    rc = ws_acquire(device, &g.w);
    if (rc) return rc;
    Workspace *w = g.w;
    const size_t base = src_off[0], total = src_off[nblk] - base;
    rc = ws_reserve(w, total, 16, (size_t)nblk);
    if (rc) return rc;
    const size_t T = w->cap_tab;
    for (int i = 0; i <= nblk; i++) w->h_tab[i] = src_off[i] - base;
    uint32_t *d_crc = reinterpret_cast<uint32_t *>(w->d_tab + 4 * T);
    if (total) CU_TRY(cudaMemcpyAsync(w->d_src, src + base, total, cudaMemcpyHostToDevice, w->stream));
    CU_TRY(cudaMemcpyAsync(w->d_tab, w->h_tab, (size_t)(nblk + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream));
    rc = launch_crc(device, nblk, w->d_src, w->d_tab, w->d_tab + 1, d_crc, w->stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(crc, d_crc, (size_t)nblk * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    CU_TRY(cudaStreamSynchronize(w->stream));
    return MZCU_OK;
}

int mzcu_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                       const uint64_t *dst_off, int32_t *status) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !status))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    return host_decode_ranges(device, nblk, src, src_off, src_off + 1, dst, dst_off, dst_off + 1, status);
}

int mzcu_encode_batch(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                      const uint64_t *dst_off, uint64_t *enc_len) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !enc_len))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (level != MZCU_LEVEL_SUPERFAST && level != MZCU_LEVEL_UNCOMPRESSED && level != MZCU_LEVEL_FASTEST && level != MZCU_LEVEL_BALANCED)
        return fail(MZCU_ERR_INVALID_LEVEL, "invalid level %d", level);
    // Per block: dst = 0x00 uvarint(n) tokens  (encode.go:87-90), with the
    // token stream produced right behind the header.
    std::vector<uint64_t> tsrc(nblk + 1), tdst((size_t)nblk + 1);
    std::vector<uint32_t> out(nblk ? nblk : 1);
    std::vector<int> hdr(nblk ? nblk : 1);
    for (int i = 0; i < nblk; i++) {
        size_t n = src_off[i + 1] - src_off[i];
        size_t cap = dst_off[i + 1] - dst_off[i];
        int64_t need = mzcu_max_encoded_len((int64_t)n);
        if (need < 0) return fail(MZCU_ERR_TOO_LARGE, "block %d larger than 8 MiB", i);
        if ((int64_t)cap < need) return fail(MZCU_ERR_DST_TOO_SMALL, "block %d: dst capacity < MaxEncodedLen", i);
    }
    bool any = false;
    if (level != MZCU_LEVEL_UNCOMPRESSED)
        for (int i = 0; i < nblk; i++) any |= (src_off[i + 1] - src_off[i]) >= 16;
    std::vector<uint8_t> tmp;
    std::vector<uint64_t> toff((size_t)nblk + 1, 0);
    if (any) {
        // token streams go to a scratch area with full MaxEncodedLen capacity
        // (the header eats up to 5 bytes of the caller's n+2).
        for (int i = 0; i < nblk; i++) toff[i + 1] = toff[i] + (src_off[i + 1] - src_off[i]) + 2;
        tmp.resize(toff[nblk] + 16);
        int rc = host_encode_blocks(device, level, nblk, src, src_off, tmp.data(), toff.data(), out.data());
        if (rc) return rc;
    }
    for (int i = 0; i < nblk; i++) {
        const uint8_t *s = src + src_off[i];
        size_t n = src_off[i + 1] - src_off[i];
        uint8_t *d = dst + dst_off[i];
        size_t cap = dst_off[i + 1] - dst_off[i];
        if (any && n >= 16 && out[i] > 0) {
            d[0] = 0;
            int h = 1 + put_uvarint(d + 1, n);
            memcpy(d + h, tmp.data() + toff[i], out[i]);
            enc_len[i] = (uint64_t)h + out[i];
        } else {
            int64_t m = store_uncompressed(d, cap, s, n);
            if (m < 0) return (int)m;
            enc_len[i] = (uint64_t)m;
        }
    }
    return MZCU_OK;
}

int mzcu_decode_batch(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                      const uint64_t *dst_off, int64_t *dec_len) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !dec_len))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
    std::vector<uint64_t> sbeg(nblk ? nblk : 1), send(nblk ? nblk : 1), dbeg(nblk ? nblk : 1), dend(nblk ? nblk : 1);
    std::vector<int32_t> status(nblk ? nblk : 1, 0);
    bool any = false;
    for (int i = 0; i < nblk; i++) {
        const uint8_t *b = src + src_off[i];
        size_t n = src_off[i + 1] - src_off[i];
        size_t cap = dst_off[i + 1] - dst_off[i];
        int mlz, lits;
        size_t hdr;
        int64_t sz;
        int rc = parse_block_header(b, n, &mlz, &lits, &hdr, &sz);
        sbeg[i] = send[i] = src_off[i];
        dbeg[i] = dend[i] = dst_off[i];
        if (rc) {
            dec_len[i] = rc;
        } else if (!mlz) {
            dec_len[i] = MZCU_ERR_UNSUPPORTED;
        } else if ((size_t)sz > cap) {
            dec_len[i] = MZCU_ERR_DST_TOO_SMALL;
        } else if (lits) {
            memcpy(dst + dst_off[i], b + hdr, (size_t)sz);  // decode.go:55-57
            dec_len[i] = sz;
        } else {
            sbeg[i] = src_off[i] + hdr;
            send[i] = src_off[i + 1];
            dend[i] = dst_off[i] + (uint64_t)sz;
            dec_len[i] = sz;
            any = true;
        }
    }
    if (any) {
        int rc = host_decode_ranges(device, nblk, src, sbeg.data(), send.data(), dst, dbeg.data(), dend.data(), status.data());
        if (rc) return rc;
        for (int i = 0; i < nblk; i++)
            if (status[i] != 0) dec_len[i] = MZCU_ERR_CORRUPT;
    }
    return MZCU_OK;
}

int64_t mzcu_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level) {
    int64_t need = mzcu_max_encoded_len((int64_t)n);
    if (need < 0) return fail(MZCU_ERR_TOO_LARGE, "block larger than 8 MiB");
    if (n < 16) return store_uncompressed(dst, dst_cap, src, n);  // encode.go:83-85
    if ((int64_t)dst_cap < need) return fail(MZCU_ERR_DST_TOO_SMALL, "dst capacity < MaxEncodedLen");
    uint64_t so[2] = {0, n}, dof[2] = {0, dst_cap}, el = 0;
    int rc = mzcu_encode_batch(-1, level, 1, src, so, dst, dof, &el);
    return rc ? rc : (int64_t)el;
}

int64_t mzcu_try_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level) {
    // encode.go:168-207: nil unless compressible and d+n < len(src)
    int64_t need = mzcu_max_encoded_len((int64_t)n);
    if (need < 0 || (int64_t)dst_cap < need || n < 16) return 0;
    if (level != MZCU_LEVEL_SUPERFAST && level != MZCU_LEVEL_FASTEST && level != MZCU_LEVEL_BALANCED) return 0;
    std::vector<uint8_t> tmp(n + 18);
    uint64_t so[2] = {0, n}, dof[2] = {0, n + 2};
    uint32_t out = 0;
    int rc = host_encode_blocks(-1, level, 1, src, so, tmp.data(), dof, &out);
    if (rc) return rc;
    dst[0] = 0;
    int h = 1 + put_uvarint(dst + 1, n);
    if (out > 0 && (size_t)h + out < n) {
        memcpy(dst + h, tmp.data(), out);
        return (int64_t)h + out;
    }
    return 0;
}

int64_t mzcu_decode(uint8_t *dst, size_t dst_cap, const uint8_t *block, size_t n) {
    uint64_t so[2] = {0, n}, dof[2] = {0, dst_cap};
    int64_t dl = 0;
    int rc = mzcu_decode_batch(-1, 1, block, so, dst, dof, &dl);
    if (rc) return rc;
    if (dl < 0) return fail((int)dl, "decode failed (%lld)", (long long)dl);
    return dl;
}

int mzcu_set_validate(int on) {
    g_validate.store(on ? 1 : 0, std::memory_order_relaxed);
    return MZCU_OK;
}

int mzcu_get_validate(void) { return validate_on() ? 1 : 0; }

int mzcu_validate_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, const uint8_t *enc,
                             const uint64_t *enc_off, const uint32_t *out_len, void *stream) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_off || !enc || !enc_off || !out_len)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    device = resolve_device(device);
    if (device < 0) return fail(MZCU_ERR_CUDA, "no CUDA device");
    int rc = init_device(device);
    if (rc) return rc;
    return validate_encoded(device, nblk, src, src_off, src_off + 1, enc, enc_off, out_len, (cudaStream_t)stream);
}

// ---- host placement -----------------------------------------------------------
// Pins the calling thread (and the threads it creates later) to the CPUs of the NUMA node the
// device hangs off, and prefers that node for its future allocations (pinned staging buffers are
// first-touched by the thread that allocates them).  Returns the node, or -1 when the platform
// exposes none / the device has no affinity -- in which case nothing is changed.
int mzcu_bind_host_to_device(int device) {
    device = resolve_device(device);
    if (device < 0) return -1;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; c++)
        if (*c >= 'A' && *c <= 'Z') *c += 'a' - 'A';
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0 || node >= 1024) return -1;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return -1;
    cpu_set_t set;
    CPU_ZERO(&set);
    int a, b, any = 0;
    for (;;) {  // "0-15,32-47"
        if (fscanf(f, "%d", &a) != 1) break;
        b = a;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &b) != 1) break;
            c = fgetc(f);
        }
        for (int k = a; k <= b && k < CPU_SETSIZE; k++) {
            CPU_SET(k, &set);
            any = 1;
        }
        if (c != ',') break;
    }
    fclose(f);
    if (!any) return -1;
    sched_setaffinity(0, sizeof set, &set);  // best effort (a container may forbid it)
#ifdef SYS_set_mempolicy
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof(mask) * 8);
#endif
    return node;
}

// ---- several devices behind one call (mzcu_init(ndev, devs)) ----------
// A Go host cannot use torch.distributed: these entry points shard one batch over a device list
// inside the library -- contiguous block ranges in stream order (device k gets blocks
// [nblk*k/ndev, nblk*(k+1)/ndev)), one host thread per device, each running the single-device
// host call on its range.  Blocks are independent, so there is no exchange between devices.

int mzcu_stream_encode_blocks_multi(int ndev, const int *devices, int level, int nblk, const uint8_t *src,
                                    const uint64_t *src_off, uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out,
                                    uint32_t *out_len, uint32_t *crc_out) {
    if (nblk < 0 || (nblk > 0 && (!src_off || !dst || !dst_off_out || !out_len)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (nblk == 0) return MZCU_OK;
    // device k packs its range into the part of dst that mirrors its source range (a packed block
    // is never longer than its source), so the ranges of different devices cannot collide
    if (dst_cap < src_off[nblk] - src_off[0]) return fail(MZCU_ERR_DST_TOO_SMALL, "dst_cap < total source bytes");
    return run_on_devices(ndev, devices, nblk, [&](int dev, int lo, int hi) -> int {
        const int m = hi - lo;
        std::vector<uint64_t> off((size_t)m + 1);
        const size_t base = src_off[lo] - src_off[0];
        int rc = host_encode_blocks_packed(dev, level, m, src, src_off + lo, dst + base, src_off[hi] - src_off[lo], off.data(),
                                           crc_out ? crc_out + lo : nullptr);
        if (rc) return rc;
        for (int i = 0; i < m; i++) {
            dst_off_out[lo + i] = base + off[i];
            out_len[lo + i] = (uint32_t)(off[i + 1] - off[i]);
        }
        return MZCU_OK;
    });
}

int mzcu_stream_decode_blocks_multi(int ndev, const int *devices, int nblk, const uint8_t *src, const uint64_t *src_beg,
                                    const uint32_t *src_len, uint8_t *dst, const uint64_t *dst_off, int32_t *status,
                                    uint32_t *crc_out) {
    if (nblk < 0 || (nblk > 0 && (!src || !src_beg || !src_len || !dst || !dst_off || !status)))
        return fail(MZCU_ERR_INVALID_ARG, "null argument");
    if (nblk == 0) return MZCU_OK;
    return run_on_devices(ndev, devices, nblk, [&](int dev, int lo, int hi) -> int {
        const int m = hi - lo;
        std::vector<uint64_t> sbeg((size_t)m), send((size_t)m);
        for (int i = 0; i < m; i++) {
            sbeg[i] = src_beg[lo + i];
            send[i] = src_beg[lo + i] + src_len[lo + i];
        }
        return host_decode_ranges(dev, m, src, sbeg.data(), send.data(), dst, dst_off + lo, dst_off + lo + 1, status + lo,
                                  crc_out ? crc_out + lo : nullptr);
    });
}

// ---- asynchronous host calls ---------------------------------------------------
// submit returns at once with a job id (> 0); the call runs on a library thread with its own
// pooled workspace and streams, so the download of call k overlaps the upload and the kernels of
// call k+1 (PCIe is full duplex).  mzcu_wait blocks until the job is done, returns the call's
// result (its message is copied to this thread's mzcu_last_error) and retires the job.

int64_t mzcu_submit_stream_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                                         uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out, uint32_t *crc_out) {
    return submit_job([=]() -> int {
        return host_encode_blocks_packed(device, level, nblk, src, src_off, dst, dst_cap, dst_off_out, crc_out);
    });
}

int64_t mzcu_submit_stream_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                                         const uint64_t *dst_off, int32_t *status, uint32_t *crc_out) {
    return submit_job([=]() -> int {
        if (nblk < 0 || (nblk > 0 && (!src_off || !dst_off || !status))) return fail(MZCU_ERR_INVALID_ARG, "null argument");
        return host_decode_ranges(device, nblk, src, src_off, src_off + 1, dst, dst_off, dst_off + 1, status, crc_out);
    });
}

int mzcu_wait(int64_t job) {
    Job *j = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_job_mu);
        auto it = g_jobs.find(job);
        if (it == g_jobs.end()) return fail(MZCU_ERR_INVALID_ARG, "unknown job %lld", (long long)job);
        j = it->second;
        g_jobs.erase(it);
    }
    j->th.join();
    const int rc = j->rc;
    if (rc) snprintf(g_err, sizeof g_err, "%s", j->err);
    delete j;
    return rc;
}

void *mzcu_host_alloc(size_t n) {
    void *p = nullptr;
    if (cudaMallocHost(&p, n ? n : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void mzcu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
