// abi.cu -- the extern "C" surface of include/bbduk_b200.h: handle management, table build/replication,
// host-buffer batching (pinned staging, two streams so copies overlap kernels) and kernel dispatch.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bbduk_b200.h"
#include "hostpack.h"
#include "params.h"
#include "probe.h"
#include "table.h"

namespace {

constexpr int N_SLOTS = 4;                    // staging slots (streams) per handle
constexpr int64_t CHUNK_READS = 1 << 19;      // reads per device batch on the host path: small enough that H2D of
                                              // chunk i+1, the kernels of chunk i and D2H of chunk i-1 overlap
constexpr int64_t CHUNK_BYTES = 256ll << 20;  // bases per device batch (offsets stay 32-bit)

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *d_bases = nullptr;
    int64_t *d_off64 = nullptr;
    uint32_t *d_off32 = nullptr;
    int32_t *d_id0 = nullptr, *d_id0b = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_count = nullptr;
    uint8_t *d_flags = nullptr;
    uint32_t *d_maskbits = nullptr;
    int64_t *d_maskoff = nullptr;
    int32_t *d_handoff = nullptr;
    unsigned int *d_handoff_n = nullptr;
    // scratch of the direct (HBM-table) path: per-read first/last hit, read-start bit stream
    unsigned long long *d_first64 = nullptr;
    int *d_lastpos = nullptr;
    uint16_t *d_sbits = nullptr;
    // host-packed transfer (hostpack.h): pinned staging + device copies of the 2-bit stream F and defined bits D
    uint32_t *h_F = nullptr, *d_F = nullptr;
    uint16_t *h_D = nullptr, *d_D = nullptr;
    uint32_t *h_off32 = nullptr;
    uint64_t *h_exc = nullptr, *d_exc = nullptr;  // groups with an undefined base, (group << 16 | D): what travels instead of D
    int64_t cap_groups = 0, cap_hoff = 0, cap_exc = 0;
    int64_t cap_bases = 0, cap_reads = 0, cap_maskwords = 0, cap_sbits = 0;
    std::mutex mu;
};

}  // namespace

struct bbduk_handle {
    bbduk_cfg cfg;
    BBParams p;
    int device = 0;
    int sm_count = 148;
    bool finalized = false;
    std::vector<uint8_t> ref;
    std::vector<int64_t> ref_off{0};
    DeviceTable table;
    unsigned long long *d_scaf_reads = nullptr, *d_scaf_bases = nullptr;
    bbduk_stats *d_stats = nullptr;
    Slot slots[N_SLOTS];
    std::atomic<int> next_slot{0};
    std::atomic<int64_t> launches{0};
    std::atomic<int64_t> h2d_bytes{0}, d2h_bytes{0};  // bytes bbduk_b200_process / _process_packed moved across PCIe
    std::atomic<int> max_read_len_hint{0};
    bool trace = false;     // BBDUK_B200_TRACE=1: per-chunk host timings on stderr
    bool ascii_every_set = false;
    int ascii_every = 3;    // BBDUK_B200_ASCII_EVERY=n: every n-th chunk crosses PCIe as ASCII (0 = never); unset = adaptive
    // adaptive share of ASCII chunks: the workers' packing rate is measured (seconds per base, moving average) and set against
    // the link (BBDUK_B200_PCIE_GBS, default 52): packing a chunk costs tp of CPU and 0.375 tb of link, shipping it as ASCII
    // costs tb of link and no CPU; both resources are busy the same time when x = (tp - 0.375 tb) / (tp + 0.625 tb) of the
    // chunks go ASCII
    double pack_s_per_base = 0.0;
    double ascii_acc = 0.0;
    double pcie_gbs = 52.0;
    bool pack_host = true;  // BBDUK_B200_PACK_HOST=0 keeps the bases ASCII across PCIe
    bool sparse_d = true;   // BBDUK_B200_SPARSE_D=0 uploads the defined bits as an array instead of all-ones + exceptions
    std::mutex err_mu;
    std::string err;
    // per-thread-stream scratch for process_device
    int32_t *dev_handoff = nullptr;
    unsigned int *dev_handoff_n = nullptr;
    int64_t dev_handoff_cap = 0;
    unsigned long long *dev_first64 = nullptr;
    int *dev_lastpos = nullptr;
    uint16_t *dev_sbits = nullptr;
    int64_t dev_sbits_cap = 0, dev_first_cap = 0;
    struct TboBuf {  // staging of the synchronous bbduk_b200_tbo entry point
        uint8_t *d_bases = nullptr, *d_quals = nullptr, *d_flags = nullptr;
        uint32_t *d_off = nullptr;
        int32_t *d_lo = nullptr, *d_hi = nullptr, *d_insert = nullptr;
        int32_t *d_id0 = nullptr, *d_count = nullptr;
        uint32_t *d_mask = nullptr;   // entropy masking: mask words of the chunk
        int64_t *d_maskoff = nullptr; // and their per-read word offsets
        int64_t cap_bases = 0, cap_quals = 0, cap_flags = 0, cap_off = 0, cap_lo = 0, cap_hi = 0, cap_insert = 0, cap_id0 = 0,
                cap_count = 0, cap_mask = 0, cap_maskoff = 0;
    } tbo;
    // bbduk_b200_process_chain: second input slot + copy stream, so that the upload of chunk i+1 overlaps the kernels of chunk i
    uint8_t *chain_bases2 = nullptr, *chain_quals2 = nullptr;
    uint32_t *chain_off2 = nullptr;
    int64_t chain_cap_bases2 = 0, chain_cap_quals2 = 0, chain_cap_off2 = 0;
    // packed upload of the chain (chunks made of A C G T N only): pinned host staging + device streams per input slot
    uint32_t *chain_hF[2] = {nullptr, nullptr}, *chain_dF[2] = {nullptr, nullptr};
    uint16_t *chain_hD[2] = {nullptr, nullptr}, *chain_dD[2] = {nullptr, nullptr};
    uint32_t *chain_hoff[2] = {nullptr, nullptr};
    int64_t chain_cap_hg[2] = {0, 0}, chain_cap_dg[2] = {0, 0}, chain_cap_hoff[2] = {0, 0};
    int64_t *chain_dst = nullptr;
    cudaStream_t chain_copy = nullptr;
    cudaEvent_t chain_in[2] = {nullptr, nullptr}, chain_free[2] = {nullptr, nullptr};
    std::mutex tbo_mu;
    int replicated_via = 0;    // 0 = built here, 1 = NCCL broadcast, 2 = peer copy (bbduk_b200_replicate)
    HostPool *pool = nullptr;  // host packing workers, created on first use
    std::mutex pool_mu;
    std::mutex dev_mu;
};

namespace {

thread_local std::string g_err;  // errors raised before a handle exists

int set_err(bbduk_handle *h, const std::string &m) {
    if (h) {
        std::lock_guard<std::mutex> g(h->err_mu);
        h->err = m;
    }
    g_err = m;
    return 1;
}

#define CKH(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return set_err(h, b_);                                                                         \
        }                                                                                                  \
    } while (0)

__global__ void off64_to_32_kernel(const int64_t *__restrict__ off64, int64_t base, uint32_t *__restrict__ off32, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off32[i] = (uint32_t)(off64[i] - base);
}
// defined bits on the device = all ones (a memset) + the listed exceptions
__global__ void defined_exceptions_kernel(const uint64_t *__restrict__ exc, int64_t n, uint16_t *__restrict__ D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) D[exc[i] >> 16] = (uint16_t)(exc[i] & 0xFFFFu);
}
// host-packed input for a mode the tuned kernels do not serve: spell the 2-bit stream out again (undefined -> 'N')
__global__ void unpack_kernel(const uint32_t *__restrict__ F, const uint16_t *__restrict__ D, uint8_t *__restrict__ bases, int64_t groups) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const uint32_t f = F[g], d = D[g];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = 4 * q + j;
            const uint32_t ch = ((d >> (15 - b)) & 1u) ? ((0x54474341u >> (8 * ((f >> (30 - 2 * b)) & 3u))) & 0xFFu) : (uint32_t)'N';
            x |= ch << (8 * j);
        }
        w[q] = x;
    }
    reinterpret_cast<uint4 *>(bases)[g] = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void maskoff_rebase_kernel(int64_t *off, int64_t base, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] -= base;
}
__global__ void max_len_kernel(const uint32_t *__restrict__ off, int64_t n_reads, unsigned int *out) {
    unsigned int m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, off[i + 1] - off[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

void free_slot(Slot &s) {
    cudaFree(s.d_bases);
    cudaFree(s.d_off64);
    cudaFree(s.d_off32);
    cudaFree(s.d_id0);
    cudaFree(s.d_id0b);
    cudaFree(s.d_lo);
    cudaFree(s.d_hi);
    cudaFree(s.d_count);
    cudaFree(s.d_flags);
    cudaFree(s.d_maskbits);
    cudaFree(s.d_maskoff);
    cudaFree(s.d_handoff);
    cudaFree(s.d_handoff_n);
    cudaFree(s.d_first64);
    cudaFree(s.d_lastpos);
    cudaFree(s.d_sbits);
    cudaFree(s.d_F);
    cudaFree(s.d_D);
    cudaFreeHost(s.h_F);
    cudaFreeHost(s.h_D);
    cudaFreeHost(s.h_off32);
    cudaFree(s.d_exc);
    cudaFreeHost(s.h_exc);
    s.d_exc = s.h_exc = nullptr;
    s.cap_exc = 0;
    s.d_F = s.h_F = s.h_off32 = nullptr;
    s.d_D = s.h_D = nullptr;
    s.cap_groups = s.cap_hoff = 0;
    s.d_first64 = nullptr;
    s.d_lastpos = nullptr;
    s.d_sbits = nullptr;
    s.cap_sbits = 0;
    s.d_bases = nullptr;
    s.d_off64 = nullptr;
    s.d_off32 = nullptr;
    s.d_id0 = s.d_id0b = s.d_lo = s.d_hi = s.d_count = nullptr;
    s.d_flags = nullptr;
    s.d_maskbits = nullptr;
    s.d_maskoff = nullptr;
    s.d_handoff = nullptr;
    s.d_handoff_n = nullptr;
    s.cap_bases = s.cap_reads = s.cap_maskwords = 0;
}

int ensure_slot(bbduk_handle *h, Slot &s, int64_t n_reads, int64_t n_bases, int64_t n_maskwords, bool direct) {
    if (direct && (int64_t)direct_sbits_bytes(n_bases) > s.cap_sbits) {
        cudaFree(s.d_sbits);
        s.d_sbits = nullptr;
        s.cap_sbits = (int64_t)direct_sbits_bytes(n_bases + n_bases / 8 + 4096);
        CKH(cudaMalloc(&s.d_sbits, (size_t)s.cap_sbits));
    }
    if (!s.st) {
        CKH(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        CKH(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    if (n_bases + 64 > s.cap_bases) {
        cudaFree(s.d_bases);
        s.cap_bases = n_bases + n_bases / 8 + 4096;
        CKH(cudaMalloc(&s.d_bases, (size_t)s.cap_bases));
    }
    if (n_reads + 1 > s.cap_reads) {
        const int64_t c = n_reads + n_reads / 8 + 1024;
        cudaFree(s.d_off64);
        cudaFree(s.d_off32);
        cudaFree(s.d_id0);
        cudaFree(s.d_id0b);
        cudaFree(s.d_lo);
        cudaFree(s.d_hi);
        cudaFree(s.d_count);
        cudaFree(s.d_flags);
        cudaFree(s.d_maskoff);
        cudaFree(s.d_handoff);
        cudaFree(s.d_handoff_n);
        cudaFree(s.d_first64);
        cudaFree(s.d_lastpos);
        CKH(cudaMalloc(&s.d_first64, sizeof(unsigned long long) * c));
        CKH(cudaMalloc(&s.d_lastpos, sizeof(int) * c));
        CKH(cudaMalloc(&s.d_off64, sizeof(int64_t) * c));
        CKH(cudaMalloc(&s.d_off32, sizeof(uint32_t) * c));
        CKH(cudaMalloc(&s.d_id0, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_id0b, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_lo, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_hi, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_count, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_flags, (size_t)c));
        CKH(cudaMalloc(&s.d_maskoff, sizeof(int64_t) * c));
        CKH(cudaMalloc(&s.d_handoff, sizeof(int32_t) * c));
        CKH(cudaMalloc(&s.d_handoff_n, sizeof(unsigned int) * 4));
        s.cap_reads = c;
    }
    if (n_maskwords > s.cap_maskwords) {
        cudaFree(s.d_maskbits);
        s.cap_maskwords = n_maskwords + n_maskwords / 8 + 1024;
        CKH(cudaMalloc(&s.d_maskbits, sizeof(uint32_t) * (size_t)s.cap_maskwords));
    }
    return 0;
}

int ensure_packed(bbduk_handle *h, Slot &s, int64_t n_reads, int64_t n_bases) {
    const int64_t groups = (n_bases + 15) / 16 + 8;
    if (groups > s.cap_groups) {
        cudaFree(s.d_F);
        cudaFree(s.d_D);
        cudaFreeHost(s.h_F);
        cudaFreeHost(s.h_D);
        s.d_F = s.h_F = nullptr;
        s.d_D = s.h_D = nullptr;
        s.cap_groups = groups + groups / 8 + 1024;
        CKH(cudaMalloc(&s.d_F, sizeof(uint32_t) * s.cap_groups));
        CKH(cudaMalloc(&s.d_D, sizeof(uint16_t) * s.cap_groups));
        CKH(cudaHostAlloc(&s.h_F, sizeof(uint32_t) * s.cap_groups, cudaHostAllocDefault));
        CKH(cudaHostAlloc(&s.h_D, sizeof(uint16_t) * s.cap_groups, cudaHostAllocDefault));
        cudaFree(s.d_exc);
        cudaFreeHost(s.h_exc);
        s.d_exc = s.h_exc = nullptr;
        s.cap_exc = s.cap_groups / 16 + 64 * 64;  // more undefined groups than that: the array travels instead
        CKH(cudaMalloc(&s.d_exc, sizeof(uint64_t) * s.cap_exc));
        CKH(cudaHostAlloc(&s.h_exc, sizeof(uint64_t) * s.cap_exc, cudaHostAllocDefault));
    }
    if (n_reads + 1 > s.cap_hoff) {
        cudaFreeHost(s.h_off32);
        s.h_off32 = nullptr;
        s.cap_hoff = n_reads + n_reads / 8 + 1024;
        CKH(cudaHostAlloc(&s.h_off32, sizeof(uint32_t) * s.cap_hoff, cudaHostAllocDefault));
    }
    return 0;
}

// dispatch one device-resident batch: fast kernel where it applies, generic kernel for the rest
struct DirectScratch {
    unsigned long long *first64;
    int *lastpos;
    uint16_t *sbits;
};

// which kernel family a batch of this handle goes to
bool uses_direct(const bbduk_handle *h, int max_read_len) {
    const BBTable t = h->table.view();
    return !plan_fast2(h->p, t, max_read_len).usable && !plan_fast(h->p, t, max_read_len).usable && plan_direct(h->p, t);
}

int run_batch(bbduk_handle *h, const uint8_t *d_bases, const uint32_t *d_off, int64_t n_reads, int64_t n_bases, int paired,
              const bbduk_out &dout, bbduk_stats *d_stats, int max_read_len, int32_t *d_handoff,
              unsigned int *d_handoff_n, const DirectScratch &ds, cudaStream_t st, const uint32_t *pk_F = nullptr,
              const uint16_t *pk_D = nullptr) {
    if (n_reads <= 0) return 0;
    const BBTable t = h->table.view();
    const int64_t n_units = paired ? n_reads / 2 : n_reads;
    const FastPlan plan2 = plan_fast2(h->p, t, max_read_len);
    const FastPlan plan = plan2.usable ? plan2 : plan_fast(h->p, t, max_read_len);
    if (plan.usable && d_handoff) {
        CKH(cudaMemsetAsync(d_handoff_n, 0, sizeof(unsigned int), st));
        const int nl = plan2.usable ? launch_fast2(plan, d_bases, d_off, n_reads, paired, h->p, t, dout, d_stats, h->d_scaf_reads,
                                                   h->d_scaf_bases, d_handoff, d_handoff_n, h->sm_count, st, pk_F, pk_D)
                                    : launch_fast(plan, d_bases, d_off, n_reads, paired, h->p, t, dout, d_stats, h->d_scaf_reads,
                                                  h->d_scaf_bases, d_handoff, d_handoff_n, h->sm_count, st, pk_F, pk_D);
        if (nl < 0) return set_err(h, std::string("fast kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        h->launches += nl;
        if (pk_F) return 0;  // packed batches are sized so that no tile is handed off
        // hand-offs (tiles that do not fit the staging): count stays on the device, no host sync
        if (launch_generic(d_bases, d_off, n_units, paired, d_handoff, d_handoff_n, h->p, t, dout, d_stats, h->d_scaf_reads,
                           h->d_scaf_bases, h->sm_count, st))
            return set_err(h, std::string("generic kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        h->launches += 1;
        return 0;
    }
    if (!plan.usable && plan_direct(h->p, t) && ds.first64 && ds.sbits && n_bases >= 0) {
        // HBM-resident table: flat position-parallel probe, then the per-read epilogue
        const int n1 = launch_direct(d_bases, d_off, n_reads, n_bases, paired, h->p, t, ds.first64, ds.lastpos, ds.sbits,
                                     h->sm_count, st);
        if (n1 < 0) return set_err(h, std::string("direct kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        const int n2 = launch_epilogue(d_bases, d_off, n_reads, paired, h->p, t, dout, d_stats, h->d_scaf_reads,
                                       h->d_scaf_bases, ds.first64, ds.lastpos, h->sm_count, st);
        if (n2 < 0) return set_err(h, std::string("epilogue kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        h->launches += n1 + n2;
        return 0;
    }
    if (launch_generic(d_bases, d_off, n_units, paired, nullptr, nullptr, h->p, t, dout, d_stats, h->d_scaf_reads,
                       h->d_scaf_bases, h->sm_count, st))
        return set_err(h, std::string("generic kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->launches += 1;
    return 0;
}

int check_mode_inputs(bbduk_handle *h, int paired, const bbduk_out *out) {
    if (!out) return set_err(h, "out is NULL");
    if (h->p.mode == MODE_KSPLIT && paired) return set_err(h, "Kmer splitting should only be performed on unpaired reads.");
    if (h->p.mode == MODE_KMASK && h->table.stored > 0 && (!out->maskbits || !out->mask_off))
        return set_err(h, "kmask mode needs out->maskbits and out->mask_off");
    return 0;
}

}  // namespace

extern "C" {

int bbduk_b200_version(void) { return BBDUK_B200_ABI_VERSION; }

void bbduk_b200_cfg_default(bbduk_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->generation = BBDUK_GEN_JGI;
    c->mink = -1;
    c->hdist2 = c->edist2 = c->qhdist2 = -1;
    c->rcomp = 1;
    c->mask_middle = 1;
    c->qskip = 1;
    c->min_skip = c->max_skip = 1;
    c->trim_symbol = 'N';
    c->min_read_length = 10;
    c->device = -1;
}

int bbduk_b200_describe_cfg(const bbduk_cfg *cfg, int64_t *v) {
    BBParams p;
    char eb[512] = {0};
    if (!v) return set_err(nullptr, "v is NULL");
    if (derive_params(cfg, &p, eb, sizeof eb)) return set_err(nullptr, eb);
    const bool kfilter = p.mode >= MODE_KFILTER;
    const int64_t out[16] = {p.k, p.kbig, p.mink, p.useShortKmers, p.maskMiddle, p.midMaskLen, p.minlen, p.minlen2,
                             p.minminlen, p.forbidNs, p.hammingDistance, p.hammingDistance2, (int64_t)p.middleMask,
                             (int64_t)p.mask, kfilter, p.removePairsIfEitherBad};
    memcpy(v, out, sizeof out);
    return 0;
}

const char *bbduk_b200_last_error(bbduk_handle *h) {
    if (h) {
        std::lock_guard<std::mutex> g(h->err_mu);
        g_err = h->err;
    }
    return g_err.c_str();
}

int bbduk_b200_create(const bbduk_cfg *cfg, bbduk_handle **out) {
    if (!out) return set_err(nullptr, "out is NULL");
    *out = nullptr;
    BBParams p;
    char eb[512] = {0};
    if (derive_params(cfg, &p, eb, sizeof eb)) return set_err(nullptr, eb);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return set_err(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                    " (libbbduk_b200 has no CPU fallback)");
    int dev = cfg->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= ndev) return set_err(nullptr, "cfg.device out of range");
    bbduk_handle *h = new bbduk_handle();
    h->cfg = *cfg;
    h->p = p;
    h->device = dev;
    if (cudaSetDevice(dev) != cudaSuccess) {
        delete h;
        return set_err(nullptr, "cudaSetDevice failed");
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("BBDUK_B200_PACK_HOST")) h->pack_host = atoi(e) != 0;
    if (const char *e = getenv("BBDUK_B200_SPARSE_D")) h->sparse_d = atoi(e) != 0;
    if (const char *e = getenv("BBDUK_B200_TRACE")) h->trace = atoi(e) != 0;
    if (const char *e = getenv("BBDUK_B200_ASCII_EVERY")) {
        h->ascii_every = std::max(0, atoi(e));
        h->ascii_every_set = true;
    }
    if (const char *e = getenv("BBDUK_B200_PCIE_GBS")) {
        if (atof(e) > 1.0) h->pcie_gbs = atof(e);
    }
    *out = h;
    return 0;
}

int bbduk_b200_add_ref(bbduk_handle *h, const uint8_t *bases, const int64_t *offsets, int32_t n_seqs) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (h->finalized) return set_err(h, "add_ref after finalize");
    if (n_seqs < 0 || (n_seqs > 0 && (!bases || !offsets))) return set_err(h, "bad add_ref arguments");
    for (int32_t s = 0; s < n_seqs; s++) {
        const int64_t a = offsets[s], b = offsets[s + 1];
        if (b < a) return set_err(h, "offsets must be non-decreasing");
        for (int64_t i = a; i < b; i++)
            if (bases[i] >= 128) return set_err(h, "reference contains a non-ASCII byte (the reference tool would throw)");
        h->ref.insert(h->ref.end(), bases + a, bases + b);
        h->ref_off.push_back((int64_t)h->ref.size());
    }
    return 0;
}

static int alloc_counters(bbduk_handle *h) {
    const size_t n = (size_t)h->table.n_scaffolds + 1;
    cudaFree(h->d_scaf_reads);
    cudaFree(h->d_scaf_bases);
    cudaFree(h->d_stats);
    CKH(cudaMalloc(&h->d_scaf_reads, sizeof(unsigned long long) * n));
    CKH(cudaMalloc(&h->d_scaf_bases, sizeof(unsigned long long) * n));
    CKH(cudaMalloc(&h->d_stats, sizeof(bbduk_stats)));
    CKH(cudaMemset(h->d_scaf_reads, 0, sizeof(unsigned long long) * n));
    CKH(cudaMemset(h->d_scaf_bases, 0, sizeof(unsigned long long) * n));
    CKH(cudaMemset(h->d_stats, 0, sizeof(bbduk_stats)));
    return 0;
}

int bbduk_b200_finalize(bbduk_handle *h, int64_t *stored_kmers) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (h->finalized) {
        if (stored_kmers) *stored_kmers = h->table.stored;
        return 0;
    }
    CKH(cudaSetDevice(h->device));
    char eb[512] = {0};
    int64_t nl = 0;
    // on-chip filter image: sized for the fast kernel's shared memory budget
    const uint32_t filter_words = 40960;  // 160 KB
    if (h->table.build(h->p, h->ref, h->ref_off, h->cfg.table_load_pct, filter_words, nullptr, &nl, eb, sizeof eb))
        return set_err(h, eb);
    h->launches += nl;
    if (alloc_counters(h)) return 1;
    h->finalized = true;
    std::vector<uint8_t>().swap(h->ref);
    if (stored_kmers) *stored_kmers = h->table.stored;
    return 0;
}

int bbduk_b200_table_describe(bbduk_handle *h, bbduk_table_desc *d) {
    if (!h || !d) return set_err(h, "NULL argument");
    if (!h->finalized) return set_err(h, "table_describe before finalize");
    memset(d, 0, sizeof *d);
    d->n_slots = h->table.n_slots;
    d->n_filter_words = h->table.total_filter_words();
    d->stored_kmers = h->table.stored;
    d->n_scaffolds = h->table.n_scaffolds;
    d->d_keys = h->table.d_keys;
    d->d_vals = h->table.d_vals;
    d->d_filter = h->table.d_filter;
    d->scalars[0] = h->table.ref_kmers;
    d->scalars[1] = h->table.n_filter_words;
    d->scalars[2] = h->table.part_words;
    d->scalars[3] = h->table.short_words;
    d->scalars[4] = h->table.n_parts | ((int64_t)h->table.part_w << 8);
    d->scalars[5] = (int64_t)h->table.part_lag[0] | ((int64_t)h->table.part_lag[1] << 8) |
                    ((int64_t)h->table.part_lag[2] << 16) | ((int64_t)h->table.part_lag[3] << 24);
    d->scalars[6] = h->table.big_words;
    d->scalars[7] = (int64_t)h->table.samp_words | ((int64_t)h->table.tail_words << 24) | ((int64_t)h->table.tail_q << 56);
    return 0;
}

int bbduk_b200_table_alloc(bbduk_handle *h, bbduk_table_desc *d) {
    if (!h || !d) return set_err(h, "NULL argument");
    if (h->finalized) return set_err(h, "table_alloc on a finalized handle");
    if (d->n_slots < 1024 || (d->n_slots & (d->n_slots - 1))) return set_err(h, "n_slots must be a power of two >= 1024");
    CKH(cudaSetDevice(h->device));
    char eb[512] = {0};
    if (h->table.alloc(d->n_slots, (uint32_t)d->n_filter_words, eb, sizeof eb)) return set_err(h, eb);
    h->table.stored = d->stored_kmers;
    h->table.n_scaffolds = d->n_scaffolds;
    h->table.ref_kmers = d->scalars[0];
    h->table.n_filter_words = (uint32_t)d->scalars[1];
    h->table.part_words = (uint32_t)d->scalars[2];
    h->table.short_words = (uint32_t)d->scalars[3];
    h->table.big_words = (uint32_t)d->scalars[6];
    h->table.samp_words = (uint32_t)(d->scalars[7] & 0xFFFFFF);
    h->table.tail_words = (uint32_t)((d->scalars[7] >> 24) & 0xFFFFFFFFll);
    h->table.tail_q = (int32_t)((d->scalars[7] >> 56) & 0x7F);
    h->table.n_parts = (int32_t)(d->scalars[4] & 0xFF);
    h->table.part_w = (int32_t)(d->scalars[4] >> 8);
    for (int j = 0; j < 4; j++) h->table.part_lag[j] = (int32_t)((d->scalars[5] >> (8 * j)) & 0xFF);
    if ((int64_t)h->table.total_filter_words() != d->n_filter_words) return set_err(h, "inconsistent filter geometry");
    d->d_keys = h->table.d_keys;
    d->d_vals = h->table.d_vals;
    d->d_filter = h->table.d_filter;
    return 0;
}

int bbduk_b200_table_commit(bbduk_handle *h) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (h->finalized) return 0;
    if (!h->table.d_keys) return set_err(h, "table_commit without table_alloc");
    CKH(cudaSetDevice(h->device));
    if (alloc_counters(h)) return 1;
    h->finalized = true;
    return 0;
}

// max_len_known > 0: the caller knows the longest read of the batch (the chain measures every chunk on the host);
// 0: the handle-wide hint, else one reduction kernel + a stream synchronisation
static int process_device_impl(bbduk_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads,
                               int32_t paired, const bbduk_out *d_out, bbduk_stats *d_stats, void *stream, int max_len_known,
                               const uint32_t *pk_F = nullptr, const uint16_t *pk_D = nullptr) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!h->finalized) return set_err(h, "process before finalize");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads");
    if (check_mode_inputs(h, paired, d_out)) return 1;
    if (n_reads == 0) return 0;
    CKH(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(h->dev_mu);
    const int64_t n_units = paired ? n_reads / 2 : n_reads;
    if (n_units + 8 > h->dev_handoff_cap) {
        cudaFree(h->dev_handoff);
        cudaFree(h->dev_handoff_n);
        h->dev_handoff_cap = n_units + n_units / 8 + 1024;
        CKH(cudaMalloc(&h->dev_handoff, sizeof(int32_t) * h->dev_handoff_cap));
        CKH(cudaMalloc(&h->dev_handoff_n, sizeof(unsigned int) * 4));
    }
    // longest read (sizes the fast kernel's staging): caller's hint, else one reduction + sync
    unsigned int mx = max_len_known > 0 ? (unsigned int)max_len_known : (unsigned int)std::max(0, h->max_read_len_hint.load());
    if (mx == 0) {
        CKH(cudaMemsetAsync(h->dev_handoff_n + 1, 0, sizeof(unsigned int), st));
        max_len_kernel<<<296, 256, 0, st>>>(d_offsets, n_reads, h->dev_handoff_n + 1);
        h->launches += 1;
        CKH(cudaMemcpyAsync(&mx, h->dev_handoff_n + 1, sizeof mx, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    DirectScratch ds{nullptr, nullptr, nullptr};
    int64_t n_bases = -1;
    if (uses_direct(h, (int)mx)) {
        // the flat scan needs the batch's total base count: one 4-byte read-back
        uint32_t last = 0;
        CKH(cudaMemcpyAsync(&last, d_offsets + n_reads, sizeof last, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
        n_bases = (int64_t)last;
        if (reinterpret_cast<uintptr_t>(d_bases) & 15) return set_err(h, "d_bases must be 16-byte aligned");
        if (n_reads + 8 > h->dev_first_cap) {
            cudaFree(h->dev_first64);
            cudaFree(h->dev_lastpos);
            h->dev_first_cap = n_reads + n_reads / 8 + 1024;
            CKH(cudaMalloc(&h->dev_first64, sizeof(unsigned long long) * h->dev_first_cap));
            CKH(cudaMalloc(&h->dev_lastpos, sizeof(int) * h->dev_first_cap));
        }
        if ((int64_t)direct_sbits_bytes(n_bases) > h->dev_sbits_cap) {
            cudaFree(h->dev_sbits);
            h->dev_sbits_cap = (int64_t)direct_sbits_bytes(n_bases + n_bases / 8 + 4096);
            CKH(cudaMalloc(&h->dev_sbits, (size_t)h->dev_sbits_cap));
        }
        ds = DirectScratch{h->dev_first64, h->dev_lastpos, h->dev_sbits};
    }
    if (pk_F) {  // the tuned kernels read the 2-bit stream directly when no tile can be handed off; else the ASCII copy serves
        const BBTable tv = h->table.view();
        const bool ok = pk_D && (int)mx <= FAST_MAX_READ_LEN && (plan_fast2(h->p, tv, (int)mx).usable || plan_fast(h->p, tv, (int)mx).usable) &&
                        packed_ok(h->p, tv) && !(d_out->maskbits);
        if (!ok) pk_F = nullptr, pk_D = nullptr;
    }
    return run_batch(h, d_bases, d_offsets, n_reads, n_bases, paired, *d_out, d_stats, (int)mx, h->dev_handoff,
                     h->dev_handoff_n, ds, st, pk_F, pk_D);
}

int bbduk_b200_process_device(bbduk_handle *h, const uint8_t *d_bases, const uint32_t *d_offsets, int64_t n_reads,
                              int32_t paired, const bbduk_out *d_out, bbduk_stats *d_stats, void *stream) {
    return process_device_impl(h, d_bases, d_offsets, n_reads, paired, d_out, d_stats, stream, 0);
}

// bases != NULL: ASCII input (bbduk_b200_process). Otherwise pre_F / pre_D: the caller's own 2-bit stream + defined bits over the
// concatenated bases (bbduk_b200_process_packed); a chunk then starts at the 16-base group its first read begins in.
// Defined bits of a packed chunk onto the device. cnt != NULL: the workers listed the groups with an undefined base in their
// regions of s.h_exc (R entries each); if none overflowed, the device array is a memset to all ones + those exceptions
// (0.05 % undefined bases: 1 % of the groups, 8 bytes each, instead of 2 bytes for every group). Otherwise the array itself
// travels: from `direct` (the caller's page-locked array) if given, else from s.h_D (if `fill` is given it is copied there
// first -- a worker skipped its share because it expected the exception path).
static int upload_defined(bbduk_handle *h, Slot &s, const int64_t *cnt, int n_parts, int64_t R, int64_t groups, const uint16_t *direct,
                          const uint16_t *fill, cudaStream_t st) {
    bool sparse = cnt != nullptr;
    int64_t total = 0;
    if (sparse)
        for (int p = 0; p < n_parts; p++) {
            if (cnt[p] < 0) sparse = false;
            else total += cnt[p];
        }
    if (sparse) {
        int64_t at = 0;
        for (int p = 0; p < n_parts; p++) {
            if (cnt[p] > 0 && at != (int64_t)p * R) memmove(s.h_exc + at, s.h_exc + (int64_t)p * R, sizeof(uint64_t) * (size_t)cnt[p]);
            at += cnt[p];
        }
        CKH(cudaMemsetAsync(s.d_D, 0xFF, sizeof(uint16_t) * (size_t)groups, st));
        if (total > 0) {
            h->h2d_bytes += (int64_t)(sizeof(uint64_t) * total);
            CKH(cudaMemcpyAsync(s.d_exc, s.h_exc, sizeof(uint64_t) * (size_t)total, cudaMemcpyHostToDevice, st));
            defined_exceptions_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s.d_exc, total, s.d_D);
            h->launches += 1;
        }
        return 0;
    }
    if (!direct && fill && cnt) memcpy(s.h_D, fill, sizeof(uint16_t) * (size_t)groups);  // rare: dense undefined bases in a staged chunk
    h->h2d_bytes += (int64_t)(sizeof(uint16_t) * groups);
    CKH(cudaMemcpyAsync(s.d_D, direct ? direct : s.h_D, sizeof(uint16_t) * (size_t)groups, cudaMemcpyHostToDevice, st));
    return 0;
}

// the handle's host worker pool (caller holds h->pool_mu)
static void ensure_pool(bbduk_handle *h) {
    if (h->pool) return;
    // one process per GPU: share the host cores among the ranks of this node (torchrun exports the count)
    int share = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = std::max(1, atoi(e));
    if (const char *e = getenv("BBDUK_B200_HOST_THREADS")) share = -atoi(e);
    const int hc = (int)std::thread::hardware_concurrency();
    h->pool = new HostPool(share < 0 ? std::max(1, -share) : std::max(1, std::min(32, hc / share)));
    // few packing workers per GPU (many ranks on one host): lean on PCIe more
    if (!h->ascii_every_set) h->ascii_every = h->pool->size() >= 12 ? 3 : h->pool->size() >= 6 ? 2 : 1;
    // four workers or fewer per GPU (eight ranks on a 32-core host): packing loses to plain DMA whatever the
    // share (tools/e2e_mix_sweep.py, profiles/r02p_e2e_mix_sweep_8gpu.jsonl), so the share stays fixed
    if (!h->ascii_every_set && h->pool->size() < 6) h->ascii_every_set = true;
}

static int process_host_impl(bbduk_handle *h, const uint8_t *bases, const uint32_t *pre_F, const uint16_t *pre_D, const int64_t *offsets,
                             int64_t n_reads, int32_t paired, const bbduk_out *out, bbduk_stats *stats) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!h->finalized) return set_err(h, "process before finalize");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    const bool pre = bases == nullptr;
    if (n_reads > 0 && ((!bases && !(pre_F && pre_D)) || !offsets)) return set_err(h, "NULL input");
    if (check_mode_inputs(h, paired, out)) return 1;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_reads == 0) return 0;
    const auto t_call0 = std::chrono::steady_clock::now();
    CKH(cudaSetDevice(h->device));
    const bool want_mask = h->p.mode == MODE_KMASK && out->maskbits && out->mask_off;

    // private device-side stats for this call so concurrent callers do not mix
    bbduk_stats *d_stats = nullptr;
    if (stats) {
        CKH(cudaMalloc(&d_stats, sizeof(bbduk_stats)));
        CKH(cudaMemset(d_stats, 0, sizeof(bbduk_stats)));
    }
    int rc = 0;
    int chunk_no = 0;
    // ASCII chunks are DMA'd straight from the caller's buffer: only worth it when that buffer is page-locked (a Java
    // array reached through GetPrimitiveArrayCritical is not; then every chunk is packed into the pinned staging)
    bool src_pinned = false;
    {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, pre ? (const void *)pre_F : (const void *)bases) == cudaSuccess) src_pinned = pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pre && src_pinned) {
            src_pinned = cudaPointerGetAttributes(&pa, pre_D) == cudaSuccess && pa.type == cudaMemoryTypeHost;
            cudaGetLastError();
        }
    }
    std::vector<Slot *> used;
    const int per = paired ? 2 : 1;
    int64_t r0 = 0;
    while (r0 < n_reads && !rc) {
        // chunk [r0, r1): bounded reads and bytes, pairs never split
        int64_t r1 = std::min(n_reads, r0 + CHUNK_READS);
        while (r1 > r0 + per && offsets[r1] - offsets[r0] > CHUNK_BYTES) r1 = r0 + std::max<int64_t>(per, ((r1 - r0) / 2 / per) * per);
        // packed input: the chunk's stream starts at the group its first read begins in, so offsets are rebased to that group
        const int64_t gfirst = pre ? (offsets[r0] >> 4) : 0;
        const int64_t obase = pre ? 16 * gfirst : offsets[r0];
        const int64_t nb = offsets[r1] - obase;
        if (nb >= (1ll << 32) - 64) {
            rc = set_err(h, "a single read (pair) exceeds 4 GiB");
            break;
        }
        const int64_t nr = r1 - r0;
        const int64_t mw0 = want_mask ? out->mask_off[r0] : 0, mw = want_mask ? out->mask_off[r1] - mw0 : 0;
        Slot &s = h->slots[h->next_slot++ % N_SLOTS];
        std::lock_guard<std::mutex> g(s.mu);
        const auto t_wait0 = std::chrono::steady_clock::now();
        if (s.done) cudaEventSynchronize(s.done);  // previous use of this slot has drained
        if (h->trace)
            fprintf(stderr, "[bbduk_b200] slot wait %.3f ms\n",
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wait0).count());
        const bool direct = uses_direct(h, 1 << 20);  // conservative: decided again per chunk in run_batch
        if ((rc = ensure_slot(h, s, nr, nb, mw, direct))) break;
        // longest read of the chunk: one pass over the offsets, shared among the host workers
        int max_len = 0;
        {
            std::lock_guard<std::mutex> pg(h->pool_mu);
            ensure_pool(h);
            std::atomic<int> mx{0};
            std::atomic<bool> bad{false};
            const int64_t *osrc = offsets + r0;
            h->pool->run([&](int part, int n_parts) {
                const int64_t i0 = nr * part / n_parts, i1 = nr * (part + 1) / n_parts;
                int64_t m = 0;
                for (int64_t i = i0; i < i1; i++) {
                    const int64_t l = osrc[i + 1] - osrc[i];
                    if (l < 0 || l > 0x7FFFFFFF) bad = true;
                    if (l > m) m = l;
                }
                int cur = mx.load();
                while ((int)std::min<int64_t>(m, 0x7FFFFFFF) > cur && !mx.compare_exchange_weak(cur, (int)std::min<int64_t>(m, 0x7FFFFFFF))) {
                }
            });
            if (bad) rc = set_err(h, "bad read length");
            max_len = mx.load();
        }
        if (rc) break;
        cudaStream_t st = s.st;
#define CKL(call)                                                                                          \
    if (!rc) {                                                                                             \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            rc = set_err(h, b_);                                                                           \
        }                                                                                                  \
    }
        // PCIe leg: when the tuned kernel takes the whole chunk, the bases cross as 2-bit codes + defined bits
        // (0.375 B/base instead of 1) packed by the host workers, and the offsets as 32-bit words
        bool packed = false;
        {
            const BBTable tv = h->table.view();
            packed = (pre || (h->pack_host && nb >= (1 << 16))) && !want_mask && max_len <= FAST_MAX_READ_LEN &&
                     (plan_fast2(h->p, tv, max_len).usable || plan_fast(h->p, tv, max_len).usable) && packed_ok(h->p, tv);
            // host packing is bound by the host's memory bandwidth, the ASCII path by PCIe: every n-th chunk goes
            // ASCII (no CPU work, DMA straight from the caller's buffer) so that both resources are used
            // (never the last chunk of a call: its transfer is the tail nothing overlaps with)
            if (!pre && packed && src_pinned && r1 < n_reads) {
                if (h->ascii_every_set || h->pack_s_per_base <= 0.0) {  // fixed share (or nothing measured yet)
                    if (h->ascii_every > 0 && (chunk_no % h->ascii_every) == h->ascii_every - 1) packed = false;
                } else {
                    // a packed base costs tp of the workers and cp bytes of the link (2 bits + its defined bit, or next to
                    // nothing for the latter when only the exceptions travel); an ASCII base costs one byte of the link
                    const double tp = h->pack_s_per_base, tb = 1.0 / (h->pcie_gbs * 1e9), cp = h->sparse_d ? 0.26 : 0.375;
                    const double x = std::min(0.75, std::max(0.0, (tp - cp * tb) / (tp + (1.0 - cp) * tb)));
                    h->ascii_acc += x;
                    if (h->ascii_acc >= 1.0) {
                        h->ascii_acc -= 1.0;
                        packed = false;
                    }
                }
            }
            chunk_no++;
        }
        if ((packed || pre) && !rc) rc = ensure_packed(h, s, nr, nb);
        if (pre && !rc) {
            // the caller's stream: rebase the offsets (and, when the arrays are not page-locked, stage them) on the workers
            std::lock_guard<std::mutex> pg(h->pool_mu);
            const int64_t *osrc = offsets + r0;
            const int64_t groups = (nb + 15) / 16;
            Slot *sp = &s;
            const bool stage = !src_pinned;
            const uint32_t *fsrc = pre_F + gfirst;
            const uint16_t *dsrc = pre_D + gfirst;
            std::vector<int64_t> cnt(h->pool->size(), 0);
            int64_t *cntp = cnt.data();
            const int64_t R = s.cap_exc / h->pool->size();
            const bool sparse = h->sparse_d;
            h->pool->run([=](int part, int n_parts) {
                const int64_t i0 = (nr + 1) * part / n_parts, i1 = (nr + 1) * (part + 1) / n_parts;
                for (int64_t i = i0; i < i1; i++) sp->h_off32[i] = (uint32_t)(osrc[i] - obase);
                const int64_t g0 = groups * part / n_parts, g1 = groups * (part + 1) / n_parts;
                if (stage) memcpy(sp->h_F + g0, fsrc + g0, sizeof(uint32_t) * (size_t)(g1 - g0));
                if (sparse) cntp[part] = list_undefined_groups(dsrc, g0, g1, sp->h_exc + part * R, R);
                if (stage && (!sparse || cntp[part] < 0)) memcpy(sp->h_D + g0, dsrc + g0, sizeof(uint16_t) * (size_t)(g1 - g0));
            });
            CKL((h->h2d_bytes += (int64_t)(sizeof(uint32_t) * groups), cudaMemcpyAsync(s.d_F, stage ? s.h_F : fsrc, sizeof(uint32_t) * groups, cudaMemcpyHostToDevice, st)));
            if (!rc) rc = upload_defined(h, s, sparse ? cntp : nullptr, (int)cnt.size(), R, groups, stage ? nullptr : dsrc, dsrc, st);
            CKL((h->h2d_bytes += (int64_t)(sizeof(uint32_t) * (nr + 1)), cudaMemcpyAsync(s.d_off32, s.h_off32, sizeof(uint32_t) * (nr + 1), cudaMemcpyHostToDevice, st)));
            if (!packed && !rc) {  // a mode the tuned kernels do not serve: spell the stream out on the device
                unpack_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(s.d_F, s.d_D, s.d_bases, groups);
                h->launches += 1;
            }
        } else if (packed && !rc) {
            std::lock_guard<std::mutex> pg(h->pool_mu);
            const uint8_t *src = bases + offsets[r0];
            const int64_t *osrc = offsets + r0;
            const int64_t groups = (nb + 15) / 16;
            Slot *sp = &s;
            std::vector<int64_t> cnt(h->pool->size(), 0);
            int64_t *cntp = cnt.data();
            const int64_t R = s.cap_exc / h->pool->size();
            const bool sparse = h->sparse_d;
            const auto t_pack0 = std::chrono::steady_clock::now();
            h->pool->run([=](int part, int n_parts) {
                // worker ranges start on 512-base boundaries so that the packer can stream whole cache lines
                const int64_t gb = (groups + 31) / 32;
                const int64_t g0 = std::min(groups, gb * part / n_parts * 32), g1 = std::min(groups, gb * (part + 1) / n_parts * 32);
                if (sparse) cntp[part] = pack_bases_range_listing(src, nb, g0, g1, sp->h_F, sp->h_D, sp->h_exc + part * R, R);
                else pack_bases_range(src, nb, g0, g1, sp->h_F, sp->h_D);
                const int64_t i0 = (nr + 1) * part / n_parts, i1 = (nr + 1) * (part + 1) / n_parts;
                for (int64_t i = i0; i < i1; i++) sp->h_off32[i] = (uint32_t)(osrc[i] - osrc[0]);
            });
            {
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pack0).count();
                if (nb > (1 << 20)) h->pack_s_per_base = h->pack_s_per_base <= 0.0 ? dt / (double)nb : 0.75 * h->pack_s_per_base + 0.25 * dt / (double)nb;
                if (h->trace)
                    fprintf(stderr, "[bbduk_b200] packed %lld bases in %.3f ms on %d threads\n", (long long)nb, 1e3 * dt, h->pool->size());
            }
            CKL((h->h2d_bytes += (int64_t)(sizeof(uint32_t) * groups), cudaMemcpyAsync(s.d_F, s.h_F, sizeof(uint32_t) * groups, cudaMemcpyHostToDevice, st)));
            if (!rc) rc = upload_defined(h, s, sparse ? cntp : nullptr, (int)cnt.size(), R, groups, nullptr, nullptr, st);
            CKL((h->h2d_bytes += (int64_t)(sizeof(uint32_t) * (nr + 1)), cudaMemcpyAsync(s.d_off32, s.h_off32, sizeof(uint32_t) * (nr + 1), cudaMemcpyHostToDevice, st)));
        } else {
            CKL((h->h2d_bytes += (int64_t)((size_t)nb), cudaMemcpyAsync(s.d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st)));
            CKL((h->h2d_bytes += (int64_t)(sizeof(int64_t) * (nr + 1)), cudaMemcpyAsync(s.d_off64, offsets + r0, sizeof(int64_t) * (nr + 1), cudaMemcpyHostToDevice, st)));
            if (!rc) {
                off64_to_32_kernel<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(s.d_off64, offsets[r0], s.d_off32, nr + 1);
                h->launches += 1;
            }
        }
        bbduk_out dout;
        memset(&dout, 0, sizeof dout);
        dout.id0 = out->id0 ? s.d_id0 : nullptr;
        dout.id0b = out->id0b ? s.d_id0b : nullptr;
        dout.lo = out->lo ? s.d_lo : nullptr;
        dout.hi = out->hi ? s.d_hi : nullptr;
        dout.flags = out->flags ? s.d_flags : nullptr;
        dout.count = out->count ? s.d_count : nullptr;
        if (want_mask) {
            CKL((h->h2d_bytes += (int64_t)(sizeof(int64_t) * (nr + 1)), cudaMemcpyAsync(s.d_maskoff, out->mask_off + r0, sizeof(int64_t) * (nr + 1), cudaMemcpyHostToDevice, st)));
            if (!rc) {
                maskoff_rebase_kernel<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(s.d_maskoff, mw0, nr + 1);
                h->launches += 1;
            }
            dout.maskbits = s.d_maskbits;
            dout.mask_off = s.d_maskoff;
        }
        if (!rc)
            rc = run_batch(h, s.d_bases, s.d_off32, nr, nb, paired, dout, d_stats, max_len, s.d_handoff, s.d_handoff_n,
                           DirectScratch{s.d_first64, s.d_lastpos, s.d_sbits}, st, packed ? s.d_F : nullptr,
                           packed ? s.d_D : nullptr);
        if (out->id0) CKL((h->d2h_bytes += (int64_t)(sizeof(int32_t) * nr), cudaMemcpyAsync(out->id0 + r0, s.d_id0, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st)));
        if (out->id0b) CKL((h->d2h_bytes += (int64_t)(sizeof(int32_t) * nr), cudaMemcpyAsync(out->id0b + r0, s.d_id0b, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st)));
        if (out->lo) CKL((h->d2h_bytes += (int64_t)(sizeof(int32_t) * nr), cudaMemcpyAsync(out->lo + r0, s.d_lo, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st)));
        if (out->hi) CKL((h->d2h_bytes += (int64_t)(sizeof(int32_t) * nr), cudaMemcpyAsync(out->hi + r0, s.d_hi, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st)));
        if (out->count) CKL((h->d2h_bytes += (int64_t)(sizeof(int32_t) * nr), cudaMemcpyAsync(out->count + r0, s.d_count, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st)));
        if (out->flags) CKL((h->d2h_bytes += (int64_t)((size_t)nr), cudaMemcpyAsync(out->flags + r0, s.d_flags, (size_t)nr, cudaMemcpyDeviceToHost, st)));
        if (want_mask && mw > 0)
            CKL((h->d2h_bytes += (int64_t)(sizeof(uint32_t) * mw), cudaMemcpyAsync(out->maskbits + mw0, s.d_maskbits, sizeof(uint32_t) * mw, cudaMemcpyDeviceToHost, st)));
        CKL(cudaEventRecord(s.done, st));
#undef CKL
        used.push_back(&s);
        r0 = r1;
    }
    const auto t_sync0 = std::chrono::steady_clock::now();
    for (Slot *s : used) cudaStreamSynchronize(s->st);
    if (h->trace)
        fprintf(stderr, "[bbduk_b200] process: enqueue %.3f ms, final sync %.3f ms\n",
                1e3 * std::chrono::duration<double>(t_sync0 - t_call0).count(),
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sync0).count());
    if (!rc && stats) {
        if (cudaMemcpy(stats, d_stats, sizeof *stats, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(h, "stats copy failed");
    }
    cudaFree(d_stats);
    if (!rc) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = set_err(h, std::string("CUDA error after process: ") + cudaGetErrorString(e));
    }
    return rc;
}

int bbduk_b200_process(bbduk_handle *h, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int32_t paired,
                       const bbduk_out *out, bbduk_stats *stats) {
    if (h && n_reads > 0 && !bases) return set_err(h, "NULL input");
    static const uint8_t none = 0;
    return process_host_impl(h, bases ? bases : &none, nullptr, nullptr, offsets, n_reads, paired, out, stats);
}

int bbduk_b200_process_packed(bbduk_handle *h, const uint32_t *F, const uint16_t *D, const int64_t *offsets, int64_t n_reads,
                              int32_t paired, const bbduk_out *out, bbduk_stats *stats) {
    if (h && n_reads > 0 && (!F || !D)) return set_err(h, "NULL input");
    if (h && h->finalized && h->p.mode == MODE_KMASK) return set_err(h, "packed input does not carry the bases' case: kmask needs the ASCII entry");
    static const uint32_t nf = 0;
    static const uint16_t nd = 0;
    return process_host_impl(h, nullptr, F ? F : &nf, D ? D : &nd, offsets, n_reads, paired, out, stats);
}

int bbduk_b200_scaffold_counts(bbduk_handle *h, int64_t *read_counts, int64_t *base_counts, int32_t n) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!h->finalized) return set_err(h, "scaffold_counts before finalize");
    CKH(cudaSetDevice(h->device));
    CKH(cudaDeviceSynchronize());
    const int32_t m = std::min(n, h->table.n_scaffolds + 1);
    if (m <= 0) return 0;
    if (read_counts) CKH(cudaMemcpy(read_counts, h->d_scaf_reads, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
    if (base_counts) CKH(cudaMemcpy(base_counts, h->d_scaf_bases, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
    return 0;
}

int bbduk_b200_set_max_read_len(bbduk_handle *h, int32_t max_read_len) {
    if (!h) return set_err(nullptr, "handle is NULL");
    h->max_read_len_hint = max_read_len < 0 ? 0 : max_read_len;
    return 0;
}

void bbduk_b200_tbo_cfg_default(bbduk_tbo_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->strict_overlap = 1;
    c->min_overlap0 = c->min_overlap = c->min_insert0 = c->min_insert = -1;
    c->qual_offset = 33;
}

int bbduk_b200_tbo_device(bbduk_handle *h, const bbduk_tbo_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
                          const uint32_t *d_offsets, int64_t n_reads, int32_t max_read_len, const int32_t *d_lo, int32_t *d_hi,
                          uint8_t *d_flags, int32_t *d_insert, int64_t *d_stats2, void *stream) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_tbo_cfg");
    if (n_reads < 0 || (n_reads & 1)) return set_err(h, "tbo needs paired reads (an even count)");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets || !d_lo || !d_hi || !d_flags) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    const int rc = launch_tbo(h->device, h->sm_count, cfg, d_bases, d_quals, d_offsets, n_reads, max_read_len, d_lo, d_hi, d_flags,
                              d_insert, reinterpret_cast<unsigned long long *>(d_stats2), (cudaStream_t)stream);
    if (rc == 2) return set_err(h, "tbo: a read is longer than 1008 bases (no device path, and no CPU fallback)");
    if (rc) return set_err(h, std::string("tbo kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->launches += 4;
    return 0;
}

int bbduk_b200_tbo(bbduk_handle *h, const bbduk_tbo_cfg *cfg, const uint8_t *bases, const uint8_t *quals, const int64_t *offsets,
                   int64_t n_reads, const int32_t *lo, int32_t *hi, uint8_t *flags, int32_t *insert, int64_t *stats2) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (n_reads < 0 || (n_reads & 1)) return set_err(h, "tbo needs paired reads (an even count)");
    if (n_reads == 0) return 0;
    if (!bases || !offsets || !lo || !hi || !flags) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> g(h->tbo_mu);
    cudaStream_t st = nullptr;  // the legacy default stream: this entry point is synchronous
    int64_t *d_stats = nullptr;
    CKH(cudaMalloc(&d_stats, 2 * sizeof(int64_t)));
    CKH(cudaMemset(d_stats, 0, 2 * sizeof(int64_t)));
    int rc = 0;
    int64_t r0 = 0;
    std::vector<uint32_t> off32;
    while (r0 < n_reads && !rc) {
        int64_t r1 = std::min(n_reads, r0 + (CHUNK_READS << 1));
        while (r1 > r0 + 2 && offsets[r1] - offsets[r0] > CHUNK_BYTES) r1 = r0 + std::max<int64_t>(2, ((r1 - r0) / 4) * 2);
        const int64_t nr = r1 - r0, nb = offsets[r1] - offsets[r0];
        if (nb < 0 || nb >= (1ll << 32) - 64) {
            rc = set_err(h, "a pair exceeds 4 GiB (or offsets decrease)");
            break;
        }
        int max_len = 0;
        off32.resize(nr + 1);
        for (int64_t i = 0; i <= nr; i++) off32[i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        for (int64_t i = 0; i < nr; i++) max_len = std::max(max_len, (int)(hi[r0 + i] - lo[r0 + i]));
        auto need = [&](void **p, int64_t *cap, int64_t bytes) -> int {
            if (bytes <= *cap) return 0;
            cudaFree(*p);
            *p = nullptr;
            *cap = bytes + bytes / 8 + 4096;
            return cudaMalloc(p, (size_t)*cap) == cudaSuccess ? 0 : 1;
        };
        auto &tb = h->tbo;
        if (need((void **)&tb.d_bases, &tb.cap_bases, nb + 64) || (quals && need((void **)&tb.d_quals, &tb.cap_quals, nb + 64)) ||
            need((void **)&tb.d_off, &tb.cap_off, 4 * (nr + 1)) || need((void **)&tb.d_lo, &tb.cap_lo, 4 * nr) ||
            need((void **)&tb.d_hi, &tb.cap_hi, 4 * nr) || need((void **)&tb.d_flags, &tb.cap_flags, nr) ||
            need((void **)&tb.d_insert, &tb.cap_insert, 2 * nr + 8)) {
            rc = set_err(h, "tbo: device allocation failed");
            break;
        }
#define CKT(call)                                                                       \
    if (!rc && (call) != cudaSuccess) rc = set_err(h, std::string(#call " failed: ") + cudaGetErrorString(cudaGetLastError()))
        CKT(cudaMemcpyAsync(tb.d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st));
        if (quals) CKT(cudaMemcpyAsync(tb.d_quals, quals + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(tb.d_off, off32.data(), 4 * (size_t)(nr + 1), cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(tb.d_lo, lo + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(tb.d_hi, hi + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(tb.d_flags, flags + r0, (size_t)nr, cudaMemcpyHostToDevice, st));
        if (!rc)
            rc = bbduk_b200_tbo_device(h, cfg, tb.d_bases, quals ? tb.d_quals : nullptr, tb.d_off, nr, max_len, tb.d_lo, tb.d_hi,
                                       tb.d_flags, tb.d_insert, d_stats, st);
        CKT(cudaMemcpyAsync(hi + r0, tb.d_hi, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKT(cudaMemcpyAsync(flags + r0, tb.d_flags, (size_t)nr, cudaMemcpyDeviceToHost, st));
        if (insert) CKT(cudaMemcpyAsync(insert + r0 / 2, tb.d_insert, 4 * (size_t)(nr / 2), cudaMemcpyDeviceToHost, st));
        CKT(cudaStreamSynchronize(st));
#undef CKT
        r0 = r1;
    }
    if (!rc && stats2) {
        int64_t v[2] = {0, 0};
        if (cudaMemcpy(v, d_stats, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(h, "tbo: stats copy failed");
        stats2[0] += v[0];
        stats2[1] += v[1];
    }
    cudaFree(d_stats);
    return rc;
}

void bbduk_b200_qtrim_cfg_default(bbduk_qtrim_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->trimq = 6.0f;  // jgi/BBDuk.java:126
    c->max_ns = -1;
    c->qual_offset = 33;
    c->max_non_poly = 1;  // parse/Parser.java:1831
    c->max_n_rate = 1.0f; // jgi/BBDuk.java:629: unset = 1 = no-op
    c->window_length = 4;     // shared/TrimRead.java:961
    c->min_good_interval = 2; // shared/TrimRead.java:949
}

static int check_qtrim_cfg(bbduk_handle *h, const bbduk_qtrim_cfg *cfg, const void *quals) {
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_qtrim_cfg");
    (void)quals;  // reads without qualities: trimming falls back to N's, mbq / maq do not apply (as in the reference)
    if (cfg->trim_mode < 0 || cfg->trim_mode > 2 || cfg->window_length < 1 || cfg->min_good_interval < 0)
        return set_err(h, "bad trim_mode / window_length / min_good_interval");
    if (cfg->qual_offset < 0 || cfg->qual_offset > 127) return set_err(h, "bad qual_offset");
    return 0;
}

int bbduk_b200_qtrim_device(bbduk_handle *h, const bbduk_qtrim_cfg *cfg, const uint8_t *d_bases, const uint8_t *d_quals,
                            const uint32_t *d_offsets, int64_t n_reads, int32_t paired, int32_t *d_lo, int32_t *d_hi,
                            uint8_t *d_flags, int64_t *d_stats8, void *stream) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (check_qtrim_cfg(h, cfg, d_quals)) return 1;
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets || !d_lo || !d_hi || !d_flags) return set_err(h, "NULL input");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_quals) & 15))
        return set_err(h, "d_bases and d_quals must be 16-byte aligned");
    CKH(cudaSetDevice(h->device));
    if (launch_qtrim(h->sm_count, cfg, h->p, d_bases, d_quals, d_offsets, n_reads, paired ? 1 : 0, d_lo, d_hi, d_flags,
                     reinterpret_cast<unsigned long long *>(d_stats8), (cudaStream_t)stream))
        return set_err(h, std::string("qtrim kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->launches += 1;
    return 0;
}

int bbduk_b200_qtrim(bbduk_handle *h, const bbduk_qtrim_cfg *cfg, const uint8_t *bases, const uint8_t *quals,
                     const int64_t *offsets, int64_t n_reads, int32_t paired, int32_t *lo, int32_t *hi, uint8_t *flags,
                     int64_t *stats8) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (check_qtrim_cfg(h, cfg, quals)) return 1;
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (n_reads == 0) return 0;
    if (!bases || !offsets || !lo || !hi || !flags) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> g(h->tbo_mu);  // shares the staging of the tbo entry point
    cudaStream_t st = nullptr;
    int64_t *d_stats = nullptr;
    CKH(cudaMalloc(&d_stats, 8 * sizeof(int64_t)));
    CKH(cudaMemset(d_stats, 0, 8 * sizeof(int64_t)));
    int rc = 0;
    int64_t r0 = 0;
    const int per = paired ? 2 : 1;
    std::vector<uint32_t> off32;
    while (r0 < n_reads && !rc) {
        int64_t r1 = std::min(n_reads, r0 + (CHUNK_READS << 1));
        while (r1 > r0 + per && offsets[r1] - offsets[r0] > CHUNK_BYTES) r1 = r0 + std::max<int64_t>(per, ((r1 - r0) / 2 / per) * per);
        const int64_t nr = r1 - r0, nb = offsets[r1] - offsets[r0];
        if (nb < 0 || nb >= (1ll << 32) - 64) {
            rc = set_err(h, "a read (pair) exceeds 4 GiB (or offsets decrease)");
            break;
        }
        off32.resize(nr + 1);
        for (int64_t i = 0; i <= nr; i++) off32[i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        auto need = [&](void **p, int64_t *cap, int64_t bytes) -> int {
            if (bytes <= *cap) return 0;
            cudaFree(*p);
            *p = nullptr;
            *cap = bytes + bytes / 8 + 4096;
            return cudaMalloc(p, (size_t)*cap) == cudaSuccess ? 0 : 1;
        };
        auto &tb = h->tbo;
        if (need((void **)&tb.d_bases, &tb.cap_bases, nb + 64) || (quals && need((void **)&tb.d_quals, &tb.cap_quals, nb + 64)) ||
            need((void **)&tb.d_off, &tb.cap_off, 4 * (nr + 1)) || need((void **)&tb.d_lo, &tb.cap_lo, 4 * nr) ||
            need((void **)&tb.d_hi, &tb.cap_hi, 4 * nr) || need((void **)&tb.d_flags, &tb.cap_flags, nr)) {
            rc = set_err(h, "qtrim: device allocation failed");
            break;
        }
#define CKQ(call)                                                                       \
    if (!rc && (call) != cudaSuccess) rc = set_err(h, std::string(#call " failed: ") + cudaGetErrorString(cudaGetLastError()))
        CKQ(cudaMemcpyAsync(tb.d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st));
        if (quals) CKQ(cudaMemcpyAsync(tb.d_quals, quals + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st));
        CKQ(cudaMemcpyAsync(tb.d_off, off32.data(), 4 * (size_t)(nr + 1), cudaMemcpyHostToDevice, st));
        CKQ(cudaMemcpyAsync(tb.d_lo, lo + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKQ(cudaMemcpyAsync(tb.d_hi, hi + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKQ(cudaMemcpyAsync(tb.d_flags, flags + r0, (size_t)nr, cudaMemcpyHostToDevice, st));
        if (!rc)
            rc = bbduk_b200_qtrim_device(h, cfg, tb.d_bases, quals ? tb.d_quals : nullptr, tb.d_off, nr, paired, tb.d_lo, tb.d_hi,
                                         tb.d_flags, d_stats, st);
        CKQ(cudaMemcpyAsync(lo + r0, tb.d_lo, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKQ(cudaMemcpyAsync(hi + r0, tb.d_hi, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKQ(cudaMemcpyAsync(flags + r0, tb.d_flags, (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKQ(cudaStreamSynchronize(st));
#undef CKQ
        r0 = r1;
    }
    if (!rc && stats8) {
        int64_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpy(v, d_stats, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(h, "qtrim: stats copy failed");
        for (int i = 0; i < 8; i++) stats8[i] += v[i];
    }
    cudaFree(d_stats);
    return rc;
}

void bbduk_b200_entropy_cfg_default(bbduk_entropy_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->cutoff = -1.0f;  // jgi/BBDuk.java:5024
    c->k = 5;
    c->window = 50;
    c->high_pass = 1;
}

int bbduk_b200_entropy_device(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *d_bases, const uint32_t *d_offsets,
                              int64_t n_reads, int32_t paired, const int32_t *d_lo, int32_t *d_hi, uint8_t *d_flags,
                              int64_t *d_stats2, void *stream) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_entropy_cfg");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets || !d_lo || !d_hi || !d_flags) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    const int rc = launch_entropy(h->sm_count, cfg, h->p, d_bases, d_offsets, n_reads, paired ? 1 : 0, d_lo, d_hi, d_flags,
                                  reinterpret_cast<unsigned long long *>(d_stats2), (cudaStream_t)stream);
    if (rc == 2) return set_err(h, "entropy: k > 5 or window - k + 1 > 254 has no device path (and there is no CPU fallback)");
    if (rc) return set_err(h, std::string("entropy kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->launches += 1;
    return 0;
}

int bbduk_b200_entropy_mask_device(bbduk_handle *h, const bbduk_entropy_cfg *cfg, int32_t mode, const uint8_t *d_bases,
                                   const uint32_t *d_offsets, int64_t n_reads, int32_t paired, int32_t *d_lo, int32_t *d_hi,
                                   const uint8_t *d_flags, uint32_t *d_maskbits, const int64_t *d_mask_off, int64_t *d_stats2,
                                   void *stream) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_entropy_cfg");
    if (mode < 1 || mode > 3) return set_err(h, "entropy_mask: mode must be 1 (mask to N), 2 (mask to lower case) or 3 (trim)");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets || !d_lo || !d_hi || !d_flags || !d_maskbits || !d_mask_off) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    const int rc = launch_entropy_mask(h->sm_count, cfg, h->p, d_bases, d_offsets, n_reads, paired ? 1 : 0, d_lo, d_hi,
                                       const_cast<uint8_t *>(d_flags), mode, d_maskbits, d_mask_off,
                                       reinterpret_cast<unsigned long long *>(d_stats2), (cudaStream_t)stream);
    if (rc == 2) return set_err(h, "entropy: k > 5 or window - k + 1 > 254 has no device path (and there is no CPU fallback)");
    if (rc) return set_err(h, std::string("entropy mask kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->launches += 1;
    return 0;
}

static int entropy_host_impl(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                             int32_t paired, int32_t *lo, int32_t *hi, uint8_t *flags, int64_t *stats2, int mode, uint32_t *maskbits,
                             const int64_t *mask_off);

int bbduk_b200_entropy(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                       int32_t paired, const int32_t *lo, int32_t *hi, uint8_t *flags, int64_t *stats2) {
    return entropy_host_impl(h, cfg, bases, offsets, n_reads, paired, const_cast<int32_t *>(lo), hi, flags, stats2, 0, nullptr, nullptr);
}

int bbduk_b200_entropy_mask(bbduk_handle *h, const bbduk_entropy_cfg *cfg, int32_t mode, const uint8_t *bases, const int64_t *offsets,
                            int64_t n_reads, int32_t paired, int32_t *lo, int32_t *hi, const uint8_t *flags, uint32_t *maskbits,
                            const int64_t *mask_off, int64_t *stats2) {
    if (h && (mode < 1 || mode > 3)) return set_err(h, "entropy_mask: mode must be 1 (mask to N), 2 (mask to lower case) or 3 (trim)");
    if (h && n_reads > 0 && (!maskbits || !mask_off)) return set_err(h, "NULL input");
    return entropy_host_impl(h, cfg, bases, offsets, n_reads, paired, lo, hi, const_cast<uint8_t *>(flags), stats2, mode, maskbits, mask_off);
}

static int entropy_host_impl(bbduk_handle *h, const bbduk_entropy_cfg *cfg, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                             int32_t paired, int32_t *lo, int32_t *hi, uint8_t *flags, int64_t *stats2, int mode, uint32_t *maskbits,
                             const int64_t *mask_off) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_entropy_cfg");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (n_reads == 0) return 0;
    if (!bases || !offsets || !lo || !hi || !flags) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> g(h->tbo_mu);  // shares the staging of the tbo entry point
    cudaStream_t st = nullptr;
    int64_t *d_stats = nullptr;
    CKH(cudaMalloc(&d_stats, 2 * sizeof(int64_t)));
    CKH(cudaMemset(d_stats, 0, 2 * sizeof(int64_t)));
    int rc = 0;
    int64_t r0 = 0;
    const int per = paired ? 2 : 1;
    std::vector<uint32_t> off32;
    while (r0 < n_reads && !rc) {
        int64_t r1 = std::min(n_reads, r0 + (CHUNK_READS << 1));
        while (r1 > r0 + per && offsets[r1] - offsets[r0] > CHUNK_BYTES) r1 = r0 + std::max<int64_t>(per, ((r1 - r0) / 2 / per) * per);
        const int64_t nr = r1 - r0, nb = offsets[r1] - offsets[r0];
        if (nb < 0 || nb >= (1ll << 32) - 64) {
            rc = set_err(h, "a read (pair) exceeds 4 GiB (or offsets decrease)");
            break;
        }
        off32.resize(nr + 1);
        for (int64_t i = 0; i <= nr; i++) off32[i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        auto need = [&](void **p, int64_t *cap, int64_t bytes) -> int {
            if (bytes <= *cap) return 0;
            cudaFree(*p);
            *p = nullptr;
            *cap = bytes + bytes / 8 + 4096;
            return cudaMalloc(p, (size_t)*cap) == cudaSuccess ? 0 : 1;
        };
        auto &tb = h->tbo;
        if (need((void **)&tb.d_bases, &tb.cap_bases, nb + 64) || need((void **)&tb.d_off, &tb.cap_off, 4 * (nr + 1)) ||
            need((void **)&tb.d_lo, &tb.cap_lo, 4 * nr) || need((void **)&tb.d_hi, &tb.cap_hi, 4 * nr) ||
            need((void **)&tb.d_flags, &tb.cap_flags, nr)) {
            rc = set_err(h, "entropy: device allocation failed");
            break;
        }
        const int64_t mw0 = mode ? mask_off[r0] : 0, mw = mode ? mask_off[r1] - mw0 : 0;
        if (mode && (mw < 0 || need((void **)&tb.d_mask, &tb.cap_mask, 4 * (mw + 1)) || need((void **)&tb.d_maskoff, &tb.cap_maskoff, 8 * (nr + 1)))) {
            rc = set_err(h, "entropy: mask staging failed");
            break;
        }
#define CKE(call)                                                                       \
    if (!rc && (call) != cudaSuccess) rc = set_err(h, std::string(#call " failed: ") + cudaGetErrorString(cudaGetLastError()))
        CKE(cudaMemcpyAsync(tb.d_bases, bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, st));
        CKE(cudaMemcpyAsync(tb.d_off, off32.data(), 4 * (size_t)(nr + 1), cudaMemcpyHostToDevice, st));
        CKE(cudaMemcpyAsync(tb.d_lo, lo + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKE(cudaMemcpyAsync(tb.d_hi, hi + r0, 4 * (size_t)nr, cudaMemcpyHostToDevice, st));
        CKE(cudaMemcpyAsync(tb.d_flags, flags + r0, (size_t)nr, cudaMemcpyHostToDevice, st));
        if (mode) {
            CKE(cudaMemcpyAsync(tb.d_maskoff, mask_off + r0, 8 * (size_t)(nr + 1), cudaMemcpyHostToDevice, st));
            if (!rc) {
                maskoff_rebase_kernel<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(tb.d_maskoff, mw0, nr + 1);
                h->launches += 1;
            }
            if (!rc)
                rc = bbduk_b200_entropy_mask_device(h, cfg, mode, tb.d_bases, tb.d_off, nr, paired, tb.d_lo, tb.d_hi, tb.d_flags, tb.d_mask,
                                                    tb.d_maskoff, d_stats, st);
            if (mw > 0) CKE(cudaMemcpyAsync(maskbits + mw0, tb.d_mask, 4 * (size_t)mw, cudaMemcpyDeviceToHost, st));
            if (mode == 3) CKE(cudaMemcpyAsync(lo + r0, tb.d_lo, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
            if (mode == 3) CKE(cudaMemcpyAsync(hi + r0, tb.d_hi, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        } else {
        if (!rc) rc = bbduk_b200_entropy_device(h, cfg, tb.d_bases, tb.d_off, nr, paired, tb.d_lo, tb.d_hi, tb.d_flags, d_stats, st);
        CKE(cudaMemcpyAsync(hi + r0, tb.d_hi, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKE(cudaMemcpyAsync(flags + r0, tb.d_flags, (size_t)nr, cudaMemcpyDeviceToHost, st));
        }
        CKE(cudaStreamSynchronize(st));
#undef CKE
        r0 = r1;
    }
    if (!rc && stats2) {
        int64_t v[2] = {0, 0};
        if (cudaMemcpy(v, d_stats, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(h, "entropy: stats copy failed");
        stats2[0] += v[0];
        stats2[1] += v[1];
    }
    cudaFree(d_stats);
    return rc;
}

void bbduk_b200_chain_cfg_default(bbduk_chain_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    bbduk_b200_tbo_cfg_default(&c->tbo);
    bbduk_b200_qtrim_cfg_default(&c->qtrim);
    bbduk_b200_entropy_cfg_default(&c->entropy);
}

int bbduk_b200_process_chain(bbduk_handle *h, const bbduk_chain_cfg *cfg, const uint8_t *bases, const uint8_t *quals,
                             const int64_t *offsets, int64_t n_reads, int32_t paired, const bbduk_out *out, bbduk_stats *stats,
                             int64_t *tbo_stats2, int64_t *qtrim_stats8, int64_t *entropy_stats2) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!cfg || cfg->struct_size != (int32_t)sizeof *cfg) return set_err(h, "bad bbduk_chain_cfg");
    if (!h->finalized) return set_err(h, "process before finalize");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (h->p.mode == MODE_KMASK || h->p.mode == MODE_KSPLIT) return set_err(h, "kmask / ksplit are not chained (they rewrite bases)");
    if (!out || !out->lo || !out->hi || !out->flags) return set_err(h, "the chain needs out->lo, out->hi and out->flags");
    if (out->id0b && h->p.mode == MODE_KTRIM_TIPS) return set_err(h, "the chain does not return out->id0b (ktrim=rl left-tip credit): pass NULL or use bbduk_b200_process");
    if (out->id0b && n_reads > 0) memset(out->id0b, 0xFF, sizeof(int32_t) * (size_t)n_reads);  // no left-tip credit outside ktrim=rl: -1
    if (cfg->do_tbo && !paired) return set_err(h, "tbo needs paired reads");
    const bool need_q = (cfg->do_tbo && quals) || (cfg->do_qtrim && (cfg->qtrim.qtrim_left || cfg->qtrim.qtrim_right ||
                                                                     cfg->qtrim.min_base_quality > 0 || cfg->qtrim.min_avg_quality > 0));
    if (need_q && !quals) return set_err(h, "qtrim / mbq / maq need quality bytes");
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_reads == 0) return 0;
    if (!bases || !offsets) return set_err(h, "NULL input");
    CKH(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> g(h->tbo_mu);
    cudaStream_t st = nullptr;  // synchronous entry point: the legacy default stream
    if (!h->chain_dst) CKH(cudaMalloc(&h->chain_dst, 20 * sizeof(int64_t)));  // kept with the handle: cudaFree would synchronise every call
    int64_t *d_st = h->chain_dst;  // [0..7] k-mer block, [8..9] tbo, [10..17] qtrim, [18..19] entropy
    CKH(cudaMemset(d_st, 0, 20 * sizeof(int64_t)));
    int rc = 0;
    const int per = paired ? 2 : 1;
    // chunks: <= CHUNK_READS / 2 reads and <= CHUNK_BYTES bases each (offsets stay 32-bit on the device); small, so that the
    // host packing of chunk i+1, the upload of chunk i and the kernels of chunk i-1 overlap already in a call of a few chunks
    std::vector<int64_t> cut{0};
    while (cut.back() < n_reads) {
        const int64_t r0 = cut.back();
        int64_t r1 = std::min(n_reads, r0 + (CHUNK_READS >> 1));
        while (r1 > r0 + per && offsets[r1] - offsets[r0] > CHUNK_BYTES) r1 = r0 + std::max<int64_t>(per, ((r1 - r0) / 2 / per) * per);
        const int64_t nb = offsets[r1] - offsets[r0];
        if (nb < 0 || nb >= (1ll << 32) - 64) {
            return set_err(h, "a read (pair) exceeds 4 GiB (or offsets decrease)");
        }
        cut.push_back(r1);
    }
    const int n_chunks = (int)cut.size() - 1;
    auto &tb = h->tbo;
#define CKC(call)                                                                       \
    if (!rc && (call) != cudaSuccess) rc = set_err(h, std::string(#call " failed: ") + cudaGetErrorString(cudaGetLastError()))
    if (!h->chain_copy) {
        CKC(cudaStreamCreateWithFlags(&h->chain_copy, cudaStreamNonBlocking));  // no implicit ordering with the legacy stream
        for (int i = 0; i < 2; i++) {
            CKC(cudaEventCreateWithFlags(&h->chain_in[i], cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&h->chain_free[i], cudaEventDisableTiming));
        }
    }
    auto need = [&](void **p, int64_t *cap, int64_t bytes) -> int {
        if (bytes <= *cap) return 0;
        cudaFree(*p);  // synchronises the device: nothing in flight still uses the old buffer
        *p = nullptr;
        *cap = bytes + bytes / 8 + 4096;
        return cudaMalloc(p, (size_t)*cap) == cudaSuccess ? 0 : 1;
    };
    // Two input slots (bases, qualities, offsets); results live in one set of buffers, which the compute stream
    // downloads before the next chunk's kernels run. The copy stream uploads chunk i+1 while chunk i computes.
    uint8_t **s_bases[2] = {&tb.d_bases, &h->chain_bases2}, **s_quals[2] = {&tb.d_quals, &h->chain_quals2};
    uint32_t **s_off[2] = {&tb.d_off, &h->chain_off2};
    int64_t *c_bases[2] = {&tb.cap_bases, &h->chain_cap_bases2}, *c_quals[2] = {&tb.cap_quals, &h->chain_cap_quals2},
            *c_off[2] = {&tb.cap_off, &h->chain_cap_off2};
    int max_len_of[2] = {0, 0};
    bool packed_of[2] = {false, false};
    // Chunks made of A C G T N only cross PCIe as 2-bit codes + defined bits (0.375 B per base instead of 1), packed by the
    // host workers, and are spelled out again on the device for the steps that read ASCII; the first chunk that holds
    // anything else (lower case, IUPAC, U) ends the packing for the rest of the call.
    bool try_pack = h->pack_host;
    {
        std::lock_guard<std::mutex> pg(h->pool_mu);
        ensure_pool(h);
    }
    auto need_host = [&](void **p, int64_t *cap, int64_t bytes) -> int {
        if (bytes <= *cap) return 0;
        if (*p) cudaFreeHost(*p);
        *p = nullptr;
        *cap = bytes + bytes / 8 + 4096;
        return cudaMallocHost(p, (size_t)*cap) == cudaSuccess ? 0 : 1;
    };
    auto stage = [&](int ci) {  // upload chunk ci into slot ci & 1 on the copy stream
        const int sl = ci & 1;
        const int64_t r0 = cut[ci], nr = cut[ci + 1] - r0, nb = offsets[cut[ci + 1]] - offsets[r0];
        const int64_t groups = (nb + 15) / 16;
        if (need((void **)s_bases[sl], c_bases[sl], nb + 64) || (need_q && need((void **)s_quals[sl], c_quals[sl], nb + 64)) ||
            need((void **)s_off[sl], c_off[sl], 4 * (nr + 1)) || need_host((void **)&h->chain_hoff[sl], &h->chain_cap_hoff[sl], 4 * (nr + 1))) {
            rc = set_err(h, "chain: allocation failed");
            return;
        }
        // every ascii_every-th chunk crosses as ASCII (no CPU work): packing is bound by the host's memory bandwidth, the
        // ASCII chunks by PCIe, and both run at the same time
        bool packed = try_pack && nb >= (1 << 16) && !(h->ascii_every > 0 && n_chunks > 2 && (ci % h->ascii_every) == h->ascii_every - 1);
        if (packed) {
            int64_t cap_h = h->chain_cap_hg[sl], cap_h2 = h->chain_cap_hg[sl], cap_d = h->chain_cap_dg[sl], cap_d2 = h->chain_cap_dg[sl];
            if (need_host((void **)&h->chain_hF[sl], &cap_h, 4 * (groups + 64)) || need_host((void **)&h->chain_hD[sl], &cap_h2, 4 * (groups + 64)) ||
                need((void **)&h->chain_dF[sl], &cap_d, 4 * (groups + 64)) || need((void **)&h->chain_dD[sl], &cap_d2, 4 * (groups + 64))) {
                rc = set_err(h, "chain: allocation failed");
                return;
            }
            h->chain_cap_hg[sl] = std::min(cap_h, cap_h2);
            h->chain_cap_dg[sl] = std::min(cap_d, cap_d2);
        }
        // the slot's host staging is free once chunk ci-2's uploads are done
        if (ci >= 2) cudaEventSynchronize(h->chain_in[sl]);
        std::atomic<int> mx{0};
        std::atomic<bool> plain{true};
        {
            std::lock_guard<std::mutex> pg(h->pool_mu);
            const uint8_t *src = bases + offsets[r0];
            const int64_t *osrc = offsets + r0;
            uint32_t *o32 = h->chain_hoff[sl], *hF = h->chain_hF[sl];
            uint16_t *hD = h->chain_hD[sl];
            h->pool->run([&, src, osrc, o32, hF, hD](int part, int n_parts) {
                const int64_t i0 = (nr + 1) * part / n_parts, i1 = (nr + 1) * (part + 1) / n_parts;
                int m = 0;
                for (int64_t i = i0; i < i1; i++) {
                    o32[i] = (uint32_t)(osrc[i] - osrc[0]);
                    if (i < nr) m = std::max<int64_t>(m, std::min<int64_t>(osrc[i + 1] - osrc[i], 0x7FFFFFFF));
                }
                int cur = mx.load();
                while (m > cur && !mx.compare_exchange_weak(cur, m)) {
                }
                if (packed) {
                    const int64_t gb = (groups + 31) / 32;
                    const int64_t g0 = std::min(groups, gb * part / n_parts * 32), g1 = std::min(groups, gb * (part + 1) / n_parts * 32);
                    if (!pack_bases_range_plain(src, nb, g0, g1, hF, hD)) plain = false;
                }
            });
        }
        if (packed && !plain.load()) {
            packed = false;
            try_pack = false;
        }
        max_len_of[sl] = mx.load();
        packed_of[sl] = packed;
        if (ci >= 2) CKC(cudaStreamWaitEvent(h->chain_copy, h->chain_free[sl], 0));  // chunk ci-2's kernels are done with the slot
        if (packed) {
            CKC((h->h2d_bytes += 4 * groups, cudaMemcpyAsync(h->chain_dF[sl], h->chain_hF[sl], 4 * (size_t)groups, cudaMemcpyHostToDevice, h->chain_copy)));
            CKC((h->h2d_bytes += 2 * groups, cudaMemcpyAsync(h->chain_dD[sl], h->chain_hD[sl], 2 * (size_t)groups, cudaMemcpyHostToDevice, h->chain_copy)));
            if (!rc) {
                unpack_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, h->chain_copy>>>(h->chain_dF[sl], h->chain_dD[sl], *s_bases[sl], groups);
                h->launches += 1;
            }
        } else {
            CKC((h->h2d_bytes += nb, cudaMemcpyAsync(*s_bases[sl], bases + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, h->chain_copy)));
        }
        if (need_q) CKC((h->h2d_bytes += nb, cudaMemcpyAsync(*s_quals[sl], quals + offsets[r0], (size_t)nb, cudaMemcpyHostToDevice, h->chain_copy)));
        CKC((h->h2d_bytes += 4 * (nr + 1), cudaMemcpyAsync(*s_off[sl], h->chain_hoff[sl], 4 * (size_t)(nr + 1), cudaMemcpyHostToDevice, h->chain_copy)));
        CKC(cudaEventRecord(h->chain_in[sl], h->chain_copy));
    };
    if (!rc && n_chunks > 0) stage(0);
    for (int ci = 0; ci < n_chunks && !rc; ci++) {
        const int sl = ci & 1;
        const int64_t r0 = cut[ci], nr = cut[ci + 1] - r0;
        const int max_len = max_len_of[sl];
        if (need((void **)&tb.d_lo, &tb.cap_lo, 4 * nr) || need((void **)&tb.d_hi, &tb.cap_hi, 4 * nr) ||
            need((void **)&tb.d_flags, &tb.cap_flags, nr) || need((void **)&tb.d_insert, &tb.cap_insert, 2 * nr + 8) ||
            (out->id0 && need((void **)&tb.d_id0, &tb.cap_id0, 4 * nr)) || (out->count && need((void **)&tb.d_count, &tb.cap_count, 4 * nr))) {
            rc = set_err(h, "chain: device allocation failed");
            break;
        }
        uint8_t *d_b = *s_bases[sl], *d_q = need_q ? *s_quals[sl] : nullptr;
        uint32_t *d_o = *s_off[sl];
        CKC(cudaStreamWaitEvent(st, h->chain_in[sl], 0));
        bbduk_out dout;
        memset(&dout, 0, sizeof dout);
        dout.lo = tb.d_lo;
        dout.hi = tb.d_hi;
        dout.flags = tb.d_flags;
        dout.id0 = out->id0 ? tb.d_id0 : nullptr;
        dout.count = out->count ? tb.d_count : nullptr;
        if (!rc)
            rc = process_device_impl(h, d_b, d_o, nr, paired, &dout, reinterpret_cast<bbduk_stats *>(d_st), st, std::max(max_len, 1),
                                     packed_of[sl] ? h->chain_dF[sl] : nullptr, packed_of[sl] ? h->chain_dD[sl] : nullptr);
        if (!rc && cfg->do_tbo)
            rc = bbduk_b200_tbo_device(h, &cfg->tbo, d_b, quals ? d_q : nullptr, d_o, nr, max_len, tb.d_lo, tb.d_hi, tb.d_flags,
                                       tb.d_insert, d_st + 8, st);
        if (!rc && cfg->do_qtrim)
            rc = bbduk_b200_qtrim_device(h, &cfg->qtrim, d_b, d_q, d_o, nr, paired, tb.d_lo, tb.d_hi, tb.d_flags, d_st + 10, st);
        if (!rc && cfg->do_entropy)
            rc = bbduk_b200_entropy_device(h, &cfg->entropy, d_b, d_o, nr, paired, tb.d_lo, tb.d_hi, tb.d_flags, d_st + 18, st);
        CKC(cudaEventRecord(h->chain_free[sl], st));
        if (ci + 1 < n_chunks && !rc) stage(ci + 1);  // host-side offsets + upload of the next chunk while this one computes
        CKC(cudaMemcpyAsync(out->lo + r0, tb.d_lo, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKC(cudaMemcpyAsync(out->hi + r0, tb.d_hi, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        CKC(cudaMemcpyAsync(out->flags + r0, tb.d_flags, (size_t)nr, cudaMemcpyDeviceToHost, st));
        if (out->id0) CKC(cudaMemcpyAsync(out->id0 + r0, tb.d_id0, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
        if (out->count) CKC(cudaMemcpyAsync(out->count + r0, tb.d_count, 4 * (size_t)nr, cudaMemcpyDeviceToHost, st));
    }
    cudaStreamSynchronize(h->chain_copy);
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = set_err(h, std::string("chain failed: ") + cudaGetErrorString(cudaGetLastError()));
#undef CKC
    if (!rc) {
        int64_t v[20];
        if (cudaMemcpy(v, d_st, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(h, "chain: stats copy failed");
        if (!rc) {
            if (stats) memcpy(stats, v, 8 * sizeof(int64_t));
            if (tbo_stats2)
                for (int i = 0; i < 2; i++) tbo_stats2[i] += v[8 + i];
            if (qtrim_stats8)
                for (int i = 0; i < 8; i++) qtrim_stats8[i] += v[10 + i];
            if (entropy_stats2)
                for (int i = 0; i < 2; i++) entropy_stats2[i] += v[18 + i];
        }
    }
    return rc;
}

int bbduk_b200_pack_bases(const uint8_t *bases, int64_t n, uint32_t *F, uint16_t *D) {
    if (n < 0 || (n > 0 && (!bases || !F || !D))) return set_err(nullptr, "bad pack_bases arguments");
    pack_bases(bases, n, F, D);
    return 0;
}

int bbduk_b200_transfer_bytes(bbduk_handle *h, int64_t *h2d, int64_t *d2h) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (h2d) *h2d = h->h2d_bytes.load();
    if (d2h) *d2h = h->d2h_bytes.load();
    return 0;
}

int64_t bbduk_b200_launch_count(bbduk_handle *h) { return h ? h->launches.load() : 0; }

void bbduk_b200_destroy(bbduk_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &s : h->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        free_slot(s);
        if (s.done) cudaEventDestroy(s.done);
        if (s.st) cudaStreamDestroy(s.st);
    }
    h->table.release();
    cudaFree(h->d_scaf_reads);
    cudaFree(h->d_scaf_bases);
    cudaFree(h->d_stats);
    cudaFree(h->dev_handoff);
    cudaFree(h->dev_handoff_n);
    cudaFree(h->dev_first64);
    cudaFree(h->dev_lastpos);
    cudaFree(h->dev_sbits);
    cudaFree(h->tbo.d_bases);
    cudaFree(h->tbo.d_quals);
    cudaFree(h->tbo.d_flags);
    cudaFree(h->tbo.d_off);
    cudaFree(h->tbo.d_lo);
    cudaFree(h->tbo.d_hi);
    cudaFree(h->tbo.d_insert);
    cudaFree(h->tbo.d_id0);
    cudaFree(h->tbo.d_count);
    cudaFree(h->chain_bases2);
    cudaFree(h->chain_quals2);
    cudaFree(h->chain_off2);
    cudaFree(h->chain_dst);
    for (int i = 0; i < 2; i++) {
        cudaFree(h->chain_dF[i]);
        cudaFree(h->chain_dD[i]);
        if (h->chain_hF[i]) cudaFreeHost(h->chain_hF[i]);
        if (h->chain_hD[i]) cudaFreeHost(h->chain_hD[i]);
        if (h->chain_hoff[i]) cudaFreeHost(h->chain_hoff[i]);
    }
    if (h->chain_copy) cudaStreamDestroy(h->chain_copy);
    for (int i = 0; i < 2; i++) {
        if (h->chain_in[i]) cudaEventDestroy(h->chain_in[i]);
        if (h->chain_free[i]) cudaEventDestroy(h->chain_free[i]);
    }
    delete h->pool;
    delete h;
}

}  // extern "C"

// ---- single-process multi-GPU: table replication and read sharding inside the library --------------------------------
// The reference drives ONE index from THREADS Java threads of one JVM (bbduk/BBDukS.java:317-319); a JNI caller has no
// torchrun. bbduk_b200_replicate copies a finished table to other GPUs of the box -- one NCCL broadcast per blob over
// NVLink (libnccl is resolved at run time with dlopen, so the library still loads where NCCL is absent), or
// cudaMemcpyPeerAsync where NCCL cannot be used (a target on the source's own GPU, duplicate devices, no libnccl).
namespace {

typedef struct ncclComm *bb_ncclComm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(bb_ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(bb_ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, bb_ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(api.lib, "ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.lib, "ncclBroadcast"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
        api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast;
    });
    return api.ok ? &api : nullptr;
}

struct Blob {
    const void *src;
    size_t bytes;
};

}  // namespace

extern "C" {

int bbduk_b200_replicate(bbduk_handle *src, const int32_t *device_ids, int32_t n_devices, bbduk_handle **out) {
    if (!src) return set_err(nullptr, "handle is NULL");
    if (!src->finalized) return set_err(src, "replicate before finalize");
    if (n_devices < 1 || !device_ids || !out) return set_err(src, "bad replicate arguments");
    for (int i = 0; i < n_devices; i++) out[i] = nullptr;
    bbduk_table_desc d0;
    if (bbduk_b200_table_describe(src, &d0)) return 1;
    std::vector<bbduk_table_desc> dd(n_devices);
    auto fail = [&](const std::string &m) {
        for (int i = 0; i < n_devices; i++) {
            bbduk_b200_destroy(out[i]);
            out[i] = nullptr;
        }
        cudaSetDevice(src->device);
        return set_err(src, m);
    };
    for (int i = 0; i < n_devices; i++) {
        bbduk_cfg c = src->cfg;
        c.device = device_ids[i];
        if (bbduk_b200_create(&c, &out[i])) return fail(std::string("replicate: create on device ") + std::to_string(device_ids[i]) + ": " + g_err);
        dd[i] = d0;
        dd[i].d_keys = dd[i].d_vals = dd[i].d_filter = nullptr;
        if (bbduk_b200_table_alloc(out[i], &dd[i])) return fail(std::string("replicate: table_alloc: ") + bbduk_b200_last_error(out[i]));
    }
    const Blob blobs[3] = {{d0.d_keys, sizeof(uint64_t) * (size_t)d0.n_slots},
                           {d0.d_vals, sizeof(int32_t) * (size_t)d0.n_slots},
                           {d0.d_filter, sizeof(uint32_t) * (size_t)d0.n_filter_words}};
    auto dst_of = [&](int i, int b) -> void * { return b == 0 ? dd[i].d_keys : b == 1 ? dd[i].d_vals : dd[i].d_filter; };
    // targets NCCL can serve: distinct devices, none of them the source's
    std::vector<int> via_nccl, via_peer;
    const char *force = getenv("BBDUK_B200_REPLICATE");  // "peer" | "nccl" (default: nccl where possible)
    NcclApi *api = (force && !strcmp(force, "peer")) ? nullptr : nccl_api();
    for (int i = 0; i < n_devices; i++) {
        bool dup = device_ids[i] == src->device;
        for (int j : via_nccl) dup = dup || device_ids[j] == device_ids[i];
        (api && !dup ? via_nccl : via_peer).push_back(i);
    }
    if (force && !strcmp(force, "nccl") && !via_peer.empty() && !api) return fail("replicate: BBDUK_B200_REPLICATE=nccl but libnccl could not be loaded");
    if (!via_nccl.empty()) {
        const int m = (int)via_nccl.size() + 1;
        std::vector<int> devs(m);
        devs[0] = src->device;
        for (int r = 1; r < m; r++) devs[r] = device_ids[via_nccl[r - 1]];
        std::vector<bb_ncclComm_t> comms(m, nullptr);
        std::vector<cudaStream_t> sts(m, nullptr);
        int nrc = api->CommInitAll(comms.data(), m, devs.data());
        std::string emsg;
        if (nrc != 0) emsg = std::string("ncclCommInitAll: ") + (api->GetErrorString ? api->GetErrorString(nrc) : "error");
        for (int r = 0; r < m && emsg.empty(); r++) {
            if (cudaSetDevice(devs[r]) != cudaSuccess || cudaStreamCreateWithFlags(&sts[r], cudaStreamNonBlocking) != cudaSuccess)
                emsg = "replicate: stream creation failed";
        }
        if (emsg.empty()) {
            api->GroupStart();
            for (int b = 0; b < 3 && nrc == 0; b++) {
                if (blobs[b].bytes == 0) continue;
                for (int r = 0; r < m && nrc == 0; r++) {
                    void *mine = r == 0 ? const_cast<void *>(blobs[b].src) : dst_of(via_nccl[r - 1], b);
                    nrc = api->Broadcast(mine, mine, blobs[b].bytes, 1 /* ncclUint8 */, 0, comms[r], sts[r]);
                }
            }
            const int grc = api->GroupEnd();
            if (nrc == 0) nrc = grc;
            if (nrc != 0) emsg = std::string("ncclBroadcast: ") + (api->GetErrorString ? api->GetErrorString(nrc) : "error");
        }
        for (int r = 0; r < m; r++) {
            if (!sts[r]) continue;
            cudaSetDevice(devs[r]);
            if (cudaStreamSynchronize(sts[r]) != cudaSuccess && emsg.empty()) emsg = std::string("replicate: ") + cudaGetErrorString(cudaGetLastError());
            cudaStreamDestroy(sts[r]);
        }
        for (int r = 0; r < m; r++)
            if (comms[r]) api->CommDestroy(comms[r]);
        if (!emsg.empty()) return fail(emsg);
        for (int i : via_nccl) out[i]->replicated_via = 1;
    }
    if (!via_peer.empty()) {
        if (cudaSetDevice(src->device) != cudaSuccess) return fail("replicate: cudaSetDevice failed");
        cudaStream_t st = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return fail("replicate: stream creation failed");
        cudaError_t e = cudaSuccess;
        for (int i : via_peer) {
            if (device_ids[i] != src->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, src->device, device_ids[i]);
                if (can) {
                    cudaError_t pe = cudaDeviceEnablePeerAccess(device_ids[i], 0);  // already enabled is fine
                    if (pe != cudaSuccess) cudaGetLastError();
                }
            }
            for (int b = 0; b < 3 && e == cudaSuccess; b++)
                if (blobs[b].bytes) e = cudaMemcpyPeerAsync(dst_of(i, b), device_ids[i], blobs[b].src, src->device, blobs[b].bytes, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        if (e != cudaSuccess) return fail(std::string("replicate: peer copy failed: ") + cudaGetErrorString(e));
        for (int i : via_peer) out[i]->replicated_via = 2;
    }
    for (int i = 0; i < n_devices; i++)
        if (bbduk_b200_table_commit(out[i])) return fail(std::string("replicate: table_commit: ") + bbduk_b200_last_error(out[i]));
    cudaSetDevice(src->device);
    return 0;
}

int bbduk_b200_replica_transport(bbduk_handle *h) { return h ? h->replicated_via : -1; }

int64_t bbduk_b200_ref_kmers(bbduk_handle *h) { return (h && h->finalized) ? h->table.ref_kmers : -1; }

int bbduk_b200_table_export(bbduk_handle *h, uint64_t *keys, int32_t *ids, int64_t cap, int64_t *n_out) {
    if (!h) return set_err(nullptr, "handle is NULL");
    if (!h->finalized) return set_err(h, "table_export before finalize");
    if (cap < 0 || (cap > 0 && (!keys || !ids))) return set_err(h, "bad table_export arguments");
    CKH(cudaSetDevice(h->device));
    CKH(cudaDeviceSynchronize());
    int64_t n = 0;
    const int64_t step = 1 << 22;  // slots per staging copy
    std::vector<uint64_t> hk((size_t)std::min<int64_t>(step, h->table.n_slots));
    std::vector<int32_t> hv(hk.size());
    for (int64_t a = 0; a < h->table.n_slots; a += step) {
        const int64_t m = std::min<int64_t>(step, h->table.n_slots - a);
        CKH(cudaMemcpy(hk.data(), h->table.d_keys + a, sizeof(uint64_t) * (size_t)m, cudaMemcpyDeviceToHost));
        CKH(cudaMemcpy(hv.data(), h->table.d_vals + a, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < m; i++) {
            if (hk[i] == BB_EMPTY_KEY) continue;
            if (n < cap) {
                keys[n] = hk[i];
                ids[n] = hv[i];
            }
            n++;
        }
    }
    if (n_out) *n_out = n;
    return 0;
}

int bbduk_b200_process_sharded(bbduk_handle **handles, int32_t n_handles, const uint8_t *bases, const int64_t *offsets,
                               int64_t n_reads, int32_t paired, const bbduk_out *out, bbduk_stats *stats) {
    if (!handles || n_handles < 1 || !handles[0]) return set_err(nullptr, "bad process_sharded arguments");
    bbduk_handle *h = handles[0];
    for (int i = 0; i < n_handles; i++)
        if (!handles[i]) return set_err(h, "process_sharded: NULL handle");
    if (n_reads < 0 || (paired && (n_reads & 1))) return set_err(h, "bad n_reads (paired input needs an even count)");
    if (!out) return set_err(h, "out is NULL");
    if (stats) memset(stats, 0, sizeof *stats);
    if (n_reads == 0) return 0;
    const int per = paired ? 2 : 1;
    const int64_t units = n_reads / per;
    std::vector<int> rcs(n_handles, 0);
    std::vector<bbduk_stats> sts(n_handles);
    std::vector<std::thread> th;
    for (int i = 0; i < n_handles; i++) {
        const int64_t r0 = units * i / n_handles * per, r1 = units * (i + 1) / n_handles * per;
        th.emplace_back([=, &rcs, &sts] {
            memset(&sts[i], 0, sizeof(bbduk_stats));
            if (r1 <= r0) return;
            bbduk_out o = *out;  // slices of the caller's arrays; kmask words are addressed through absolute mask_off entries
            if (o.id0) o.id0 += r0;
            if (o.id0b) o.id0b += r0;
            if (o.lo) o.lo += r0;
            if (o.hi) o.hi += r0;
            if (o.flags) o.flags += r0;
            if (o.count) o.count += r0;
            if (o.mask_off) o.mask_off += r0;
            rcs[i] = bbduk_b200_process(handles[i], bases, offsets + r0, r1 - r0, paired, &o, &sts[i]);
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < n_handles; i++)
        if (rcs[i]) return set_err(h, std::string("process_sharded: shard ") + std::to_string(i) + ": " + bbduk_b200_last_error(handles[i]));
    if (stats) {
        int64_t *acc = reinterpret_cast<int64_t *>(stats);
        for (int i = 0; i < n_handles; i++) {
            const int64_t *v = reinterpret_cast<const int64_t *>(&sts[i]);
            for (size_t q = 0; q < sizeof(bbduk_stats) / sizeof(int64_t); q++) acc[q] += v[q];
        }
    }
    return 0;
}

int bbduk_b200_scaffold_counts_sum(bbduk_handle **handles, int32_t n_handles, int64_t *read_counts, int64_t *base_counts, int32_t n) {
    if (!handles || n_handles < 1 || !handles[0]) return set_err(nullptr, "bad scaffold_counts_sum arguments");
    std::vector<int64_t> r(std::max(n, 0)), b(std::max(n, 0));
    if (read_counts) std::fill(read_counts, read_counts + std::max(n, 0), 0);
    if (base_counts) std::fill(base_counts, base_counts + std::max(n, 0), 0);
    for (int i = 0; i < n_handles; i++) {
        std::fill(r.begin(), r.end(), 0);
        std::fill(b.begin(), b.end(), 0);
        if (bbduk_b200_scaffold_counts(handles[i], r.data(), b.data(), n)) return set_err(handles[0], bbduk_b200_last_error(handles[i]));
        for (int j = 0; j < n; j++) {
            if (read_counts) read_counts[j] += r[j];
            if (base_counts) base_counts[j] += b[j];
        }
    }
    return 0;
}

}  // extern "C"
