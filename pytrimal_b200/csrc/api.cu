// api.cu -- the C ABI of libtrimal_cuda.so (include/trimal_cuda.h): handle
// management, host<->device staging, and the call sequences around the kernels.
// No CPU implementation of any statistic lives here: without a device every
// compute entry point fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "nccl_dyn.h"
#include "tcu_internal.cuh"

using namespace tcu;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

static int cuda_fail(cudaError_t e, const char *what)
{
    // clear the sticky-free error state so the next call starts clean
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation)
        return fail(TCU_ERR_OOM, "%s: out of device memory", what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return fail(TCU_ERR_NO_DEVICE, "%s: %s", what, cudaGetErrorString(e));
    return fail(TCU_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);        \
    } while (0)

extern "C" const char *tcu_last_error(void) { return g_last_error.c_str(); }

// NVTX ranges around the entry points (the reference's StartTiming scopes):
// libnvToolsExt is looked up at run time, a profiler that wants the ranges preloads it;
// without it the calls cost one branch.
namespace {
struct NvtxApi {
    int (*push)(const char *) = nullptr;
    int (*pop)() = nullptr;
    NvtxApi()
    {
        if (getenv("TRIMAL_CUDA_NVTX") == nullptr) return;
        void *h = dlopen("libnvToolsExt.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnvToolsExt.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        push = (int (*)(const char *))dlsym(h, "nvtxRangePushA");
        pop = (int (*)())dlsym(h, "nvtxRangePop");
        if (!push || !pop) push = nullptr, pop = nullptr;
    }
};
const NvtxApi &nvtx_api()
{
    static NvtxApi api;
    return api;
}
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char *name) : on(nvtx_api().push != nullptr)
    {
        if (on) nvtx_api().push(name);
    }
    ~NvtxRange()
    {
        if (on) nvtx_api().pop();
    }
};
}  // namespace
extern "C" const char *tcu_version(void) { return "trimal_cuda 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------
// devices
// ---------------------------------------------------------------------------
// cudaGetDeviceProperties queries the whole device (it took 9-300 ms per call inside a
// busy process): two attributes are all that is needed, and they are looked up once.
static bool device_usable(int dev, int *sms)
{
    struct Info {
        int state = 0;  // 0 unknown, 1 usable, 2 not
        int sms = 0;
    };
    static Info info[64];
    static std::mutex mu;
    if (dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lk(mu);
    Info &i = info[dev];
    if (i.state == 0) {
        int major = 0, n = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            cudaGetLastError();
            return false;  // not cached: the driver may come up later
        }
        i.sms = n;
        i.state = major == 10 ? 1 : 2;  // the library carries sm_100a code only
    }
    if (sms) *sms = i.sms;
    return i.state == 1;
}

extern "C" int tcu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < n; d++) usable += device_usable(d, nullptr);
    return usable;
}

// ---------------------------------------------------------------------------
// pinned staging pool (process-wide; buffers are reused across handles because
// cudaHostAlloc costs milliseconds)
// ---------------------------------------------------------------------------
namespace {
constexpr size_t STAGE_BYTES = 16u << 20;
std::mutex g_pool_mutex;
std::vector<void *> g_pool;

void *stage_acquire()
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        if (!g_pool.empty()) {
            void *p = g_pool.back();
            g_pool.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void stage_release(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    g_pool.push_back(p);
}
}  // namespace

// ---------------------------------------------------------------------------
// device buffer cache.  cudaMalloc / cudaFree synchronise the device and cost
// milliseconds for the large buffers (the packed identity array is 5 GB at 50 000 rows,
// the threshold bit matrix 313 MB), sometimes hundreds of milliseconds inside a process
// that holds other big allocations.  Every device buffer of a handle therefore comes
// from, and on tcu_msa_destroy goes back to, a small per-device pool of freed buffers
// (at most POOL_SLOTS of them; tcu_release_cached_memory() returns them to the driver),
// so that a create / compute / destroy cycle on same-sized alignments allocates nothing.
// ---------------------------------------------------------------------------
namespace {
struct DevSlot {
    void *ptr = nullptr;
    size_t cap = 0;
};
constexpr int POOL_SLOTS = 24;
std::mutex g_dev_cache_mutex;
DevSlot g_dev_cache[64][POOL_SLOTS];

// best fit: the smallest cached buffer that holds `need` without wasting more than
// half of itself (plus 1 MB of slack for the small ones)
void *dev_cache_take(int device, size_t need, size_t *cap)
{
    std::lock_guard<std::mutex> lk(g_dev_cache_mutex);
    DevSlot *best = nullptr;
    for (DevSlot &s : g_dev_cache[device & 63])
        if (s.ptr && s.cap >= need && s.cap <= 2 * need + (1u << 20) && (!best || s.cap < best->cap))
            best = &s;
    if (!best) return nullptr;
    void *p = best->ptr;
    *cap = best->cap;
    *best = DevSlot{};
    return p;
}
void dev_cache_give(int device, void *ptr, size_t cap)
{
    if (!ptr) return;
    void *drop = ptr;
    {
        std::lock_guard<std::mutex> lk(g_dev_cache_mutex);
        DevSlot *slot = nullptr, *smallest = nullptr;
        for (DevSlot &s : g_dev_cache[device & 63]) {
            if (!s.ptr) {
                slot = &s;
                break;
            }
            if (!smallest || s.cap < smallest->cap) smallest = &s;
        }
        if (slot) {
            *slot = DevSlot{ptr, cap};
            drop = nullptr;
        } else if (smallest && smallest->cap < cap) {  // full: keep the larger buffers
            drop = smallest->ptr;
            *smallest = DevSlot{ptr, cap};
        }
    }
    if (drop) cudaFree(drop);
}
}  // namespace

// Streams and events of destroyed handles, per device: creating and destroying two streams and
// eight events costs 0.2-0.3 ms per handle life cycle, which is a tenth of an 8-GPU step.
namespace {
struct StreamSet {
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t cev[2] = {nullptr, nullptr};
};
std::mutex g_streams_mutex;
std::vector<StreamSet> g_streams[64];

bool streams_take(int device, StreamSet &out)
{
    std::lock_guard<std::mutex> lk(g_streams_mutex);
    auto &v = g_streams[device & 63];
    if (v.empty()) return false;
    out = v.back();
    v.pop_back();
    return true;
}
void streams_destroy(StreamSet &s)
{
    for (auto &e : s.ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : s.cev)
        if (e) cudaEventDestroy(e);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
    s = StreamSet{};
}
void streams_give(int device, StreamSet &s)  // the streams must be idle
{
    bool complete = s.stream && s.copy_stream;
    for (auto &e : s.ev) complete = complete && e;
    for (auto &e : s.cev) complete = complete && e;
    if (complete) {
        std::lock_guard<std::mutex> lk(g_streams_mutex);
        auto &v = g_streams[device & 63];
        if (v.size() < 16) {
            v.push_back(s);
            s = StreamSet{};
            return;
        }
    }
    streams_destroy(s);
}
}  // namespace

extern "C" void tcu_release_cached_memory(void)
{
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; d++) {
        for (int k = 0; k < POOL_SLOTS; k++) {
            void *p = nullptr;
            {
                std::lock_guard<std::mutex> lk(g_dev_cache_mutex);
                p = g_dev_cache[d][k].ptr;
                g_dev_cache[d][k] = DevSlot{};
            }
            if (p) {
                cudaSetDevice(d);
                cudaFree(p);
            }
        }
    }
    for (int d = 0; d < 64; d++) {
        std::vector<StreamSet> v;
        {
            std::lock_guard<std::mutex> lk(g_streams_mutex);
            v.swap(g_streams[d]);
        }
        if (!v.empty()) cudaSetDevice(d);
        for (StreamSet &s : v) streams_destroy(s);
    }
    cudaSetDevice(cur);
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        for (void *p : g_pool) cudaFreeHost(p);
        g_pool.clear();
    }
    cudaGetLastError();
}

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct tcu_msa {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int nseq = 0, ncol = 0;
    size_t pitch = 0;
    uint8_t *d_raw = nullptr;
    size_t raw_cap = 0;

    // byte presence (mask independent)
    bool have_present = false;
    unsigned int present[256];

    // identity operand
    bool prepared = false;
    int np = 0, nk = 0, nb = 0, nchunks = 0;
    int nsb = 0, nb2 = 0;     // v2: 128-row super-blocks, allocated 64-row blocks (even)
    uint32_t *d_planes = nullptr;
    size_t planes_cap = 0;
    uint8_t *d_gbytes = nullptr;
    size_t gbytes_cap = 0;
    int *d_kept_rows = nullptr;    // the three small tables of the operand live in one pooled
    uint8_t *d_col_drop = nullptr; // buffer (d_small): cudaMalloc / cudaFree per handle would
    uint8_t *d_lut = nullptr;      // synchronise the whole device
    void *d_small = nullptr;
    size_t small_cap = 0;
    uint8_t prepared_indet = 0;

    // identities kept on the device
    float *d_ident = nullptr;
    size_t ident_cap = 0;
    bool ident_full = false;  // unmasked rows: usable by tcu_similarity

    // threshold bit matrix (slab layout, tcu_internal.cuh) of the clustering calls
    uint32_t *d_bits = nullptr;   // threshold bits as K1 leaves them (slab layout, column words)
    size_t bits_cap = 0;
    uint32_t *d_brows = nullptr;  // full symmetric bit matrix, one row per sequence
    size_t brows_cap = 0;

    // generic scratch
    void *d_scratch = nullptr;
    size_t scratch_cap = 0;

    // several GPUs in one process (tcu_set_devices / TRIMAL_CUDA_DEVICES): the handle the
    // caller holds lives on the first device; one replica per further device hangs off it
    std::vector<tcu_msa *> peers;
    bool peer_direct = false;  // kernels of the peers' devices may store into this device's memory

    cudaStream_t copy_stream = nullptr;       // D2H of finished sub-bands, overlapping the kernel
    std::vector<cudaEvent_t> band_done;       // one per sub-band (no timing)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t cev[2] = {nullptr, nullptr};  // first / last D2H on copy_stream
    tcu_timings timings{};
    bool pending_device_timing = false;  // ev[2]..ev[3] bracket an async tcu_identity_device
};

// grow-only device buffer, pooled (see the device buffer cache above)
static int ensure_dev(int device, void **p, size_t *cap, size_t need, tcu_msa *owner = nullptr)
{
    if (*cap >= need && *p) return TCU_OK;
    if (*p) {
        // growth: queued work may still read the old buffer.  Only this handle's own streams
        // can hold such work (a buffer belongs to one handle), so only they are waited for --
        // not the whole device, which would stall every other thread's handle.
        if (owner) {
            if (owner->stream) cudaStreamSynchronize(owner->stream);
            if (owner->copy_stream) cudaStreamSynchronize(owner->copy_stream);
        } else {
            cudaDeviceSynchronize();
        }
    }
    dev_cache_give(device, *p, *cap);
    *p = nullptr;
    *cap = 0;
    need = (std::max<size_t>(need, 1) + (1u << 20) - 1) >> 20 << 20;  // 1 MB granules: poolable
    size_t got = 0;
    if (void *q = dev_cache_take(device, need, &got)) {
        *p = q;
        *cap = got;
        return TCU_OK;
    }
    CK(cudaMalloc(p, need));
    *cap = need;
    return TCU_OK;
}

static int ensure_ident(tcu_msa *m, size_t need)
{
    return ensure_dev(m->device, (void **)&m->d_ident, &m->ident_cap, need, m);
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

// ---------------------------------------------------------------------------
// several GPUs, one process per GPU: a communicator is an NCCL communicator
// bound to this rank's device.  Used by the *_all entry points only.
// ---------------------------------------------------------------------------
struct tcu_comm {
    NcclComm comm = nullptr;
    int rank = 0, world = 1, device = 0;
    // peer memory over CUDA IPC: device allocations of the other ranks mapped into this
    // process (per rank: handle bytes -> mapped base), so that bands are pulled device to
    // device by the copy engines over NVLink instead of going through NCCL's transfer kernels
    std::vector<std::map<std::string, void *>> opened;
    // a communicator is used by one call at a time (NCCL requires the ranks to issue the same
    // operations in the same order; two threads of one rank interleaving theirs would break that)
    std::mutex mutex;
    uint8_t *d_sync = nullptr;   // handle exchange area + the word of the stream-ordered barriers
    uint8_t *h_sync = nullptr;   // pinned mirror
    // peer memory between the ranks; cleared for good on the first failure (agreed by all
    // ranks), or from the start with TRIMAL_CUDA_NO_PEER=1 in the environment (every rank's)
    bool ipc_usable = getenv("TRIMAL_CUDA_NO_PEER") == nullptr;
};
constexpr size_t COMM_SYNC_BYTES = 64 * 64 + 256;  // up to 64 ranks x 64-byte IPC handles + flags

static int nccl_fail(int r, const char *what)
{
    const NcclApi &api = nccl_api();
    return fail(TCU_ERR_NCCL, "%s: %s", what,
                api.ok && api.GetErrorString ? api.GetErrorString(r) : api.why);
}

#define NK(call)                                                \
    do {                                                        \
        int r__ = (call);                                       \
        if (r__ != NCCL_SUCCESS) return nccl_fail(r__, #call);  \
    } while (0)

static int comm_check(const tcu_msa *m, const tcu_comm *c)
{
    if (c && c->device != m->device)
        return fail(TCU_ERR_INVALID, "communicator is bound to device %d, the alignment to %d",
                    c->device, m->device);
    return TCU_OK;
}

// Share of rank `rank` when `total` units are cut into `world` contiguous ranges whose
// boundaries are multiples of `granule`: [begin, end), empty for surplus ranks.
extern "C" int tcu_shard_range(int total, int granule, int rank, int world, int *begin, int *end)
{
    if (total < 0 || granule < 1 || world < 1 || rank < 0 || rank >= world || !begin || !end)
        return fail(TCU_ERR_INVALID, "bad shard arguments");
    const long long units = ((long long)total + granule - 1) / granule;
    const long long lo = units * rank / world, hi = units * (rank + 1) / world;
    *begin = (int)std::min<long long>(lo * granule, total);
    *end = (int)std::min<long long>(hi * granule, total);
    return TCU_OK;
}

// Row-blocks [begin, end) of the pair matrix for rank `rank`: contiguous bands holding
// (nearly) the same number of 128x64 tiles, i.e. the same work.
extern "C" int tcu_shard_blocks(int kept_rows, int rank, int world, int *block_begin,
                                int *block_end)
{
    if (kept_rows < 0 || world < 1 || rank < 0 || rank >= world || !block_begin || !block_end)
        return fail(TCU_ERR_INVALID, "bad shard arguments");
    const int nb = (kept_rows + RB - 1) / RB, nsb = (kept_rows + IB - 1) / IB;
    const long long total = tiles_before2(nsb, nb);
    auto bound = [&](int g) {
        if (g <= 0) return 0;
        if (g >= world) return nsb;
        // first block whose preceding work reaches g/world of the total
        int lo = 0, hi = nsb;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (tiles_before2(mid, nb) * world < total * g) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    *block_begin = bound(rank);
    *block_end = bound(rank + 1);
    return TCU_OK;
}

extern "C" int tcu_comm_id(void *id)
{
    if (!id) return fail(TCU_ERR_INVALID, "id is NULL");
    const NcclApi &api = nccl_api();
    if (!api.ok) return fail(TCU_ERR_NCCL, "NCCL unavailable: %s", api.why);
    NcclUniqueId u;
    NK(api.GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return TCU_OK;
}

extern "C" int tcu_comm_create(const void *id, int rank, int world, int device, tcu_comm **out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world)
        return fail(TCU_ERR_INVALID, "bad communicator arguments");
    *out = nullptr;
    const NcclApi &api = nccl_api();
    if (!api.ok) return fail(TCU_ERR_NCCL, "NCCL unavailable: %s", api.why);
    if (!device_usable(device, nullptr))
        return fail(TCU_ERR_NO_DEVICE, "device %d is not an sm_100 GPU", device);
    CK(cudaSetDevice(device));
    tcu_comm *c = new (std::nothrow) tcu_comm();
    if (!c) return fail(TCU_ERR_OOM, "host allocation failed");
    c->rank = rank;
    c->world = world;
    c->device = device;
    NcclUniqueId u;
    memcpy(&u, id, sizeof u);
    int r = api.CommInitRank(&c->comm, world, u, rank);
    if (r != NCCL_SUCCESS) {
        delete c;
        return nccl_fail(r, "ncclCommInitRank");
    }
    c->opened.resize((size_t)world);
    if (world > 64 || cudaMalloc((void **)&c->d_sync, COMM_SYNC_BYTES) != cudaSuccess ||
        cudaHostAlloc((void **)&c->h_sync, COMM_SYNC_BYTES, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        c->ipc_usable = false;  // the NCCL path needs neither
    } else {
        cudaMemset(c->d_sync, 0, COMM_SYNC_BYTES);
    }
    *out = c;
    return TCU_OK;
}

extern "C" void tcu_comm_destroy(tcu_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &per_rank : c->opened)
        for (auto &kv : per_rank) cudaIpcCloseMemHandle(kv.second);
    if (c->d_sync) cudaFree(c->d_sync);
    if (c->h_sync) cudaFreeHost(c->h_sync);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    cudaGetLastError();
    delete c;
}

extern "C" int tcu_comm_rank(const tcu_comm *c) { return c ? c->rank : -1; }
extern "C" int tcu_comm_world(const tcu_comm *c) { return c ? c->world : 0; }

// In-place all-gather of unequal contiguous byte ranges of `buf` (rank r owns
// [off[r], off[r]+cnt[r])): every rank sends its range to every other rank and receives
// theirs, all point-to-point transfers fused into one NCCL group so that they run
// concurrently over NVLink / NVSwitch (a group of broadcasts, the first version, moved
// 78 MB in 0.67 ms; NVSwitch gives every pair its own full-bandwidth path).
static int comm_allgatherv(tcu_comm *c, void *buf, const size_t *off, const size_t *cnt,
                           cudaStream_t stream)
{
    const NcclApi &api = nccl_api();
    const int me = c->rank;
    NK(api.GroupStart());
    for (int d = 1; d < c->world; d++) {
        // staggered partners: at step d everybody sends to rank + d and receives from rank - d
        const int to = (me + d) % c->world, from = (me - d + c->world) % c->world;
        int rc = NCCL_SUCCESS;
        if (cnt[me]) rc = api.Send((const uint8_t *)buf + off[me], cnt[me], NCCL_UINT8, to, c->comm, stream);
        if (rc == NCCL_SUCCESS && cnt[from])
            rc = api.Recv((uint8_t *)buf + off[from], cnt[from], NCCL_UINT8, from, c->comm, stream);
        if (rc != NCCL_SUCCESS) {
            api.GroupEnd();
            return nccl_fail(rc, "ncclSend/ncclRecv");
        }
    }
    NK(api.GroupEnd());
    return TCU_OK;
}

// ---------------------------------------------------------------------------
// Peer memory between ranks.  `buf` is the base of a cudaMalloc allocation of the same size
// on every rank.  peer_prepare (before the producing kernel is launched): the ranks exchange
// CUDA IPC handles of `buf` and map the ones they have not seen yet (allocations come from
// the library's pool, so after the first call nothing is mapped any more); all ranks agree on
// whether that worked.  The producing kernel then stores its results into every rank's copy
// (NVLink / NVSwitch), and a one-word all-reduce behind it tells everybody that all parts
// have landed.  (Round 2 first PULLED the finished bands with 2-D device-to-device copies
// between two such barriers: 0.9 ms at 8 GPUs that the stores inside K1 do not cost.)
// ---------------------------------------------------------------------------
static int comm_allreduce_i32(tcu_comm *c, int *buf, size_t count, cudaStream_t stream);
static int comm_allgatherv(tcu_comm *c, void *buf, const size_t *off, const size_t *cnt,
                           cudaStream_t stream);

static bool peer_prepare(tcu_comm *c, void *buf, std::vector<void *> &peer_base, cudaStream_t stream)
{
    if (!c->ipc_usable || !c->d_sync) return false;
    peer_base.assign((size_t)c->world, nullptr);
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (cudaIpcGetMemHandle(&mine, buf) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // my handle into my slot, zeros elsewhere; an integer sum over all ranks is then the
    // all-gather (one small all-reduce: a fraction of the latency of 2 (world - 1)
    // point-to-point transfers)
    memset(c->h_sync, 0, 64 * (size_t)c->world);
    memcpy(c->h_sync + 64 * (size_t)c->rank, &mine, 64);
    bool comm_ok = cudaMemcpyAsync(c->d_sync, c->h_sync, 64 * (size_t)c->world, cudaMemcpyHostToDevice,
                                   stream) == cudaSuccess &&
                   comm_allreduce_i32(c, (int *)c->d_sync, 16 * (size_t)c->world, stream) == TCU_OK &&
                   cudaMemcpyAsync(c->h_sync, c->d_sync, 64 * (size_t)c->world, cudaMemcpyDeviceToHost,
                                   stream) == cudaSuccess &&
                   cudaStreamSynchronize(stream) == cudaSuccess;
    if (!comm_ok) {
        cudaGetLastError();
        c->ipc_usable = false;  // a failed collective: nothing sensible can be agreed any more
        return false;
    }
    for (int r = 0; r < c->world && ok; r++) {
        if (r == c->rank) {
            peer_base[r] = buf;
            continue;
        }
        const std::string key((const char *)c->h_sync + 64 * (size_t)r, 64);
        auto it = c->opened[r].find(key);
        if (it == c->opened[r].end()) {
            cudaIpcMemHandle_t h;
            memcpy(&h, key.data(), 64);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            it = c->opened[r].emplace(key, p).first;
        }
        peer_base[r] = it->second;
    }
    // agree: the sum of the ok flags must be `world`
    int *flag = (int *)(c->h_sync + 64 * 64);
    *flag = ok;
    int *d_flag = (int *)(c->d_sync + 64 * 64);
    comm_ok = cudaMemcpyAsync(d_flag, flag, sizeof(int), cudaMemcpyHostToDevice, stream) == cudaSuccess &&
              comm_allreduce_i32(c, d_flag, 1, stream) == TCU_OK &&
              cudaMemcpyAsync(flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
              cudaStreamSynchronize(stream) == cudaSuccess;
    if (!comm_ok || *flag != c->world) {
        cudaGetLastError();
        c->ipc_usable = false;
        if (getenv("TCU_TRACE")) fprintf(stderr, "[tcu] rank %d: peer memory unavailable (ok flags %d of %d), using NCCL transfers\n", c->rank, *flag, c->world);
        return false;
    }
    return true;
}

static int comm_allreduce_i32(tcu_comm *c, int *buf, size_t count, cudaStream_t stream)
{
    const NcclApi &api = nccl_api();
    NK(api.AllReduce(buf, buf, count, NCCL_INT32, NCCL_SUM, c->comm, stream));
    return TCU_OK;
}

static int msa_alloc(int nseq, int ncol, int device, tcu_msa **out)
{
    if (!out) return fail(TCU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nseq < 0 || ncol < 0) return fail(TCU_ERR_INVALID, "negative alignment shape");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TCU_ERR_NO_DEVICE, "no CUDA device available (no CPU fallback exists)");
    }
    if (device < 0 || device >= ndev) return fail(TCU_ERR_INVALID, "device %d out of range", device);
    int sms = 0;
    if (!device_usable(device, &sms))
        return fail(TCU_ERR_NO_DEVICE, "device %d is not an sm_100 GPU", device);
    CK(cudaSetDevice(device));
    tcu_msa *m = new (std::nothrow) tcu_msa();
    if (!m) return fail(TCU_ERR_OOM, "host allocation failed");
    m->device = device;
    m->num_sms = sms;
    m->nseq = nseq;
    m->ncol = ncol;
    m->pitch = ((size_t)ncol + 127) / 128 * 128;
    if (m->pitch == 0) m->pitch = 128;
    cudaError_t e = cudaSuccess;
    StreamSet pooled;
    if (streams_take(device, pooled)) {
        m->stream = pooled.stream;
        m->copy_stream = pooled.copy_stream;
        for (int i = 0; i < 6; i++) m->ev[i] = pooled.ev[i];
        for (int i = 0; i < 2; i++) m->cev[i] = pooled.cev[i];
    } else {
        e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
        for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&m->ev[i]);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&m->cev[i]);
    }
    if (e != cudaSuccess) {
        tcu_msa_destroy(m);
        return cuda_fail(e, "tcu_msa_create");
    }
    int rc = ensure_dev(device, (void **)&m->d_raw, &m->raw_cap,
                        std::max<size_t>(1, (size_t)nseq) * m->pitch);
    if (rc != TCU_OK) {
        tcu_msa_destroy(m);
        return rc;
    }
    *out = m;
    return TCU_OK;
}

static size_t gather_threads()
{
    static const size_t n = [] {
        unsigned hc = std::thread::hardware_concurrency();
        return (size_t)std::max(1u, std::min(8u, hc ? hc / 2 : 1u));
    }();
    return n;
}

// One strided host buffer that is page-locked (cudaHostAlloc / cudaHostRegister): the DMA
// engine reads it in place, no staging copy.
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Page-locked source: ONE linear DMA of the whole strided block into scratch (a 2-D copy
// would issue a descriptor per 1000-byte row), then a device kernel lays the rows out at
// the device pitch and zero-fills the padding.
static int upload_strided_pinned(tcu_msa *m, const uint8_t *data, size_t stride, int r0, int r1,
                                 const std::vector<uint8_t *> *peer_raw = nullptr)
{
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    if (r1 <= r0) return TCU_OK;
    const size_t bytes = (size_t)(r1 - r0 - 1) * stride + (size_t)m->ncol;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, bytes, m);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaMemcpyAsync(m->d_scratch, data + (size_t)r0 * stride, bytes, cudaMemcpyHostToDevice,
                       m->stream));
    CK(launch_repitch_rows((const uint8_t *)m->d_scratch, stride, r1 - r0, m->ncol, m->d_raw, m->pitch,
                           (size_t)r0 * m->pitch, peer_raw ? peer_raw->data() : nullptr,
                           peer_raw ? (int)peer_raw->size() : 0, m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

// rows -> pinned staging (row pitch = device pitch, padding zero) -> device
template <typename RowPtr>
static int upload_rows(tcu_msa *m, RowPtr row_of, int row_begin, int row_end)
{
    CK(cudaSetDevice(m->device));
    CK(cudaEventRecord(m->ev[0], m->stream));
    const size_t pitch = m->pitch;
    const size_t rows_per_stage = std::max<size_t>(1, STAGE_BYTES / pitch);
    if (row_end <= row_begin) {
    } else if (pitch > STAGE_BYTES) {
        // very long rows: copy row by row straight from caller memory
        CK(cudaMemsetAsync(m->d_raw + (size_t)row_begin * pitch, 0,
                           (size_t)(row_end - row_begin) * pitch, m->stream));
        for (int r = row_begin; r < row_end; r++)
            CK(cudaMemcpyAsync(m->d_raw + (size_t)r * pitch, row_of(r), m->ncol,
                               cudaMemcpyHostToDevice, m->stream));
    } else {
        void *stage[2] = {stage_acquire(), stage_acquire()};
        cudaEvent_t done[2] = {m->ev[4], m->ev[5]};
        bool used[2] = {false, false};
        if (!stage[0] || !stage[1]) {
            stage_release(stage[0]);
            stage_release(stage[1]);
            return fail(TCU_ERR_OOM, "pinned staging allocation failed");
        }
        int which = 0;
        int rc = TCU_OK;
        for (size_t r0 = (size_t)row_begin; r0 < (size_t)row_end && rc == TCU_OK;
             r0 += rows_per_stage) {
            const size_t nr = std::min(rows_per_stage, (size_t)row_end - r0);
            uint8_t *s = (uint8_t *)stage[which];
            if (used[which] && cudaEventSynchronize(done[which]) != cudaSuccess) {
                rc = cuda_fail(cudaGetLastError(), "staging wait");
                break;
            }
            // gather the rows (separate heap strings in trimAl) into the stage; split over
            // a few threads when the stage is large -- one core copies ~5 GB/s, PCIe takes 50
            auto gather = [&](size_t a, size_t b) {
                for (size_t r = a; r < b; r++) {
                    memcpy(s + r * pitch, row_of((int)(r0 + r)), m->ncol);
                    if (pitch > (size_t)m->ncol)
                        memset(s + r * pitch + m->ncol, 0, pitch - m->ncol);
                }
            };
            const size_t workers = nr * pitch >= (4u << 20) ? gather_threads() : 1;
            if (workers <= 1) {
                gather(0, nr);
            } else {
                std::vector<std::thread> pool;
                size_t started = 1;  // share 0 is this thread's
                try {
                    for (; started < workers; started++)
                        pool.emplace_back(gather, nr * started / workers,
                                          nr * (started + 1) / workers);
                } catch (...) {  // no more threads to be had: the rest is done here
                }
                gather(0, nr / workers);
                if (started < workers) gather(nr * started / workers, nr);
                for (auto &t : pool) t.join();
            }
            cudaError_t e = cudaMemcpyAsync(m->d_raw + r0 * pitch, s, nr * pitch,
                                            cudaMemcpyHostToDevice, m->stream);
            if (e == cudaSuccess) e = cudaEventRecord(done[which], m->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "row upload");
            used[which] = true;
            which ^= 1;
        }
        cudaStreamSynchronize(m->stream);
        stage_release(stage[0]);
        stage_release(stage[1]);
        if (rc != TCU_OK) return rc;
    }
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings = tcu_timings{};
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    return TCU_OK;
}

static double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// ---------------------------------------------------------------------------
// Device set.  The reference has one compute platform per process and no notion of a
// device; a caller that wants several GPUs behind the same single-process API (pytrimal's
// platform="cuda") names them once -- tcu_set_devices() or the environment variable
// TRIMAL_CUDA_DEVICES ("all" or a comma separated list) -- and creates its handles with
// device = TCU_DEVICE_AUTO.
// ---------------------------------------------------------------------------
namespace {
std::mutex g_devset_mutex;
std::vector<int> g_devset;
bool g_devset_ready = false;

std::vector<int> usable_devices()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    std::vector<int> v;
    for (int d = 0; d < n; d++)
        if (device_usable(d, nullptr)) v.push_back(d);
    return v;
}

std::vector<int> configured_devices()
{
    std::lock_guard<std::mutex> lk(g_devset_mutex);
    if (!g_devset_ready) {
        g_devset_ready = true;
        g_devset.clear();
        const char *e = getenv("TRIMAL_CUDA_DEVICES");
        if (e && *e) {
            if (!strcmp(e, "all")) {
                g_devset = usable_devices();
            } else {
                for (const char *p = e; *p;) {
                    char *end = nullptr;
                    const long d = strtol(p, &end, 10);
                    if (end == p) break;
                    if (d >= 0 && device_usable((int)d, nullptr) &&
                        std::find(g_devset.begin(), g_devset.end(), (int)d) == g_devset.end())
                        g_devset.push_back((int)d);
                    p = *end == ',' ? end + 1 : end;
                    if (*end && *end != ',') break;
                }
            }
        }
        if (g_devset.empty()) g_devset.push_back(0);
    }
    return g_devset;
}

// alignments smaller than this stay on the first device of the set (threads and peer copies
// cost more than they save); TRIMAL_CUDA_MULTI_MIN_BYTES overrides (tests)
size_t multi_min_bytes()
{
    const char *e = getenv("TRIMAL_CUDA_MULTI_MIN_BYTES");
    return e && *e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(4u << 20);
}
}  // namespace

extern "C" int tcu_set_devices(const int *devices, int count)
{
    if (count < 0 || (count > 0 && !devices)) return fail(TCU_ERR_INVALID, "bad device list");
    std::vector<int> v;
    for (int k = 0; k < count; k++) {
        if (!device_usable(devices[k], nullptr))
            return fail(TCU_ERR_NO_DEVICE, "device %d is not an sm_100 GPU", devices[k]);
        if (std::find(v.begin(), v.end(), devices[k]) != v.end())
            return fail(TCU_ERR_INVALID, "device %d listed twice", devices[k]);
        v.push_back(devices[k]);
    }
    std::lock_guard<std::mutex> lk(g_devset_mutex);
    if (count == 0) {
        g_devset_ready = false;  // back to TRIMAL_CUDA_DEVICES / device 0
    } else {
        g_devset = v;
        g_devset_ready = true;
    }
    return TCU_OK;
}

extern "C" int tcu_get_devices(int *devices, int max)
{
    const std::vector<int> v = configured_devices();
    for (int k = 0; k < (int)v.size() && k < max; k++)
        if (devices) devices[k] = v[k];
    return (int)v.size();
}

extern "C" int tcu_msa_device_count(const tcu_msa *m) { return m ? 1 + (int)m->peers.size() : 0; }

// every replica of a handle, the caller's own first
static std::vector<tcu_msa *> replicas(tcu_msa *m)
{
    std::vector<tcu_msa *> v{m};
    v.insert(v.end(), m->peers.begin(), m->peers.end());
    return v;
}

// f(replica, index) on every replica at once, one host thread per further device (the calls
// block on their streams); the first failure wins and its message is carried over from the
// worker thread (error strings are thread-local).
template <class F>
static int for_each_replica(tcu_msa *m, F f)
{
    std::vector<tcu_msa *> hs = replicas(m);
    if (hs.size() == 1) return f(hs[0], 0);
    std::vector<int> rcs(hs.size(), TCU_OK);
    std::vector<std::string> errs(hs.size());
    auto run = [&](int k) {
        rcs[k] = f(hs[k], k);
        if (rcs[k] != TCU_OK) errs[k] = g_last_error;
    };
    std::vector<std::thread> pool;
    size_t started = 1;
    try {
        for (; started < hs.size(); started++) pool.emplace_back(run, (int)started);
    } catch (...) {  // no thread to be had: the rest runs here, nothing may escape the C ABI
    }
    run(0);
    for (size_t k = started; k < hs.size(); k++) run((int)k);
    for (auto &t : pool) t.join();
    for (size_t k = 0; k < hs.size(); k++)
        if (rcs[k] != TCU_OK) return fail(rcs[k], "%s", errs[k].c_str());
    return TCU_OK;
}

// direct access between every pair of devices of a set (NVLink / NVSwitch); copies between
// devices without it still work, staged by the driver
// (returns whether every device can address the memory of the first one: kernels on the
// other devices may then store into it directly)
static bool enable_peer_access(const std::vector<int> &devs)
{
    bool first_reachable = true;
    for (int a : devs) {
        if (cudaSetDevice(a) != cudaSuccess) {
            first_reachable = false;
            continue;
        }
        for (int b : devs) {
            if (a == b) continue;
            int can = 0;
            bool ok = false;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
            }
            if (b == devs[0] && !ok) first_reachable = false;
        }
    }
    cudaGetLastError();  // cudaErrorPeerAccessAlreadyEnabled
    return first_reachable;
}

// the caller's rows: separate row pointers, or one strided block (page-locked or not)
struct HostRows {
    const char *const *rows = nullptr;
    const uint8_t *data = nullptr;
    size_t stride = 0;
    bool pinned = false;
};

static int upload_range(tcu_msa *m, const HostRows &h, int r0, int r1)
{
    if (h.rows) return upload_rows(m, [&](int r) { return (const void *)h.rows[r]; }, r0, r1);
    if (h.pinned) return upload_strided_pinned(m, h.data, h.stride, r0, r1);
    return upload_rows(m, [&](int r) { return (const void *)(h.data + (size_t)r * h.stride); }, r0,
                       r1);
}

extern "C" void tcu_msa_destroy(tcu_msa *m);

static int msa_create_any(const HostRows &h, int nseq, int ncol, int device, tcu_msa **out)
{
    if (!out) return fail(TCU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    std::vector<int> devs;
    if (device == TCU_DEVICE_AUTO) {
        devs = configured_devices();
        if ((size_t)std::max(nseq, 0) * (size_t)std::max(ncol, 0) < multi_min_bytes()) devs.resize(1);
    } else {
        devs.push_back(device);
    }
    static const bool trace = getenv("TCU_TRACE") != nullptr;
    const double t0 = trace ? now_ms() : 0;
    tcu_msa *m = nullptr;
    int rc = msa_alloc(nseq, ncol, devs[0], &m);
    if (rc != TCU_OK) return rc;
    for (size_t k = 1; k < devs.size() && rc == TCU_OK; k++) {
        tcu_msa *p = nullptr;
        rc = msa_alloc(nseq, ncol, devs[k], &p);
        if (rc == TCU_OK) m->peers.push_back(p);
    }
    const double t1 = trace ? now_ms() : 0;
    if (rc == TCU_OK && devs.size() == 1) {
        rc = upload_range(m, h, 0, nseq);
    } else if (rc == TCU_OK) {
        // every device takes 1/N of the rows over its own PCIe link, at their place in its
        // copy of the matrix; the shards are then exchanged device to device
        m->peer_direct = enable_peer_access(devs);
        const int world = (int)devs.size();
        rc = for_each_replica(m, [&](tcu_msa *r, int k) {
            int r0, r1;
            tcu_shard_range(nseq, 1, k, world, &r0, &r1);
            return upload_range(r, h, r0, r1);
        });
        if (rc == TCU_OK) {
            std::vector<tcu_msa *> hs = replicas(m);
            cudaError_t e = cudaSuccess;
            for (int k = 0; k < world && e == cudaSuccess; k++) {
                e = cudaSetDevice(hs[k]->device);
                for (int o = 0; o < world && e == cudaSuccess; o++) {
                    int r0, r1;
                    tcu_shard_range(nseq, 1, o, world, &r0, &r1);
                    if (o == k || r1 <= r0) continue;
                    e = cudaMemcpyPeerAsync(hs[k]->d_raw + (size_t)r0 * m->pitch, hs[k]->device,
                                            hs[o]->d_raw + (size_t)r0 * m->pitch, hs[o]->device,
                                            (size_t)(r1 - r0) * m->pitch, hs[k]->stream);
                }
            }
            for (int k = 0; k < world; k++) {
                cudaSetDevice(hs[k]->device);
                cudaError_t e2 = cudaStreamSynchronize(hs[k]->stream);
                if (e == cudaSuccess) e = e2;
            }
            if (e != cudaSuccess) rc = cuda_fail(e, "row exchange between devices");
        }
    }
    if (trace)
        fprintf(stderr, "[tcu] create on %zu device(s): alloc %.2f ms, upload(%s) %.2f ms (h2d events %.2f)\n",
                devs.size(), t1 - t0, h.rows ? "rows" : (h.pinned ? "pinned" : "staged"),
                now_ms() - t1, rc == TCU_OK ? m->timings.h2d_ms : -1.f);
    if (rc != TCU_OK) {
        const std::string keep = g_last_error;
        tcu_msa_destroy(m);
        g_last_error = keep;
        return rc;
    }
    cudaSetDevice(m->device);
    *out = m;
    return TCU_OK;
}

extern "C" int tcu_msa_create(const char *const *rows, int nseq, int ncol, int device, tcu_msa **out)
{
    if (nseq > 0 && !rows) return fail(TCU_ERR_INVALID, "rows is NULL");
    HostRows h;
    h.rows = rows;
    return msa_create_any(h, nseq, ncol, device, out);
}

extern "C" int tcu_msa_create_strided(const uint8_t *data, int nseq, int ncol, size_t stride,
                                      int device, tcu_msa **out)
{
    if (nseq > 0 && ncol > 0 && !data) return fail(TCU_ERR_INVALID, "data is NULL");
    if (stride < (size_t)ncol) return fail(TCU_ERR_INVALID, "stride smaller than ncol");
    HostRows h;
    h.data = data;
    h.stride = stride;
    h.pinned = nseq > 0 && ncol > 0 && stride <= 2 * (size_t)ncol + 64 && is_pinned_host(data);
    return msa_create_any(h, nseq, ncol, device, out);
}

// One process per GPU: every rank holds the same alignment on the host; each uploads only
// its share of the rows and the shares are all-gathered over NVLink, instead of N copies of
// the whole alignment crossing the host's PCIe complex at once.
extern "C" int tcu_msa_create_all(tcu_comm *comm, const uint8_t *data, int nseq, int ncol,
                                  size_t stride, tcu_msa **out)
{
    NvtxRange nvtx("tcu_msa_create_all");
    if (!comm || !out) return fail(TCU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    if (nseq > 0 && ncol > 0 && !data) return fail(TCU_ERR_INVALID, "data is NULL");
    if (stride < (size_t)ncol) return fail(TCU_ERR_INVALID, "stride smaller than ncol");
    *out = nullptr;
    HostRows h;
    h.data = data;
    h.stride = stride;
    h.pinned = nseq > 0 && ncol > 0 && stride <= 2 * (size_t)ncol + 64 && is_pinned_host(data);
    tcu_msa *m = nullptr;
    int rc = msa_alloc(nseq, ncol, comm->device, &m);
    if (rc != TCU_OK) return rc;
    int r0 = 0, r1 = 0;
    tcu_shard_range(nseq, 1, comm->rank, comm->world, &r0, &r1);
    // A page-locked buffer goes up in one copy and a layout kernel; with peer memory that
    // kernel writes the rank's rows into every rank's matrix (NVLink) and a one-word
    // all-reduce is all that is left of the exchange.
    std::vector<void *> peer_base;
    if (h.pinned && nseq > 0 && ncol > 0 && comm->world > 1 && comm->world - 1 <= ID2_MAX_PEERS &&
        peer_prepare(comm, m->d_raw, peer_base, m->stream)) {
        std::vector<uint8_t *> others;
        for (int q = 1; q < comm->world; q++)
            others.push_back((uint8_t *)peer_base[(size_t)((comm->rank + q) % comm->world)]);
        rc = upload_strided_pinned(m, h.data, h.stride, r0, r1, &others);
        if (rc == TCU_OK) {
            cudaEventRecord(m->ev[4], m->stream);
            rc = comm_allreduce_i32(comm, (int *)(comm->d_sync + 64 * 64 + 64), 1, m->stream);
            cudaEventRecord(m->ev[5], m->stream);
            if (rc == TCU_OK && cudaStreamSynchronize(m->stream) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "row exchange between ranks");
            if (rc == TCU_OK) m->timings.comm_ms = ev_ms(m->ev[4], m->ev[5]);
        }
        if (rc != TCU_OK) {
            const std::string keep = g_last_error;
            tcu_msa_destroy(m);
            g_last_error = keep;
            return rc;
        }
        *out = m;
        return TCU_OK;
    }
    rc = upload_range(m, h, r0, r1);
    if (rc == TCU_OK && nseq > 0) {
        std::vector<size_t> off(comm->world), cnt(comm->world);
        for (int r = 0; r < comm->world; r++) {
            int a, b;
            tcu_shard_range(nseq, 1, r, comm->world, &a, &b);
            off[r] = (size_t)a * m->pitch;
            cnt[r] = (size_t)(b - a) * m->pitch;
        }
        cudaEventRecord(m->ev[4], m->stream);
        rc = comm_allgatherv(comm, m->d_raw, off.data(), cnt.data(), m->stream);
        cudaEventRecord(m->ev[5], m->stream);
        if (rc == TCU_OK && cudaStreamSynchronize(m->stream) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "row exchange between ranks");
        if (rc == TCU_OK) m->timings.comm_ms = ev_ms(m->ev[4], m->ev[5]);
    }
    if (rc != TCU_OK) {
        const std::string keep = g_last_error;
        tcu_msa_destroy(m);
        g_last_error = keep;
        return rc;
    }
    *out = m;
    return TCU_OK;
}

extern "C" void tcu_msa_destroy(tcu_msa *m)
{
    if (!m) return;
    for (tcu_msa *p : m->peers) tcu_msa_destroy(p);
    m->peers.clear();
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
    dev_cache_give(m->device, m->d_raw, m->raw_cap);
    dev_cache_give(m->device, m->d_planes, m->planes_cap);
    dev_cache_give(m->device, m->d_gbytes, m->gbytes_cap);
    dev_cache_give(m->device, m->d_ident, m->ident_cap);
    dev_cache_give(m->device, m->d_bits, m->bits_cap);
    dev_cache_give(m->device, m->d_brows, m->brows_cap);
    dev_cache_give(m->device, m->d_scratch, m->scratch_cap);
    dev_cache_give(m->device, m->d_small, m->small_cap);
    for (auto &e : m->band_done)
        if (e) cudaEventDestroy(e);
    {
        StreamSet s;  // both streams were synchronised above
        s.stream = m->stream;
        s.copy_stream = m->copy_stream;
        for (int i = 0; i < 6; i++) s.ev[i] = m->ev[i];
        for (int i = 0; i < 2; i++) s.cev[i] = m->cev[i];
        streams_give(m->device, s);
    }
    cudaGetLastError();
    delete m;
}

extern "C" int tcu_msa_nseq(const tcu_msa *m) { return m ? m->nseq : 0; }
extern "C" int tcu_msa_ncol(const tcu_msa *m) { return m ? m->ncol : 0; }
extern "C" void *tcu_msa_stream(tcu_msa *m) { return m ? (void *)m->stream : nullptr; }
extern "C" int tcu_msa_device(const tcu_msa *m) { return m ? m->device : -1; }

extern "C" int tcu_msa_sync(tcu_msa *m)
{
    if (!m) return fail(TCU_ERR_INVALID, "msa is NULL");
    CK(cudaSetDevice(m->device));
    CK(cudaStreamSynchronize(m->stream));
    if (m->pending_device_timing) {
        m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);
        m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);
        m->pending_device_timing = false;
    }
    return TCU_OK;
}

extern "C" int tcu_msa_timings(const tcu_msa *m, tcu_timings *out)
{
    if (!m || !out) return fail(TCU_ERR_INVALID, "NULL argument");
    *out = m->timings;
    return TCU_OK;
}

// device -> host copy of a possibly multi-GB result
static int download(tcu_msa *m, void *dst, const void *d_src, size_t bytes)
{
    if (bytes == 0) return TCU_OK;
    CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, m->stream));
    return TCU_OK;
}

// ---------------------------------------------------------------------------
// K3 gaps
// ---------------------------------------------------------------------------
// comm: this rank's share of the rows, all-reduced on the device.  Without comm: the share
// (srank of sworld) only, no reduction -- 0 of 1 is the whole statistic.
static int gaps_impl(tcu_msa *m, tcu_comm *comm, int srank, int sworld, const int *save_seq,
                     int *gaps_in_column, int *num_cols_with_gaps, int *max_gaps)
{
    if (!m || !gaps_in_column) return fail(TCU_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(m->device));
    int rc0 = comm_check(m, comm);
    if (rc0 != TCU_OK) return rc0;
    m->timings = tcu_timings{};
    const int n = m->nseq, L = m->ncol;
    if (L == 0) return TCU_OK;
    const size_t cnt_bytes = ((size_t)L * sizeof(int) + 255) / 256 * 256;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, cnt_bytes + (size_t)n + 256, m);
    if (rc != TCU_OK) return rc;
    int *d_cnt = (int *)m->d_scratch;
    uint8_t *d_drop = nullptr;
    CK(cudaEventRecord(m->ev[0], m->stream));
    std::vector<uint8_t> drop;
    if (save_seq) {
        drop.resize(n);
        for (int i = 0; i < n; i++) drop[i] = save_seq[i] == -1;
        d_drop = (uint8_t *)m->d_scratch + cnt_bytes;
        CK(cudaMemcpyAsync(d_drop, drop.data(), n, cudaMemcpyHostToDevice, m->stream));
    }
    CK(cudaMemsetAsync(d_cnt, 0, (size_t)L * sizeof(int), m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    // several ranks: each counts its share of the rows, the integer partial counts are
    // summed across GPUs (exact in any order)
    int r0 = 0, r1 = n;
    if (comm) tcu_shard_range(n, 1, comm->rank, comm->world, &r0, &r1);
    else tcu_shard_range(n, 1, srank, sworld, &r0, &r1);
    CK(launch_column_counts(m->d_raw + (size_t)r0 * m->pitch, r1 - r0, L, m->pitch,
                            d_drop ? d_drop + r0 : nullptr, '-', '-', d_cnt, nullptr, nullptr, nullptr,
                            m->num_sms, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    if (comm) {
        rc = comm_allreduce_i32(comm, d_cnt, (size_t)L, m->stream);
        if (rc != TCU_OK) return rc;
    }
    CK(cudaEventRecord(m->ev[5], m->stream));
    CK(cudaMemcpyAsync(gaps_in_column, d_cnt, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost,
                       m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.comm_ms = ev_ms(m->ev[2], m->ev[5]);
    m->timings.d2h_ms = ev_ms(m->ev[5], m->ev[3]);
    m->timings.kernel_launches = r1 > r0 ? 1 : 0;
    // histogram and maximum (template.h:496-501)
    for (int k = 0; k < L; k++) {
        if (num_cols_with_gaps) num_cols_with_gaps[gaps_in_column[k]]++;
        if (max_gaps && gaps_in_column[k] > *max_gaps) *max_gaps = gaps_in_column[k];
    }
    return TCU_OK;
}

extern "C" int tcu_gaps(tcu_msa *m, const int *save_seq, int *gaps_in_column,
                        int *num_cols_with_gaps, int *max_gaps)
{
    NvtxRange nvtx("tcu_gaps");
    if (!m || !gaps_in_column) return fail(TCU_ERR_INVALID, "NULL argument");
    if (m->peers.empty())
        return gaps_impl(m, nullptr, 0, 1, save_seq, gaps_in_column, num_cols_with_gaps, max_gaps);
    // several devices: each counts its share of the rows; the integer partial counts are
    // summed on the host (exact in any order)
    const int world = 1 + (int)m->peers.size(), L = m->ncol;
    std::vector<std::vector<int>> part((size_t)world, std::vector<int>((size_t)std::max(L, 1), 0));
    int rc = for_each_replica(m, [&](tcu_msa *r, int k) {
        return gaps_impl(r, nullptr, k, world, save_seq, part[k].data(), nullptr, nullptr);
    });
    if (rc != TCU_OK) return rc;
    for (int k = 0; k < L; k++) {
        int c = 0;
        for (int d = 0; d < world; d++) c += part[d][k];
        gaps_in_column[k] = c;
        if (num_cols_with_gaps) num_cols_with_gaps[c]++;
        if (max_gaps && c > *max_gaps) *max_gaps = c;
    }
    for (tcu_msa *p : m->peers) m->timings.kernel_ms = std::max(m->timings.kernel_ms, p->timings.kernel_ms);
    return TCU_OK;
}

extern "C" int tcu_gaps_all(tcu_msa *m, tcu_comm *comm, const int *save_seq, int *gaps_in_column,
                            int *num_cols_with_gaps, int *max_gaps)
{
    if (!comm) return fail(TCU_ERR_INVALID, "comm is NULL");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    return gaps_impl(m, comm, 0, 1, save_seq, gaps_in_column, num_cols_with_gaps, max_gaps);
}

// ---------------------------------------------------------------------------
// K0 + K1 identity
// ---------------------------------------------------------------------------
extern "C" int tcu_identity_band_rows(void) { return IB; }
extern "C" int tcu_identity_row_blocks(int kept_rows) { return (kept_rows + IB - 1) / IB; }

extern "C" long long tcu_identity_tiles_before(int kept_rows, int block)
{
    const int nb = (kept_rows + RB - 1) / RB, nsb = (kept_rows + IB - 1) / IB;
    return tiles_before2(std::max(0, std::min(block, nsb)), nb);
}

extern "C" int tcu_identity_tile(int kept_rows, int block_begin, int block_end, long long tile,
                                 int *row_block, int *col_block64)
{
    const int nb = (kept_rows + RB - 1) / RB, nsb = (kept_rows + IB - 1) / IB;
    if (!row_block || !col_block64 || block_begin < 0 || block_end > nsb || block_begin >= block_end ||
        tile < tiles_before2(block_begin, nb) || tile >= tiles_before2(block_end, nb))
        return fail(TCU_ERR_INVALID, "tile %lld outside row-blocks [%d, %d) of %d kept rows", tile,
                    block_begin, block_end, kept_rows);
    tile_to_blocks2(tile, nb, block_begin, block_end, *row_block, *col_block64);
    return TCU_OK;
}

extern "C" size_t tcu_identity_row_offset(int kept_rows, int i)
{
    const size_t n = (size_t)std::max(kept_rows, 0);
    size_t r = (size_t)std::max(i, 0);
    if (n < 2) return 0;
    if (r > n - 1) r = n - 1;
    return r * n - r * (r + 1) / 2;
}

// which of the 256 byte values occur in the alignment (mask independent, computed once)
static int ensure_present(tcu_msa *m)
{
    if (m->have_present) return TCU_OK;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, 256 * sizeof(unsigned int), m);
    if (rc != TCU_OK) return rc;
    unsigned int *d_present = (unsigned int *)m->d_scratch;
    CK(cudaMemsetAsync(d_present, 0, 256 * sizeof(unsigned int), m->stream));
    CK(launch_byte_presence(m->d_raw, m->nseq, m->ncol, m->pitch, d_present, m->num_sms, m->stream));
    CK(cudaMemcpyAsync(m->present, d_present, sizeof m->present, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->have_present = true;
    m->timings.kernel_launches++;
    return TCU_OK;
}

// more distinct non-gap byte values than the 7 code planes of the packed operand hold
static bool bytewise_alphabet(tcu_msa *m, uint8_t indet)
{
    if (cudaSetDevice(m->device) != cudaSuccess || ensure_present(m) != TCU_OK) return false;
    int count = 0;
    for (int b = 0; b < 256; b++) count += b != '-' && b != indet && m->present[b];
    return count > (1 << MAX_PLANES) - 2;
}

extern "C" int tcu_identity_prepare(tcu_msa *m, const int *save_seq, const int *save_res,
                                    uint8_t indet, int *kept_rows_out)
{
    if (!m) return fail(TCU_ERR_INVALID, "msa is NULL");
    CK(cudaSetDevice(m->device));
    m->prepared = false;
    const int n = m->nseq, L = m->ncol;

    std::vector<int> kept;
    kept.reserve(n);
    for (int i = 0; i < n; i++)
        if (!save_seq || save_seq[i] != -1) kept.push_back(i);
    std::vector<uint8_t> drop((size_t)std::max(L, 1), 0);
    if (save_res)
        for (int k = 0; k < L; k++) drop[k] = save_res[k] == -1;

    CK(cudaEventRecord(m->ev[0], m->stream));
    int prc = ensure_present(m);
    if (prc != TCU_OK) return prc;

    // dense residue codes for the bytes that occur; the gap class shares one code
    uint8_t lut[256];
    int count = 0;
    for (int b = 0; b < 256; b++) {
        if (b == '-' || b == indet) lut[b] = CODE_GAP;
        else if (m->present[b]) lut[b] = (uint8_t)std::min(count++, 254);
        else lut[b] = 0;
    }
    int np = MIN_PLANES;
    while (np <= MAX_PLANES && (1 << np) - 2 < count) np++;
    // more distinct symbols than 7 code planes hold: no bit-plane operand, the byte-wise
    // kernel (identity_bytes.cu) computes the statistic from the raw rows
    const bool bytewise = np > MAX_PLANES;
    if (bytewise) np = 0;

    m->np = np;
    m->nk = (int)kept.size();
    m->nb = (m->nk + RB - 1) / RB;
    m->nsb = (m->nk + IB - 1) / IB;
    m->nb2 = 2 * m->nsb;
    m->nchunks = (L + KC2 * 32 - 1) / (KC2 * 32);
    m->prepared_indet = indet;

    if (!m->d_small) {
        const size_t rows_b = (std::max<size_t>(1, (size_t)n) * sizeof(int) + 255) / 256 * 256;
        int src = ensure_dev(m->device, &m->d_small, &m->small_cap, rows_b + m->pitch + 256, m);
        if (src != TCU_OK) return src;
        m->d_kept_rows = (int *)m->d_small;
        m->d_col_drop = (uint8_t *)m->d_small + rows_b;
        m->d_lut = m->d_col_drop + m->pitch;
    }
    CK(cudaMemcpyAsync(m->d_lut, lut, 256, cudaMemcpyHostToDevice, m->stream));
    if (m->nk)
        CK(cudaMemcpyAsync(m->d_kept_rows, kept.data(), (size_t)m->nk * sizeof(int),
                           cudaMemcpyHostToDevice, m->stream));
    if (L) CK(cudaMemcpyAsync(m->d_col_drop, drop.data(), L, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));  // the host vectors go out of scope

    if (bytewise) {
        CK(cudaEventRecord(m->ev[1], m->stream));
        CK(cudaEventRecord(m->ev[2], m->stream));
        m->prepared = true;
        if (kept_rows_out) *kept_rows_out = m->nk;
        return TCU_OK;
    }
    int rc = ensure_dev(m->device, (void **)&m->d_planes, &m->planes_cap, (size_t)m->nb2 * m->nchunks * tile2_bytes(np), m);
    if (rc != TCU_OK) return rc;
    rc = ensure_dev(m->device, (void **)&m->d_gbytes, &m->gbytes_cap,
                (size_t)m->nb2 * m->nchunks * G_STAGES_PER_CHUNK * G_BLOCK_BYTES, m);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_pack_planes(m->d_raw, m->pitch, L, m->d_kept_rows, m->nk, m->d_col_drop, m->d_lut, np,
                          m->nb2, m->nchunks, m->d_planes, m->d_gbytes, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    if (m->nb && m->nchunks) m->timings.kernel_launches++;
    m->prepared = true;
    if (kept_rows_out) *kept_rows_out = m->nk;
    return TCU_OK;
}

// super-blocks [sb_begin, sb_end) of IB = 128 kept rows each
static int identity_launch(tcu_msa *m, int sb_begin, int sb_end, float *d_out, int *d_hit,
                           int *d_dst, uint32_t *d_bits = nullptr, float thr = 0.f,
                           const std::vector<uint32_t *> *bits_peers = nullptr)
{
    if (m->nchunks == 0 || sb_end <= sb_begin) return TCU_OK;
    if (m->np == 0)
        return fail(TCU_ERR_INVALID, "more than %d distinct symbols: only tcu_identity and "
                    "tcu_representatives handle such an alignment", (1 << MAX_PLANES) - 2);
    Identity2Params p{};
    p.planes = m->d_planes;
    p.gbytes = m->d_gbytes;
    p.out = d_out;
    p.hit_out = d_hit;
    p.dst_out = d_dst;
    p.bits_out = d_bits;
    p.thr = threshold_rule(thr);
    if (d_bits && (size_t)((m->nk + 127) / 128) * (size_t)m->nk >= 0xFFFFFFFFull)
        return fail(TCU_ERR_INVALID, "%d sequences: the threshold bit matrix is limited to 2^32 entries", m->nk);
    if (bits_peers) {
        if (bits_peers->size() > (size_t)ID2_MAX_PEERS) return fail(TCU_ERR_INVALID, "too many peer matrices");
        for (uint32_t *q : *bits_peers) p.bits_peer[p.n_bits_peer++] = q;
    }
    p.nb = m->nb;
    p.nb2 = m->nb2;
    p.nchunks = m->nchunks;
    p.nk = m->nk;
    p.total_bits = m->nchunks * KC2 * 32;
    p.tile_begin = tiles_before2(sb_begin, m->nb);
    p.tile_end = tiles_before2(sb_end, m->nb);
    p.sb_begin = sb_begin;
    p.sb_end = sb_end;
    p.out_base = tcu_identity_row_offset(m->nk, sb_begin * IB);
    if (p.tile_end <= p.tile_begin) return TCU_OK;
    CK(launch_identity2(m->np, p, m->num_sms, m->stream));
    m->timings.kernel_launches++;
    return TCU_OK;
}

// Identity of super-blocks [sb_begin, sb_end) into d_out (element 0 = packed offset of
// row sb_begin * IB) and, when `host` is given, on into host memory: the range is cut
// into sub-bands of equal work; the kernel of sub-band s+1 runs while sub-band s
// crosses PCIe on the copy stream.  ev[3] marks the end of the last kernel; cev[0..1]
// bracket the copies.  Returns with both streams idle when `host` is given.
static int identity_pipeline(tcu_msa *m, int sb_begin, int sb_end, float *d_out, float *host,
                             int *d_hit, int *d_dst)
{
    constexpr int MAX_SUB = 16;
    const long long t0 = tiles_before2(sb_begin, m->nb), t1 = tiles_before2(sb_end, m->nb);
    // sub-bands only pay off when each still fills the GPU for several waves
    int nsub = host ? (int)std::min<long long>(MAX_SUB, (t1 - t0) / (8LL * 2 * m->num_sms)) : 1;
    nsub = std::max(1, std::min(nsub, sb_end - sb_begin));
    while ((int)m->band_done.size() < nsub) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->band_done.push_back(e);
    }
    const size_t base = tcu_identity_row_offset(m->nk, sb_begin * IB);
    int b0 = sb_begin;
    for (int s = 0; s < nsub; s++) {
        int b1 = sb_end;
        if (s + 1 < nsub) {
            const long long target = t0 + (t1 - t0) * (s + 1) / nsub;
            b1 = b0;
            while (b1 < sb_end && tiles_before2(b1, m->nb) < target) b1++;
        }
        const size_t lo = tcu_identity_row_offset(m->nk, b0 * IB);
        const size_t hi = tcu_identity_row_offset(m->nk, std::min(b1 * IB, m->nk));
        if (b1 > b0) {
            int rc = identity_launch(m, b0, b1, d_out + (lo - base), d_hit, d_dst);
            if (rc != TCU_OK) return rc;
        }
        if (s + 1 == nsub) CK(cudaEventRecord(m->ev[3], m->stream));
        if (host && hi > lo) {
            CK(cudaEventRecord(m->band_done[s], m->stream));
            CK(cudaStreamWaitEvent(m->copy_stream, m->band_done[s], 0));
            if (s == 0) CK(cudaEventRecord(m->cev[0], m->copy_stream));
            CK(cudaMemcpyAsync(host + (lo - base), d_out + (lo - base), (hi - lo) * sizeof(float),
                               cudaMemcpyDeviceToHost, m->copy_stream));
        }
        b0 = b1;
    }
    if (host) {
        CK(cudaEventRecord(m->cev[1], m->copy_stream));
        CK(cudaStreamSynchronize(m->copy_stream));
        CK(cudaStreamSynchronize(m->stream));
        m->timings.d2h_ms = ev_ms(m->cev[0], m->cev[1]);
    }
    return TCU_OK;
}

extern "C" int tcu_identity_device(tcu_msa *m, int block_begin, int block_end, float *d_out)
{
    if (!m || !d_out) return fail(TCU_ERR_INVALID, "NULL argument");
    if (!m->prepared) return fail(TCU_ERR_STATE, "tcu_identity_prepare has not been called");
    if (block_begin < 0 || block_end > m->nsb || block_begin > block_end)
        return fail(TCU_ERR_INVALID, "row-block range [%d,%d) outside [0,%d)", block_begin,
                    block_end, m->nsb);
    CK(cudaSetDevice(m->device));
    m->timings.kernel_launches = 0;
    int rc = identity_launch(m, block_begin, block_end, d_out, nullptr, nullptr);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[3], m->stream));
    m->pending_device_timing = true;
    return TCU_OK;
}

// Host-facing band variant: row-blocks [block_begin, block_end) of the packed
// array into a host slice whose element 0 is packed offset
// tcu_identity_row_offset(kept, 64*block_begin).  block_end < 0 = all blocks.
extern "C" int tcu_identity_band(tcu_msa *m, const int *save_seq, const int *save_res,
                                 uint8_t indet, int block_begin, int block_end, float *identities)
{
    if (!m || !identities) return fail(TCU_ERR_INVALID, "NULL argument");
    m->timings = tcu_timings{};
    m->ident_full = false;
    int rc = tcu_identity_prepare(m, save_seq, save_res, indet, nullptr);
    if (rc != TCU_OK) return rc;
    if (block_end < 0) block_end = m->nsb;
    if (block_begin < 0 || block_end > m->nsb || block_begin > block_end)
        return fail(TCU_ERR_INVALID, "row-block range [%d,%d) outside [0,%d)", block_begin,
                    block_end, m->nsb);
    const size_t lo = tcu_identity_row_offset(m->nk, block_begin * IB);
    const size_t hi = tcu_identity_row_offset(m->nk, std::min(block_end * IB, m->nk));
    const size_t count = hi - lo;
    if (count == 0) {
        CK(cudaStreamSynchronize(m->stream));
        return TCU_OK;
    }
    rc = ensure_ident(m, count * sizeof(float));
    if (rc != TCU_OK) return rc;
    rc = identity_pipeline(m, block_begin, block_end, m->d_ident, identities, nullptr, nullptr);
    if (rc != TCU_OK) return rc;
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);  // overlaps the copies
    return TCU_OK;
}

static int identity_host(tcu_msa *m, const int *save_seq, const int *save_res, uint8_t indet,
                         float *identities, int *hit_out, int *dst_out, int keep_on_device,
                         bool bytes_kernel)
{
    if (!m) return fail(TCU_ERR_INVALID, "msa is NULL");
    if (!identities && !keep_on_device)
        return fail(TCU_ERR_INVALID, "identities is NULL and keep_on_device is 0");
    m->timings = tcu_timings{};
    m->ident_full = false;
    int rc = tcu_identity_prepare(m, save_seq, save_res, indet, nullptr);
    if (rc != TCU_OK) return rc;
    const size_t npairs = (size_t)m->nk * (size_t)std::max(m->nk - 1, 0) / 2;
    if (npairs == 0) {
        CK(cudaStreamSynchronize(m->stream));
        return TCU_OK;
    }
    rc = ensure_ident(m, npairs * sizeof(float));
    if (rc != TCU_OK) return rc;
    int *d_hit = nullptr, *d_dst = nullptr;
    if (hit_out || dst_out) {
        // the kernel writes both or neither
        rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, 2 * npairs * sizeof(int), m);
        if (rc != TCU_OK) return rc;
        d_hit = (int *)m->d_scratch;
        d_dst = d_hit + npairs;
    }
    if (bytes_kernel || m->np == 0) {
        CK(launch_identity_bytes(m->d_raw, m->pitch, m->ncol, m->d_kept_rows, m->nk, m->d_col_drop,
                                 indet, m->d_ident, d_hit, d_dst, m->num_sms, m->stream));
        m->timings.kernel_launches++;
        CK(cudaEventRecord(m->ev[3], m->stream));
        if (identities) rc = download(m, identities, m->d_ident, npairs * sizeof(float));
    } else {
        rc = identity_pipeline(m, 0, m->nsb, m->d_ident, identities, d_hit, d_dst);
    }
    if (rc != TCU_OK) return rc;
    const float pipelined_d2h_ms = m->timings.d2h_ms;
    if (rc == TCU_OK && hit_out) rc = download(m, hit_out, d_hit, npairs * sizeof(int));
    if (rc == TCU_OK && dst_out) rc = download(m, dst_out, d_dst, npairs * sizeof(int));
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[4], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.d2h_ms = pipelined_d2h_ms + ev_ms(m->ev[3], m->ev[4]);
    m->ident_full = (m->nk == m->nseq);
    if (!keep_on_device) {
        dev_cache_give(m->device, m->d_ident, m->ident_cap);
        m->d_ident = nullptr;
        m->ident_cap = 0;
        m->ident_full = false;
    }
    return TCU_OK;
}

// The same on the replicas of a multi-device handle: the pair matrix is cut into row-block
// bands of equal work (tcu_shard_blocks), every device computes its band and copies it
// into its slice of the caller's array over its own PCIe link; with keep_on_device the bands
// are also collected on the first device (device-to-device) for the consumers that follow.
static int identity_multi(tcu_msa *m, const int *save_seq, const int *save_res, uint8_t indet,
                          float *identities, int keep_on_device)
{
    if (!identities && !keep_on_device)
        return fail(TCU_ERR_INVALID, "identities is NULL and keep_on_device is 0");
    const int world = 1 + (int)m->peers.size();
    std::vector<size_t> lo((size_t)world), hi((size_t)world);
    int rc = for_each_replica(m, [&](tcu_msa *r, int k) -> int {
        r->timings = tcu_timings{};
        r->ident_full = false;
        int e = tcu_identity_prepare(r, save_seq, save_res, indet, nullptr);
        if (e != TCU_OK) return e;
        int b0 = 0, b1 = 0;
        tcu_shard_blocks(r->nk, k, world, &b0, &b1);
        lo[k] = tcu_identity_row_offset(r->nk, std::min(b0 * IB, r->nk));
        hi[k] = tcu_identity_row_offset(r->nk, std::min(b1 * IB, r->nk));
        const size_t npairs = (size_t)r->nk * (size_t)std::max(r->nk - 1, 0) / 2;
        if (npairs == 0) {
            CK(cudaStreamSynchronize(r->stream));
            return TCU_OK;
        }
        // the first device holds the whole array when the matrix is to stay resident
        const bool whole = k == 0 && keep_on_device;
        e = ensure_ident(r, std::max<size_t>(whole ? npairs : hi[k] - lo[k], 1) * sizeof(float));
        if (e != TCU_OK) return e;
        float *d_band = whole ? r->d_ident + lo[k] : r->d_ident;
        if (hi[k] > lo[k]) {
            e = identity_pipeline(r, b0, b1, d_band, identities ? identities + lo[k] : nullptr,
                                  nullptr, nullptr);
            if (e != TCU_OK) return e;
        }
        CK(cudaStreamSynchronize(r->stream));
        r->timings.h2d_ms = ev_ms(r->ev[0], r->ev[1]);
        r->timings.pack_ms = ev_ms(r->ev[1], r->ev[2]);
        if (hi[k] > lo[k]) r->timings.kernel_ms = ev_ms(r->ev[2], r->ev[3]);
        return TCU_OK;
    });
    if (rc != TCU_OK) return rc;
    if (keep_on_device && m->nk >= 2) {
        CK(cudaSetDevice(m->device));
        CK(cudaEventRecord(m->ev[4], m->stream));
        for (int k = 1; k < world; k++) {
            tcu_msa *p = m->peers[(size_t)k - 1];
            if (hi[k] > lo[k])
                CK(cudaMemcpyPeerAsync(m->d_ident + lo[k], m->device, p->d_ident, p->device,
                                       (hi[k] - lo[k]) * sizeof(float), m->stream));
        }
        CK(cudaEventRecord(m->ev[5], m->stream));
        CK(cudaStreamSynchronize(m->stream));
        m->timings.comm_ms = ev_ms(m->ev[4], m->ev[5]);
        m->ident_full = (m->nk == m->nseq);
    }
    for (tcu_msa *p : m->peers) {
        m->timings.kernel_ms = std::max(m->timings.kernel_ms, p->timings.kernel_ms);
        m->timings.d2h_ms = std::max(m->timings.d2h_ms, p->timings.d2h_ms);
        m->timings.kernel_launches += p->timings.kernel_launches;
        // the bands of the further devices are not needed there any more
        dev_cache_give(p->device, p->d_ident, p->ident_cap);
        p->d_ident = nullptr;
        p->ident_cap = 0;
    }
    if (!keep_on_device) {
        dev_cache_give(m->device, m->d_ident, m->ident_cap);
        m->d_ident = nullptr;
        m->ident_cap = 0;
        m->ident_full = false;
    }
    return TCU_OK;
}

extern "C" int tcu_identity(tcu_msa *m, const int *save_seq, const int *save_res, uint8_t indet,
                            float *identities, int *hit_out, int *dst_out, int keep_on_device)
{
    NvtxRange nvtx("tcu_identity");
    if (m && !m->peers.empty() && !hit_out && !dst_out && !bytewise_alphabet(m, indet))
        return identity_multi(m, save_seq, save_res, indet, identities, keep_on_device);
    return identity_host(m, save_seq, save_res, indet, identities, hit_out, dst_out,
                         keep_on_device, false);
}

// Every rank computes the band of the pair matrix tcu_shard_blocks() assigns it,
// straight into its place in a full-size device array; the bands are then exchanged so
// that each GPU holds the whole matrix (what tcu_similarity_all needs).
extern "C" int tcu_identity_all(tcu_msa *m, tcu_comm *comm, const int *save_seq,
                                const int *save_res, uint8_t indet, float *identities)
{
    if (!m || !comm) return fail(TCU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    int rc = comm_check(m, comm);
    if (rc != TCU_OK) return rc;
    m->timings = tcu_timings{};
    m->ident_full = false;
    rc = tcu_identity_prepare(m, save_seq, save_res, indet, nullptr);
    if (rc != TCU_OK) return rc;
    const size_t npairs = (size_t)m->nk * (size_t)std::max(m->nk - 1, 0) / 2;
    if (npairs == 0) {
        CK(cudaStreamSynchronize(m->stream));
        return TCU_OK;
    }
    rc = ensure_ident(m, npairs * sizeof(float));
    if (rc != TCU_OK) return rc;
    std::vector<size_t> off(comm->world), cnt(comm->world);
    int my_b0 = 0, my_b1 = 0;
    for (int r = 0; r < comm->world; r++) {
        int b0, b1;
        tcu_shard_blocks(m->nk, r, comm->world, &b0, &b1);
        const size_t lo = tcu_identity_row_offset(m->nk, b0 * IB);
        const size_t hi = tcu_identity_row_offset(m->nk, std::min(b1 * IB, m->nk));
        off[r] = lo * sizeof(float);
        cnt[r] = (hi - lo) * sizeof(float);
        if (r == comm->rank) {
            my_b0 = b0;
            my_b1 = b1;
        }
    }
    rc = identity_launch(m, my_b0, my_b1, m->d_ident + off[comm->rank] / sizeof(float), nullptr,
                         nullptr);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[3], m->stream));
    rc = comm_allgatherv(comm, m->d_ident, off.data(), cnt.data(), m->stream);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[4], m->stream));
    if (identities) rc = download(m, identities, m->d_ident, npairs * sizeof(float));
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[5], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.comm_ms = ev_ms(m->ev[3], m->ev[4]);
    m->timings.d2h_ms = ev_ms(m->ev[4], m->ev[5]);
    m->ident_full = (m->nk == m->nseq);
    return TCU_OK;
}

// ---------------------------------------------------------------------------
// consumers of the device-resident identity matrix (clusters.cu)
// ---------------------------------------------------------------------------
extern "C" int tcu_identity_resident(const tcu_msa *m)
{
    return m && m->ident_full && (m->d_ident || m->nseq < 2) ? 1 : 0;
}

static int need_resident(tcu_msa *m)
{
    if (!m) return fail(TCU_ERR_INVALID, "msa is NULL");
    if (!tcu_identity_resident(m))
        return fail(TCU_ERR_STATE,
                    "no device-resident identity matrix: call tcu_identity with keep_on_device=1 "
                    "and all rows kept first");
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    return TCU_OK;
}

extern "C" int tcu_identity_download(tcu_msa *m, float *identities)
{
    int rc = need_resident(m);
    if (rc != TCU_OK) return rc;
    if (!identities) return fail(TCU_ERR_INVALID, "identities is NULL");
    const size_t npairs = (size_t)m->nseq * (size_t)std::max(m->nseq - 1, 0) / 2;
    if (npairs == 0) return TCU_OK;
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaMemcpyAsync(identities, m->d_ident, npairs * sizeof(float), cudaMemcpyDeviceToHost,
                       m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.d2h_ms = ev_ms(m->ev[0], m->ev[1]);
    return TCU_OK;
}

extern "C" int tcu_identity_row_stats(tcu_msa *m, int upper_only, float *row_max, float *row_min,
                                      float *row_sum)
{
    int rc = need_resident(m);
    if (rc != TCU_OK) return rc;
    const int n = m->nseq;
    if (n == 0) return TCU_OK;
    const size_t vec = ((size_t)n * sizeof(float) + 255) / 256 * 256;
    rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, 3 * vec, m);
    if (rc != TCU_OK) return rc;
    float *d_max = (float *)m->d_scratch;
    float *d_min = (float *)((uint8_t *)m->d_scratch + vec);
    float *d_sum = (float *)((uint8_t *)m->d_scratch + 2 * vec);
    CK(cudaEventRecord(m->ev[1], m->stream));
    // n == 1: no pairs, the kernel only writes the neutral values (0, 1, 0)
    CK(launch_row_stats(m->d_ident, n, upper_only != 0, d_max, d_min, d_sum, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    if (row_max) CK(cudaMemcpyAsync(row_max, d_max, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    if (row_min) CK(cudaMemcpyAsync(row_min, d_min, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    if (row_sum) CK(cudaMemcpyAsync(row_sum, d_sum, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

// bytes of the threshold bit matrix of n sequences (slab layout, tcu_internal.cuh)
static size_t bits_bytes(int n) { return std::max<size_t>(bits_total_words(n), 4) * sizeof(uint32_t); }
// bytes of the row-per-sequence form the clustering walk reads
static size_t brows_bytes(int n) { return std::max<size_t>(brow_total_words(n), 4) * sizeof(uint32_t); }

// All-gather of the slabs the ranks' bands own (band [b0, b1) of 128-row blocks = slabs
// [b0, b1), contiguous in the slab layout) with NCCL point-to-point transfers: the path taken
// when the ranks cannot map each other's memory.
static int bits_allgather(tcu_msa *m, tcu_comm *comm, cudaStream_t stream)
{
    std::vector<size_t> off(comm->world), cnt(comm->world);
    const size_t slab_b = bits_slab_words(m->nk) * sizeof(uint32_t);
    for (int r = 0; r < comm->world; r++) {
        int b0, b1;
        tcu_shard_blocks(m->nk, r, comm->world, &b0, &b1);
        off[r] = (size_t)b0 * slab_b;
        cnt[r] = (size_t)(b1 - b0) * slab_b;
    }
    return comm_allgatherv(comm, m->d_bits, off.data(), cnt.data(), stream);
}

// Greedy clustering in the given order over the threshold bit matrix m->d_brows (one row per
// sequence).  When `id0` is given the matrix is first derived from the resident float
// identities (K5); otherwise the caller has filled it (K1's threshold epilogue + the mirror /
// relayout pass).
static int clusters_impl(tcu_msa *m, const float *id0, const int *order, int count,
                         float threshold, int *clusters, int *n_clusters)
{
    int rc = TCU_OK;
    if (!n_clusters || (count > 0 && !order)) return fail(TCU_ERR_INVALID, "NULL argument");
    const int n = m->nseq;
    if (count < 0 || count > n) return fail(TCU_ERR_INVALID, "count %d outside [0,%d]", count, n);
    for (int k = 0; k < count; k++)
        if (order[k] < 0 || order[k] >= n)
            return fail(TCU_ERR_INVALID, "order[%d] = %d outside [0,%d)", k, order[k], n);
    *n_clusters = 0;
    if (count == 0) return TCU_OK;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t ord_b = up((size_t)count * 4), walk_b = greedy_scratch_bytes(n, count);
    if (id0) {
        rc = ensure_dev(m->device, (void **)&m->d_brows, &m->brows_cap, brows_bytes(n), m);
        if (rc != TCU_OK) return rc;
    }
    rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, 256 + 2 * ord_b + walk_b, m);
    if (rc != TCU_OK) return rc;
    uint8_t *p = (uint8_t *)m->d_scratch;
    int *d_count = (int *)p;
    p += 256;
    int *d_order = (int *)p;
    p += ord_b;
    int *d_clusters = (int *)p;
    p += ord_b;
    void *d_walk = p;
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaMemcpyAsync(d_order, order, (size_t)count * 4, cudaMemcpyHostToDevice, m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    if (id0) CK(launch_identity_bits(id0, n, threshold, m->d_brows, 0, n, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(launch_greedy_clusters(m->d_brows, n, d_order, count, d_walk, d_clusters, d_count, m->num_sms,
                              m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    int found = 0;
    void *stage = (size_t)(count + 1) * sizeof(int) <= STAGE_BYTES ? stage_acquire() : nullptr;
    if (stage) {
        // one round trip: the counter and the whole candidate list into page-locked memory
        int *h = (int *)stage;
        cudaError_t e = cudaMemcpyAsync(h, d_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream);
        if (e == cudaSuccess && clusters)
            e = cudaMemcpyAsync(h + 1, d_clusters, (size_t)count * 4, cudaMemcpyDeviceToHost, m->stream);
        if (e == cudaSuccess) e = cudaEventRecord(m->ev[4], m->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
        if (e == cudaSuccess) {
            found = h[0];
            if (clusters && found > 0) memcpy(clusters, h + 1, (size_t)found * 4);
        }
        stage_release(stage);
        if (e != cudaSuccess) return cuda_fail(e, "cluster list download");
    } else {
        CK(cudaMemcpyAsync(&found, d_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        if (clusters && found > 0)
            CK(cudaMemcpyAsync(clusters, d_clusters, (size_t)found * 4, cudaMemcpyDeviceToHost,
                               m->stream));
        CK(cudaEventRecord(m->ev[4], m->stream));
        CK(cudaStreamSynchronize(m->stream));
    }
    *n_clusters = found;
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);    // threshold -> bit matrix (K5)
    m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);  // greedy clustering
    m->timings.d2h_ms = ev_ms(m->ev[3], m->ev[4]);
    m->timings.kernel_launches = (id0 ? 1 : 0) + 1;
    return TCU_OK;
}

extern "C" int tcu_identity_clusters(tcu_msa *m, const int *order, int count, float threshold,
                                     int *clusters, int *n_clusters)
{
    int rc = need_resident(m);
    if (rc != TCU_OK) return rc;
    return clusters_impl(m, m->d_ident, order, count, threshold, clusters, n_clusters);
}

extern "C" int tcu_byte_histogram(tcu_msa *m, unsigned long long *hist256)
{
    if (!m || !hist256) return fail(TCU_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    memset(hist256, 0, 256 * sizeof(unsigned long long));
    if (m->nseq == 0 || m->ncol == 0) return TCU_OK;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, 256 * sizeof(unsigned long long), m);
    if (rc != TCU_OK) return rc;
    unsigned long long *d_hist = (unsigned long long *)m->d_scratch;
    CK(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_byte_histogram(m->d_raw, m->nseq, m->ncol, m->pitch, d_hist, m->num_sms, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(hist256, d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

extern "C" int tcu_sequence_lengths(tcu_msa *m, int *lengths)
{
    if (!m || !lengths) return fail(TCU_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    if (m->nseq == 0) return TCU_OK;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, (size_t)m->nseq * sizeof(int), m);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_row_lengths(m->d_raw, m->nseq, m->ncol, m->pitch, (int *)m->d_scratch, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(lengths, m->d_scratch, (size_t)m->nseq * sizeof(int), cudaMemcpyDeviceToHost,
                       m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

// Per row, the number of non-gap bytes over the kept columns (save_res; NULL = all columns,
// i.e. tcu_sequence_lengths).  With tcu_gaps(save_seq) for the columns this is everything
// Cleaner::removeAllGapsSeqsAndCols (Cleaner.cpp:1331-1396) scans the alignment for.
extern "C" int tcu_row_residues(tcu_msa *m, const int *save_res, int *residues)
{
    NvtxRange nvtx("tcu_row_residues");
    if (!m || !residues) return fail(TCU_ERR_INVALID, "NULL argument");
    if (!save_res) return tcu_sequence_lengths(m, residues);
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    const int n = m->nseq, L = m->ncol;
    if (n == 0) return TCU_OK;
    const size_t out_b = ((size_t)n * sizeof(int) + 255) / 256 * 256;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, out_b + m->pitch, m);
    if (rc != TCU_OK) return rc;
    std::vector<uint8_t> keep(m->pitch, 0);
    for (int k = 0; k < L; k++) keep[k] = save_res[k] != -1;
    uint8_t *d_keep = (uint8_t *)m->d_scratch + out_b;
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaMemcpyAsync(d_keep, keep.data(), m->pitch, cudaMemcpyHostToDevice, m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_row_residues(m->d_raw, n, m->pitch, d_keep, (int *)m->d_scratch, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(residues, m->d_scratch, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost,
                       m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

// Two 64-bit hashes per row (hashes[2 i], hashes[2 i + 1]); equal rows have equal hashes.
// Candidates for Cleaner::removeDuplicates (Cleaner.cpp:1489-1509), which the caller then
// confirms byte for byte.
extern "C" int tcu_row_hashes(tcu_msa *m, unsigned long long *hashes)
{
    NvtxRange nvtx("tcu_row_hashes");
    if (!m || !hashes) return fail(TCU_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    const int n = m->nseq;
    if (n == 0) return TCU_OK;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, (size_t)n * 16, m);
    if (rc != TCU_OK) return rc;
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_row_hashes(m->d_raw, n, m->pitch, (unsigned long long *)m->d_scratch, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(hashes, m->d_scratch, (size_t)n * 16, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = 1;
    return TCU_OK;
}

// Host only.  The visiting order of the two clustering walks: records (length, index)
// sorted ascending by the reference's own quicksort -- pivot = last element, compared as a
// float, Hoare-style scans that stop on equal keys, not stable -- and then walked from the
// end (Cleaner.cpp:1413-1426, utils.cpp:246-273).  The permutation a non-stable sort leaves
// among equal lengths decides which sequence represents a cluster, so the partition steps
// are replayed exactly; recursion is replaced by an explicit stack (sorted inputs make the
// reference recurse n deep).
namespace {
struct LenRec {
    int key, idx;
};

// The reference's partition scheme on v[ini..fin] with an explicit stack (sorted inputs make
// the reference recurse n deep).  The two halves a partition leaves are independent of
// each other, so a large left half may be handed to another thread (`spawn` = how many more
// threads this call may start): same swaps, same result, a fraction of the wall time -- the
// sort sits on the critical path of tcu_representatives once the identity kernel is
// spread over several GPUs.
void replay_quicksort(LenRec *v, int ini0, int fin0, int spawn)
{
    std::vector<std::pair<int, int>> todo;
    std::vector<std::thread> helpers;
    todo.emplace_back(ini0, fin0);
    while (!todo.empty()) {
        const int ini = todo.back().first, fin = todo.back().second;
        todo.pop_back();
        if (ini >= fin || fin < 0) continue;
        const float pivot = (float)v[fin].key;
        int i = ini - 1, j = fin;
        for (;;) {
            while ((float)v[++i].key < pivot)
                if (i == fin) break;
            while ((float)v[--j].key > pivot)
                if (j == 0) break;
            if (i < j)
                std::swap(v[i], v[j]);
            else
                break;
        }
        std::swap(v[i], v[fin]);
        bool handed = false;
        if (spawn > 0 && i - 1 - ini >= 8192 && fin - (i + 1) >= 8192) {
            try {
                helpers.emplace_back(replay_quicksort, v, ini, i - 1, spawn / 2);
                spawn /= 2;
                handed = true;
            } catch (...) {  // no thread to be had: do it here
            }
        }
        todo.emplace_back(i + 1, fin);
        if (!handed) todo.emplace_back(ini, i - 1);
    }
    for (auto &t : helpers) t.join();
}
}  // namespace

extern "C" int tcu_cluster_order(const int *lengths, int nseq, int *order)
{
    if (nseq < 0 || (nseq > 0 && (!lengths || !order))) return fail(TCU_ERR_INVALID, "bad argument");
    std::vector<LenRec> v;
    try {
        v.resize((size_t)nseq);
    } catch (...) {
        return fail(TCU_ERR_OOM, "host allocation failed");
    }
    for (int i = 0; i < nseq; i++) v[i] = LenRec{lengths[i], i};
    try {
        replay_quicksort(v.data(), 0, nseq - 1, /*spawn=*/4);
    } catch (...) {
        return fail(TCU_ERR_OOM, "host allocation failed");
    }
    for (int i = 0; i < nseq; i++) order[i] = v[nseq - 1 - i].idx;
    return TCU_OK;
}

// The sequence lengths and (if not known yet) the set of byte values that occur, in ONE
// round trip to the device: both are needed before the identity kernel can be set up.
static int lengths_and_presence(tcu_msa *m, int *lengths)
{
    CK(cudaSetDevice(m->device));
    m->timings = tcu_timings{};
    const int n = m->nseq;
    const size_t len_b = ((size_t)n * sizeof(int) + 255) / 256 * 256;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, len_b + 256 * sizeof(unsigned int), m);
    if (rc != TCU_OK) return rc;
    unsigned int *d_present = (unsigned int *)((uint8_t *)m->d_scratch + len_b);
    const bool presence = !m->have_present;
    CK(cudaEventRecord(m->ev[1], m->stream));
    if (presence) {
        CK(cudaMemsetAsync(d_present, 0, 256 * sizeof(unsigned int), m->stream));
        CK(launch_byte_presence(m->d_raw, m->nseq, m->ncol, m->pitch, d_present, m->num_sms, m->stream));
    }
    CK(launch_row_lengths(m->d_raw, n, m->ncol, m->pitch, (int *)m->d_scratch, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    if (presence)
        CK(cudaMemcpyAsync(m->present, d_present, sizeof m->present, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaMemcpyAsync(lengths, m->d_scratch, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (presence) m->have_present = true;
    m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[2]);
    m->timings.d2h_ms = ev_ms(m->ev[2], m->ev[3]);
    m->timings.kernel_launches = presence ? 2 : 1;
    return TCU_OK;
}

extern "C" void tcu_threshold_rule(float threshold, int *mode, unsigned *mul, int *shift)
{
    const ThresholdRule r = threshold_rule(threshold);
    if (mode) *mode = r.mode;
    if (mul) *mul = r.mul;
    if (shift) *shift = r.shift;
}

// Cleaner::calculateRepresentativeSeq in one call.  The walk only ever asks "identity >
// threshold" (Cleaner.cpp:1435-1440), so the float matrix is never materialised: K1 compares
// each ratio with the threshold in its epilogue and emits one bit per pair (n^2/16 bytes
// instead of 4*P: 156 MB instead of 5 GB at 50 000 sequences, and no capacity ceiling from
// the packed array), a mirror pass completes the symmetric matrix, and the greedy
// clustering runs on it.  The sequence lengths go first (one tiny kernel) so that the
// host-side replay of the reference's sort runs on another thread under the identity kernel.
// With a communicator every rank computes the bits of its row band (a contiguous run of
// slabs), the bands are all-gathered over NVLink and every rank finishes alone.
static int representatives_impl(tcu_msa *m, tcu_comm *comm, const int *save_res, uint8_t indet,
                                float threshold, int *clusters, int *n_clusters)
{
    if (!m || !n_clusters) return fail(TCU_ERR_INVALID, "NULL argument");
    int rc = comm_check(m, comm);
    if (rc != TCU_OK) return rc;
    const int n = m->nseq;
    *n_clusters = 0;
    if (n == 0) return TCU_OK;
    tcu_timings total{};
    auto add = [&](const tcu_timings &t) {
        total.h2d_ms += t.h2d_ms;
        total.pack_ms += t.pack_ms;
        total.kernel_ms += t.kernel_ms;
        total.d2h_ms += t.d2h_ms;
        total.comm_ms += t.comm_ms;
        total.kernel_launches += t.kernel_launches;
    };
    static const bool trace = getenv("TCU_TRACE") != nullptr;
    const double t_begin = now_ms();
    // lengths and visiting order live in page-locked memory when they fit a staging buffer
    // (copies to and from it are asynchronous and run at PCIe speed)
    struct StageHold {
        void *p = nullptr;
        ~StageHold() { stage_release(p); }
    } hold;
    std::vector<int> pageable;
    int *lengths, *order;
    if ((size_t)n * 2 * sizeof(int) <= STAGE_BYTES && (hold.p = stage_acquire()) != nullptr) {
        lengths = (int *)hold.p;
        order = lengths + n;
    } else {
        try {
            pageable.resize((size_t)n * 2);
        } catch (...) {
            return fail(TCU_ERR_OOM, "host allocation failed");
        }
        lengths = pageable.data();
        order = lengths + n;
    }
    rc = lengths_and_presence(m, lengths);
    if (rc != TCU_OK) return rc;
    add(m->timings);
    const double t_lengths = now_ms();
    int sort_rc = TCU_OK;
    std::string sort_err;
    double sort_ms = 0;
    auto sort = [&]() {
        const double t0 = now_ms();
        sort_rc = tcu_cluster_order(lengths, n, order);
        if (sort_rc != TCU_OK) sort_err = g_last_error;  // thread-local: carry it over
        sort_ms = now_ms() - t0;
    };
    std::thread sorter;
    try {
        sorter = std::thread(sort);
    } catch (...) {  // no thread to be had: sort here, nothing may escape the C ABI
        sort();
    }
    auto device_part = [&]() -> int {
        m->timings = tcu_timings{};
        m->ident_full = false;
        int r = tcu_identity_prepare(m, nullptr, save_res, indet, nullptr);
        if (r != TCU_OK) return r;
        r = ensure_dev(m->device, (void **)&m->d_bits, &m->bits_cap, bits_bytes(n), m);
        if (r != TCU_OK) return r;
        r = ensure_dev(m->device, (void **)&m->d_brows, &m->brows_cap, brows_bytes(n), m);
        if (r != TCU_OK) return r;
        if (m->nchunks == 0) {
            // no columns: every identity is 0 (template.h:427-434)
            CK(cudaMemsetAsync(m->d_brows, threshold < 0.f ? 0xFF : 0, brows_bytes(n), m->stream));
            CK(cudaEventRecord(m->ev[3], m->stream));
        } else {
            int b0 = 0, b1 = m->nsb;
            if (comm) tcu_shard_blocks(m->nk, comm->rank, comm->world, &b0, &b1);
            // With peer memory (CUDA IPC over NVLink) the band exchange is part of K1: its
            // epilogue stores every column entry into all ranks' matrices, and one one-word
            // all-reduce afterwards tells every rank that all bands have landed.  (peer_prepare
            // ends with a collective the host waits for: every rank is inside this call, so
            // nobody still clusters on the matrix the stores go to.)
            std::vector<void *> peer_base;
            const bool use_peer = comm && comm->world - 1 <= ID2_MAX_PEERS &&
                                  peer_prepare(comm, m->d_bits, peer_base, m->stream);
            std::vector<uint32_t *> others;
            if (use_peer)
                for (int q = 1; q < comm->world; q++)  // staggered like the pulls were
                    others.push_back((uint32_t *)peer_base[(size_t)((comm->rank + q) % comm->world)]);
            r = identity_launch(m, b0, b1, nullptr, nullptr, nullptr, m->d_bits, threshold,
                                use_peer ? &others : nullptr);
            if (r != TCU_OK) return r;
            CK(cudaEventRecord(m->ev[3], m->stream));
            if (use_peer) {
                r = comm_allreduce_i32(comm, (int *)(comm->d_sync + 64 * 64 + 64), 1, m->stream);
                if (r != TCU_OK) return r;
            } else if (comm) {
                r = bits_allgather(m, comm, m->stream);
                if (r != TCU_OK) return r;
            }
            CK(cudaEventRecord(m->ev[4], m->stream));
            CK(launch_bits_rows(m->d_bits, n, m->d_brows, m->stream));
            m->timings.kernel_launches++;
        }
        CK(cudaEventRecord(m->ev[5], m->stream));
        CK(cudaStreamSynchronize(m->stream));
        m->timings.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
        m->timings.pack_ms = ev_ms(m->ev[1], m->ev[2]);
        m->timings.kernel_ms = ev_ms(m->ev[2], m->ev[3]);
        if (m->nchunks) {
            m->timings.comm_ms = ev_ms(m->ev[3], m->ev[4]);
            m->timings.pack_ms += ev_ms(m->ev[4], m->ev[5]);  // mirror pass
        }
        return TCU_OK;
    };
    // The replicas of a multi-device handle (one process, several GPUs): every device
    // thresholds its band.  With peer access the epilogue of K1 stores straight into the
    // first device's matrix over NVLink (nothing to exchange afterwards); otherwise a device
    // keeps its run of slabs and the first device pulls it (one 2-D peer copy each).  The
    // first device then mirrors and clusters.
    auto device_part_multi = [&]() -> int {
        const int world = 1 + (int)m->peers.size();
        const size_t slab_w = bits_slab_words(n);
        const bool direct = m->peer_direct;
        std::vector<int> bb((size_t)world + 1, 0);
        CK(cudaSetDevice(m->device));
        int r = ensure_dev(m->device, (void **)&m->d_bits, &m->bits_cap, bits_bytes(n), m);
        if (r != TCU_OK) return r;
        r = ensure_dev(m->device, (void **)&m->d_brows, &m->brows_cap, brows_bytes(n), m);
        if (r != TCU_OK) return r;
        r = for_each_replica(m, [&](tcu_msa *d, int k) -> int {
            d->timings = tcu_timings{};
            d->ident_full = false;
            int e = tcu_identity_prepare(d, nullptr, save_res, indet, nullptr);
            if (e != TCU_OK) return e;
            int b0 = 0, b1 = 0;
            tcu_shard_blocks(d->nk, k, world, &b0, &b1);
            bb[k] = b0;
            bb[k + 1] = b1;
            uint32_t *base = m->d_bits;
            if (k > 0 && !direct) {
                e = ensure_dev(d->device, (void **)&d->d_bits, &d->bits_cap,
                               std::max<size_t>((size_t)(b1 - b0) * slab_w, 4) * sizeof(uint32_t), d);
                if (e != TCU_OK) return e;
                // K1 indexes slabs absolutely: a device that holds only its band passes the
                // address slab 0 would have
                base = d->d_bits - (size_t)b0 * slab_w;
            }
            e = identity_launch(d, b0, b1, nullptr, nullptr, nullptr, base, threshold);
            if (e != TCU_OK) return e;
            CK(cudaEventRecord(d->ev[3], d->stream));
            CK(cudaStreamSynchronize(d->stream));
            d->timings.h2d_ms = ev_ms(d->ev[0], d->ev[1]);
            d->timings.pack_ms = ev_ms(d->ev[1], d->ev[2]);
            d->timings.kernel_ms = ev_ms(d->ev[2], d->ev[3]);
            return TCU_OK;
        });
        if (r != TCU_OK) return r;
        CK(cudaSetDevice(m->device));
        CK(cudaEventRecord(m->ev[3], m->stream));
        for (int k = 1; k < world; k++) {
            tcu_msa *p = m->peers[(size_t)k - 1];
            if (!direct && bb[k + 1] > bb[k]) {
                // rows below the band's first row carry nothing before the mirror pass
                const size_t first_row = std::min<size_t>((size_t)bb[k] * IB, (size_t)n);
                CK(cudaMemcpy2DAsync(m->d_bits + (size_t)bb[k] * slab_w + first_row * 4, slab_w * 4,
                                     p->d_bits + first_row * 4, slab_w * 4, ((size_t)n - first_row) * 16,
                                     (size_t)(bb[k + 1] - bb[k]), cudaMemcpyDeviceToDevice, m->stream));
            }
            m->timings.kernel_ms = std::max(m->timings.kernel_ms, p->timings.kernel_ms);
            m->timings.kernel_launches += p->timings.kernel_launches;
        }
        CK(cudaEventRecord(m->ev[4], m->stream));
        CK(launch_bits_rows(m->d_bits, n, m->d_brows, m->stream));
        m->timings.kernel_launches++;
        CK(cudaEventRecord(m->ev[5], m->stream));
        CK(cudaStreamSynchronize(m->stream));
        m->timings.comm_ms = ev_ms(m->ev[3], m->ev[4]);
        m->timings.pack_ms += ev_ms(m->ev[4], m->ev[5]);  // mirror pass
        return TCU_OK;
    };
    const bool bytewise = bytewise_alphabet(m, indet);
    if (bytewise && comm) {
        if (sorter.joinable()) sorter.join();
        return fail(TCU_ERR_INVALID, "more than %d distinct symbols: not supported across ranks",
                    (1 << MAX_PLANES) - 2);
    }
    if (bytewise) {
        // no bit-plane operand: float matrix from the byte-wise kernel, thresholded by K5
        rc = tcu_identity(m, nullptr, save_res, indet, nullptr, nullptr, nullptr, 1);
        if (sorter.joinable()) sorter.join();
        if (rc != TCU_OK) return rc;
        add(m->timings);
        if (sort_rc != TCU_OK) return fail(sort_rc, "%s", sort_err.c_str());
        rc = clusters_impl(m, m->d_ident, order, n, threshold, clusters, n_clusters);
        if (rc != TCU_OK) return rc;
        add(m->timings);
        m->timings = total;
        return TCU_OK;
    }
    rc = (!comm && !m->peers.empty() && m->ncol > 0) ? device_part_multi() : device_part();
    const double t_device = now_ms();
    if (sorter.joinable()) sorter.join();
    const double t_join = now_ms();
    if (rc != TCU_OK) return rc;
    add(m->timings);
    if (sort_rc != TCU_OK) return fail(sort_rc, "%s", sort_err.c_str());
    rc = clusters_impl(m, nullptr, order, n, threshold, clusters, n_clusters);
    if (rc != TCU_OK) return rc;
    add(m->timings);
    if (trace)
        fprintf(stderr, "[tcu] representatives (host clock, ms): lengths %.3f, identity+exchange %.3f "
                "(sort on its thread %.3f, waited for %.3f), clustering %.3f; device: pack %.3f K1 %.3f "
                "comm %.3f greedy %.3f\n", t_lengths - t_begin, t_device - t_lengths, sort_ms,
                t_join - t_device, now_ms() - t_join, total.pack_ms, total.kernel_ms - m->timings.kernel_ms,
                total.comm_ms, m->timings.kernel_ms);
    m->timings = total;
    return TCU_OK;
}

extern "C" int tcu_representatives(tcu_msa *m, const int *save_res, uint8_t indet, float threshold,
                                   int *clusters, int *n_clusters)
{
    NvtxRange nvtx("tcu_representatives");
    return representatives_impl(m, nullptr, save_res, indet, threshold, clusters, n_clusters);
}

extern "C" int tcu_representatives_all(tcu_msa *m, tcu_comm *comm, const int *save_res,
                                       uint8_t indet, float threshold, int *clusters,
                                       int *n_clusters)
{
    if (!comm) return fail(TCU_ERR_INVALID, "comm is NULL");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    return representatives_impl(m, comm, save_res, indet, threshold, clusters, n_clusters);
}

// ---------------------------------------------------------------------------
// K2 spurious
// ---------------------------------------------------------------------------
// How one device takes part in the statistic when the rows are spread over the replicas of a
// multi-device handle: phase 1 = partial column counts of the share (rank of world) to
// host_counts (2 * ncol ints: '-' then indet); phase 2 = total counts from host_counts, then
// the share's rows of the vector straight into the caller's array.  Phase 0 = all of it.
struct SpuriousShare {
    int rank = 0, world = 1, phase = 0;
    int *host_counts = nullptr;
};

static int spurious_impl(tcu_msa *m, tcu_comm *comm, const SpuriousShare &sh, uint8_t indet,
                         uint32_t ovrlap, float *spurious)
{
    if (!m || !spurious) return fail(TCU_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(m->device));
    int rc0 = comm_check(m, comm);
    if (rc0 != TCU_OK) return rc0;
    if (sh.phase != 2) m->timings = tcu_timings{};
    const int n = m->nseq, L = m->ncol;
    if (n == 0) return TCU_OK;
    if (L == 0) {
        // 0/0 in the reference's final division (template.h:309)
        if (sh.phase != 1 && sh.rank == 0)
            for (int i = 0; i < n; i++) spurious[i] = NAN;
        return TCU_OK;
    }
    const size_t cnt_bytes = ((size_t)L * sizeof(int) + 255) / 256 * 256;
    const size_t out_bytes = ((size_t)n * sizeof(float) + 255) / 256 * 256;
    const size_t flag_bytes = (3 * (m->pitch >> 5) * sizeof(uint32_t) + 255) / 256 * 256;
    const size_t plane_bytes = (size_t)n * (m->pitch >> 3);  // one bit per cell, pitch % 128 == 0
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap,
                        2 * cnt_bytes + out_bytes + flag_bytes + 2 * plane_bytes, m);
    if (rc != TCU_OK) return rc;
    int *d_cg = (int *)m->d_scratch;
    int *d_cx = (int *)((uint8_t *)m->d_scratch + cnt_bytes);
    float *d_out = (float *)((uint8_t *)m->d_scratch + 2 * cnt_bytes);
    uint32_t *d_flags = (uint32_t *)((uint8_t *)m->d_scratch + 2 * cnt_bytes + out_bytes);
    uint8_t *d_pg = (uint8_t *)d_flags + flag_bytes, *d_px = d_pg + plane_bytes;
    // several ranks: partial column counts over this rank's rows, summed across GPUs
    // (d_cg and d_cx are adjacent: one all-reduce), then this rank's rows of the
    // vector, all-gathered
    int r0 = 0, r1 = n;
    if (comm) tcu_shard_range(n, 1, comm->rank, comm->world, &r0, &r1);
    else tcu_shard_range(n, 1, sh.rank, sh.world, &r0, &r1);
    const size_t prow = m->pitch >> 3;  // bytes per plane row
    if (sh.phase != 2) {
        CK(cudaMemsetAsync(d_cg, 0, 2 * cnt_bytes, m->stream));
        CK(cudaEventRecord(m->ev[1], m->stream));
        CK(launch_column_counts(m->d_raw + (size_t)r0 * m->pitch, r1 - r0, L, m->pitch, nullptr,
                                '-', indet, d_cg, d_cx, (uint16_t *)(d_pg + (size_t)r0 * prow),
                                (uint16_t *)(d_px + (size_t)r0 * prow), m->num_sms, m->stream));
        CK(cudaEventRecord(m->ev[4], m->stream));
        if (comm) {
            rc = comm_allreduce_i32(comm, d_cg, 2 * cnt_bytes / sizeof(int), m->stream);
            if (rc != TCU_OK) return rc;
        }
        if (sh.phase == 1) {
            CK(cudaMemcpyAsync(sh.host_counts, d_cg, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost,
                               m->stream));
            CK(cudaMemcpyAsync(sh.host_counts + L, d_cx, (size_t)L * sizeof(int),
                               cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
            m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[4]);
            m->timings.kernel_launches = 1;
            return TCU_OK;
        }
    } else {
        CK(cudaMemcpyAsync(d_cg, sh.host_counts, (size_t)L * sizeof(int), cudaMemcpyHostToDevice,
                           m->stream));
        CK(cudaMemcpyAsync(d_cx, sh.host_counts + L, (size_t)L * sizeof(int), cudaMemcpyHostToDevice,
                           m->stream));
    }
    CK(cudaEventRecord(m->ev[5], m->stream));
    CK(launch_spurious_rows((const uint32_t *)d_pg, (const uint32_t *)d_px, n, r0, r1, L, m->pitch,
                            d_cg, d_cx, ovrlap, d_flags, d_out, m->num_sms, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    if (comm) {
        std::vector<size_t> off(comm->world), cnt(comm->world);
        for (int r = 0; r < comm->world; r++) {
            int a, b;
            tcu_shard_range(n, 1, r, comm->world, &a, &b);
            off[r] = (size_t)a * sizeof(float);
            cnt[r] = (size_t)(b - a) * sizeof(float);
        }
        rc = comm_allgatherv(comm, d_out, off.data(), cnt.data(), m->stream);
        if (rc != TCU_OK) return rc;
    }
    CK(cudaEventRecord(m->ev[0], m->stream));
    if (sh.phase == 2) {
        if (r1 > r0)
            CK(cudaMemcpyAsync(spurious + r0, d_out + r0, (size_t)(r1 - r0) * sizeof(float),
                               cudaMemcpyDeviceToHost, m->stream));
    } else {
        CK(cudaMemcpyAsync(spurious, d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost,
                           m->stream));
    }
    CK(cudaEventRecord(m->ev[3], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (sh.phase == 2) {
        m->timings.kernel_ms += ev_ms(m->ev[5], m->ev[2]);
        m->timings.kernel_launches += 2;
    } else {
        m->timings.kernel_ms = ev_ms(m->ev[1], m->ev[4]) + ev_ms(m->ev[5], m->ev[2]);
        m->timings.comm_ms = ev_ms(m->ev[4], m->ev[5]) + ev_ms(m->ev[2], m->ev[0]);
        m->timings.kernel_launches = 3;
    }
    m->timings.d2h_ms = ev_ms(m->ev[0], m->ev[3]);
    return TCU_OK;
}

extern "C" int tcu_spurious(tcu_msa *m, uint8_t indet, uint32_t ovrlap, float *spurious)
{
    NvtxRange nvtx("tcu_spurious");
    if (!m || !spurious) return fail(TCU_ERR_INVALID, "NULL argument");
    if (m->peers.empty()) return spurious_impl(m, nullptr, SpuriousShare{}, indet, ovrlap, spurious);
    // several devices: partial column counts per share of the rows, summed on the host
    // (integers), then every device finishes its own rows of the vector
    const int world = 1 + (int)m->peers.size(), L = m->ncol;
    std::vector<std::vector<int>> part((size_t)world, std::vector<int>((size_t)std::max(2 * L, 1), 0));
    int rc = for_each_replica(m, [&](tcu_msa *r, int k) {
        SpuriousShare sh;
        sh.rank = k, sh.world = world, sh.phase = 1, sh.host_counts = part[k].data();
        return spurious_impl(r, nullptr, sh, indet, ovrlap, spurious);
    });
    if (rc != TCU_OK) return rc;
    for (int d = 1; d < world; d++)
        for (int k = 0; k < 2 * L; k++) part[0][k] += part[d][k];
    rc = for_each_replica(m, [&](tcu_msa *r, int k) {
        SpuriousShare sh;
        sh.rank = k, sh.world = world, sh.phase = 2, sh.host_counts = part[0].data();
        return spurious_impl(r, nullptr, sh, indet, ovrlap, spurious);
    });
    if (rc != TCU_OK) return rc;
    for (tcu_msa *p : m->peers) m->timings.kernel_ms = std::max(m->timings.kernel_ms, p->timings.kernel_ms);
    return TCU_OK;
}

extern "C" int tcu_spurious_all(tcu_msa *m, tcu_comm *comm, uint8_t indet, uint32_t ovrlap,
                                float *spurious)
{
    if (!comm) return fail(TCU_ERR_INVALID, "comm is NULL");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    return spurious_impl(m, comm, SpuriousShare{}, indet, ovrlap, spurious);
}

// ---------------------------------------------------------------------------
// K4 similarity
// ---------------------------------------------------------------------------
static int similarity_impl(tcu_msa *m, tcu_comm *comm, uint8_t indet, const float *dist, int npos,
                           const int *vhash, const int *gaps, float gap_threshold,
                           const float *identities, float *num, float *den, float *mdk,
                           int *err_col, int *err_row, int *err_byte)
{
    if (!m || !dist || !vhash || !num || !den) return fail(TCU_ERR_INVALID, "NULL argument");
    if (npos < 1 || npos > SIM_MAX_POS)
        return fail(TCU_ERR_INVALID, "similarity matrix order %d outside [1,%d]", npos, SIM_MAX_POS);
    CK(cudaSetDevice(m->device));
    int rc0 = comm_check(m, comm);
    if (rc0 != TCU_OK) return rc0;
    const int n = m->nseq, L = m->ncol;
    const size_t npairs = (size_t)n * (size_t)std::max(n - 1, 0) / 2;
    tcu_timings t{};
    if (L == 0) return TCU_OK;

    if (identities) {
        int rc = ensure_ident(m, std::max<size_t>(npairs, 1) * sizeof(float));
        if (rc != TCU_OK) return rc;
        CK(cudaMemcpyAsync(m->d_ident, identities, npairs * sizeof(float), cudaMemcpyHostToDevice,
                           m->stream));
        m->ident_full = true;
    } else if (npairs && (!m->d_ident || !m->ident_full)) {
        return fail(TCU_ERR_STATE,
                    "tcu_similarity needs the identities of the unmasked alignment: call "
                    "tcu_identity(..., keep_on_device=1) first or pass them");
    }

    // byte -> code table (template.h:130-147) and skipped columns (:122-125)
    uint8_t lut[256];
    for (int b = 0; b < 256; b++) {
        int up = (b >= 'a' && b <= 'z') ? b - 32 : b;
        if (up == indet || up == '-') lut[b] = SIM_GAP;
        else if (up < 'A' || up > 'Z') lut[b] = SIM_INCORRECT;
        else if (vhash[up - 'A'] < 0 || vhash[up - 'A'] >= npos) lut[b] = SIM_UNDEFINED;
        else lut[b] = (uint8_t)vhash[up - 'A'];
    }
    std::vector<uint8_t> skip((size_t)m->pitch, 0);
    if (gaps)
        for (int k = 0; k < L; k++) skip[k] = (float)gaps[k] >= gap_threshold;

    const int npad = (n + 31) / 32 * 32;
    const int ngroups = (int)(m->pitch >> 5);
    const size_t codes_bytes = (size_t)ngroups * npad * 32;
    const size_t vec_bytes = ((size_t)L * sizeof(float) + 255) / 256 * 256;
    const size_t dist_bytes = 4096;
    const size_t skip_bytes = ((size_t)ngroups * (npad >> 5) * sizeof(uint32_t) + 255) / 256 * 256;
    const size_t nb_bytes = ((size_t)ngroups * sizeof(unsigned long long) + 255) / 256 * 256;
    const size_t ngm_bytes = (size_t)ngroups * npad * sizeof(uint32_t);
    const size_t need = codes_bytes + 2 * vec_bytes + dist_bytes + m->pitch + 256 + 64 + skip_bytes +
                        nb_bytes + ngm_bytes;
    int rc = ensure_dev(m->device, &m->d_scratch, &m->scratch_cap, need, m);
    if (rc != TCU_OK) return rc;
    uint8_t *base = (uint8_t *)m->d_scratch;
    float *d_num = (float *)base;
    float *d_den = (float *)(base + vec_bytes);
    float *d_dist = (float *)(base + 2 * vec_bytes);
    uint8_t *d_skip = base + 2 * vec_bytes + dist_bytes;
    uint8_t *d_lut = d_skip + m->pitch;
    unsigned long long *d_err = (unsigned long long *)(d_lut + 256);
    uint32_t *d_rowskip = (uint32_t *)(d_err + 8);
    unsigned long long *d_nbatches = (unsigned long long *)((uint8_t *)d_rowskip + skip_bytes);
    uint8_t *d_codes = (uint8_t *)d_nbatches + nb_bytes;
    uint32_t *d_ngmask = (uint32_t *)(d_codes + codes_bytes);

    const unsigned long long no_err = ~0ull;
    unsigned long long first_err = no_err;
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaMemsetAsync(d_num, 0, 2 * vec_bytes, m->stream));
    CK(cudaMemcpyAsync(d_dist, dist, (size_t)npos * npos * sizeof(float), cudaMemcpyHostToDevice,
                       m->stream));
    CK(cudaMemcpyAsync(d_skip, skip.data(), m->pitch, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(d_err, &no_err, sizeof no_err, cudaMemcpyHostToDevice, m->stream));
    CK(cudaEventRecord(m->ev[1], m->stream));
    CK(launch_sim_codes(m->d_raw, n, L, m->pitch, npad, d_lut, d_skip, d_codes, d_err, m->num_sms, m->stream));
    CK(launch_sim_rows(d_codes, n, npad, ngroups, d_rowskip, d_nbatches, d_ngmask, m->stream));
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(cudaMemcpyAsync(&first_err, d_err, sizeof first_err, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    t.h2d_ms = ev_ms(m->ev[0], m->ev[1]);
    t.pack_ms = ev_ms(m->ev[1], m->ev[2]);
    t.kernel_launches = n > 0 ? 2 : 0;
    if (first_err != no_err) {
        const int byte = (int)(first_err & 0xFF);
        const unsigned long long cell = first_err >> 8;
        const int col = (int)(cell / (unsigned long long)n), row = (int)(cell % (unsigned long long)n);
        if (err_col) *err_col = col;
        if (err_row) *err_row = row;
        if (err_byte) *err_byte = byte;
        m->timings = t;
        const bool incorrect = byte < 'A' || byte > 'Z';
        return fail(incorrect ? TCU_ERR_INCORRECT_SYMBOL : TCU_ERR_UNDEFINED_SYMBOL,
                    "symbol '%c' at column %d, row %d cannot be scored", byte, col, row);
    }

    // several ranks: columns are independent given the identities, so each rank runs the
    // chains of its share of the 32-column groups and the two vectors are all-gathered
    // (d_num and d_den sit vec_bytes apart: two ranges per rank)
    const int col_groups = (L + 31) / 32;
    int g0 = 0, g1 = col_groups;
    if (comm) tcu_shard_range(col_groups, 1, comm->rank, comm->world, &g0, &g1);
    CK(cudaEventRecord(m->ev[2], m->stream));
    CK(launch_similarity(d_codes, n, npad, L, m->d_ident, d_dist, npos, d_skip, d_rowskip,
                         d_ngmask, d_nbatches, g0, g1, d_num, d_den, m->num_sms, m->stream));
    CK(cudaEventRecord(m->ev[3], m->stream));
    if (comm) {
        for (int v = 0; v < 2; v++) {
            std::vector<size_t> off(comm->world), cnt(comm->world);
            for (int r = 0; r < comm->world; r++) {
                int a, b;
                tcu_shard_range(col_groups, 1, r, comm->world, &a, &b);
                off[r] = (size_t)a * 32 * sizeof(float);
                cnt[r] = (size_t)(std::min(b * 32, L) - std::min(a * 32, L)) * sizeof(float);
            }
            rc = comm_allgatherv(comm, v ? (void *)d_den : (void *)d_num, off.data(), cnt.data(),
                                 m->stream);
            if (rc != TCU_OK) return rc;
        }
    }
    CK(cudaEventRecord(m->ev[5], m->stream));
    CK(cudaMemcpyAsync(num, d_num, (size_t)L * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaMemcpyAsync(den, d_den, (size_t)L * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaEventRecord(m->ev[4], m->stream));
    CK(cudaStreamSynchronize(m->stream));
    t.kernel_ms = ev_ms(m->ev[2], m->ev[3]);
    t.comm_ms = ev_ms(m->ev[3], m->ev[5]);
    t.d2h_ms = ev_ms(m->ev[5], m->ev[4]);
    t.kernel_launches += (n > 0 && g1 > g0) ? 1 : 0;
    m->timings = t;

    if (mdk) {
        // template.h:186-200, glibc expf on the host so the last bit matches
        for (int k = 0; k < L; k++) {
            if (den[k] == 0) mdk[k] = 0.0f;
            else {
                const float q = num[k] / den[k];
                mdk[k] = q < 0 ? 1.0f : expf(-q);
            }
        }
    }
    return TCU_OK;
}

extern "C" int tcu_similarity(tcu_msa *m, uint8_t indet, const float *dist, int npos,
                              const int *vhash, const int *gaps, float gap_threshold,
                              const float *identities, float *num, float *den, float *mdk,
                              int *err_col, int *err_row, int *err_byte)
{
    NvtxRange nvtx("tcu_similarity");
    return similarity_impl(m, nullptr, indet, dist, npos, vhash, gaps, gap_threshold, identities,
                           num, den, mdk, err_col, err_row, err_byte);
}

extern "C" int tcu_similarity_all(tcu_msa *m, tcu_comm *comm, uint8_t indet, const float *dist,
                                  int npos, const int *vhash, const int *gaps, float gap_threshold,
                                  float *num, float *den, float *mdk, int *err_col, int *err_row,
                                  int *err_byte)
{
    if (!comm) return fail(TCU_ERR_INVALID, "comm is NULL");
    std::lock_guard<std::mutex> comm_lock(comm->mutex);
    return similarity_impl(m, comm, indet, dist, npos, vhash, gaps, gap_threshold, nullptr, num,
                           den, mdk, err_col, err_row, err_byte);
}
