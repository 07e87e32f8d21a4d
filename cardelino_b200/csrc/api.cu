// C ABI of libcardelino_b200.so (include/cardelino_b200.h): handle, data loading, the Gibbs driver
// loop of clone_id_Gibbs (R/clone_id.R:351-675), clone_id_EM (:240-326) and get_logLik (:734-752).
#include "../../include/cardelino_b200.h"
#include "internal.h"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

using namespace cdl;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        if (p && n == count) return;       // same size: keep the buffer (contents undefined, as after cudaMalloc)
        release();
        n = count;
        const size_t bytes = sizeof(T) * (count ? count : 1);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(CDL_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " +
                                           cudaGetErrorString(e));
        }
    }
    void zero(cudaStream_t st) { CDL_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * (n ? n : 1), st)); }
    void upload(const T* h, size_t count, cudaStream_t st) {
        CDL_CUDA(cudaMemcpyAsync(p, h, sizeof(T) * count, cudaMemcpyHostToDevice, st));
    }
    void download(T* h, size_t count, cudaStream_t st, size_t offset = 0) const {
        CDL_CUDA(cudaMemcpyAsync(h, p + offset, sizeof(T) * count, cudaMemcpyDeviceToHost, st));
        CDL_CUDA(cudaStreamSynchronize(st));
    }
};

// ---- NCCL through dlopen: the host plumbing (torch.distributed) already has libnccl loaded; we bind the
// same copy instead of linking a second one. ----
struct Id128 { char b[128]; };   // ncclUniqueId is passed BY VALUE (128 bytes)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

struct Chain {
    ChainDev dev{};
    DevBuf<double> scal, l1, l1c, theta0_all, theta1_all, loglik_all, relax_all, prob_all, prob_acc, part_d;
    DevBuf<double> el_l1, el_l1c, sd1;
    DevBuf<double> prob_sq;              // running sum of squares beside prob_acc (streaming Geweke windows)
    DevBuf<uint32_t> mask;
    DevBuf<uint8_t> z, config_all, assign_all;
    DevBuf<unsigned long long> sab, part_u, sab_zero;
    DevBuf<unsigned int> ticket, work, sell_tickets, stats_work, chg_ctl;
    DevBuf<uint32_t> chg_list;
    DevBuf<uint8_t> chg_old;
    DevBuf<double> sell_scratch;
    DevBuf<double2> tab_g;
    DevBuf<unsigned long long> dbg_s34, dbg_sab;   // test hooks (cdl_debug_capture)
    DevBuf<unsigned long long> tl;                 // device timeline of the last sweep (CDL_TIMELINE=1)
    // summaries
    DevBuf<double> prob_mean, config_prob, theta1_mean;
    double theta0_mean = 0.0, relax_mean = 0.0;
    cdl_gibbs_info info{};
    bool done = false;
};

}  // namespace

struct cdl_handle_s {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    std::string err;
    Donor donor;
    bool has_data = false;
    // last Gibbs run
    std::vector<std::unique_ptr<Chain>> chains;
    // Chains of the PREVIOUS run, kept for their device buffers (retire_chains): a run of the same shape -- the steps of a
    // clone_id(n_chain = 64) loop on a small donor, a chain per parameter setting -- then allocates and frees nothing
    // (64 chains x 25 buffers, 4.4 GB of traces on example_donor at max_iter = 5000: 0.13 s of cudaMalloc / cudaFree per
    // call against 0.07 s of sweeps).  At most SPARE_BYTES are kept; buffers of another size are reallocated one by one.
    std::vector<std::unique_ptr<Chain>> spare;
    cdl_gibbs_params gp{};
    int K = 0;
    int64_t T = 0, n_el = 0;
    DevBuf<uint32_t> cfg0;
    DevBuf<double> prior1_mat;
    SweepArgs sweep{};        // template of the last run (bench hook)
    bool have_sweep = false;
    // comm
    NcclApi nccl;
    void* comm = nullptr;
    int rank = 0, world = 1;
    int64_t launches = 0;
    // statistic-table exchange of cell-sharded runs: peer memory (fused into k_table) or NCCL all-reduce
    int exchange_mode = CDL_EXCHANGE_AUTO;
    P2PState p2p;
    bool p2p_active = false;
    std::vector<void*> p2p_opened;       // cudaIpcOpenMemHandle mappings to close
    bool debug_capture = false;          // cdl_debug_capture
    int cells_layout = CDL_CELLS_AUTO;   // cdl_set_cells_layout
    bool use_sell = false;               // decided per run
    // device scratch of the fetch calls that transpose into R's layout (kept: a clone_id(n_chain = 64) call makes 128 such
    // fetches, and a cudaMalloc / cudaFree pair each cost more than the copy); at most 64 MB stay resident
    DevBuf<double> fetch_scratch;
};

namespace {

void use_device(cdl_handle h) { CDL_CUDA(cudaSetDevice(h->device)); }

template <typename F>
int guard(cdl_handle h, F&& f) {
    try {
        if (!h) throw Error(CDL_ERR_INVALID, "null handle");
        use_device(h);
        f();
        return CDL_OK;
    } catch (const Error& e) {
        if (h) h->err = e.what(); else g_create_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        if (h) h->err = "host allocation failed";
        return CDL_ERR_NOMEM;
    } catch (const std::exception& e) {
        if (h) h->err = e.what();
        return CDL_ERR_INVALID;
    }
}

// `count` doubles of device scratch for a fetch call: the handle's kept buffer when the request is small, a one-off
// allocation (freed on scope exit) otherwise.  Every user synchronises the stream before it returns.
struct ScratchUse {
    double* p = nullptr;
    DevBuf<double> own;
    ScratchUse(cdl_handle h, size_t count) {
        if (count * sizeof(double) <= ((size_t)64 << 20)) {
            if (h->fetch_scratch.n < count) h->fetch_scratch.alloc(count);
            p = h->fetch_scratch.p;
        } else {
            own.alloc(count);
            p = own.p;
        }
    }
};

constexpr size_t SPARE_BYTES = (size_t)16 << 30;
size_t chain_bytes(const Chain& c) {
    size_t b = 0;
    auto add = [&](const auto& buf) { if (buf.p) b += sizeof(*buf.p) * (buf.n ? buf.n : 1); };
    add(c.scal); add(c.l1); add(c.l1c); add(c.theta0_all); add(c.theta1_all); add(c.loglik_all); add(c.relax_all);
    add(c.prob_all); add(c.prob_acc); add(c.part_d); add(c.el_l1); add(c.el_l1c); add(c.sd1); add(c.prob_sq);
    add(c.mask); add(c.z); add(c.config_all); add(c.assign_all); add(c.sab); add(c.part_u); add(c.sab_zero);
    add(c.ticket); add(c.work); add(c.sell_tickets); add(c.stats_work); add(c.chg_ctl); add(c.chg_list); add(c.chg_old);
    add(c.sell_scratch); add(c.tab_g); add(c.dbg_s34); add(c.dbg_sab); add(c.tl);
    add(c.prob_mean); add(c.config_prob); add(c.theta1_mean);
    return b;
}
// The results of the last run are dropped; the chains' device buffers stay with the handle (up to SPARE_BYTES) for the
// next run to take over.  Every API call leaves its stream synchronised, so nothing is still using them.
void retire_chains(cdl_handle h) {
    size_t total = 0;
    for (const auto& c : h->spare) total += chain_bytes(*c);
    for (auto& c : h->chains) {
        const size_t b = chain_bytes(*c);
        if (total + b > SPARE_BYTES) continue;          // freed with h->chains below
        total += b;
        h->spare.push_back(std::move(c));
    }
    h->chains.clear();
}

void nccl_load(cdl_handle h, const char* path) {
    if (h->nccl.lib) return;
    const char* cand[] = {path, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* c : cand) {
        if (!c || !*c) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) throw Error(CDL_ERR_COMM, std::string("cannot dlopen NCCL: ") + dlerror());
    h->nccl.lib = lib;
    h->nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    h->nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(lib, "ncclCommInitRank");
    h->nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    h->nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    h->nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    h->nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!h->nccl.GetUniqueId || !h->nccl.CommInitRank || !h->nccl.AllReduce || !h->nccl.CommDestroy)
        throw Error(CDL_ERR_COMM, "NCCL symbols missing");
}

void nccl_check(cdl_handle h, int rc, const char* what) {
    if (rc != 0) {
        const char* s = h->nccl.GetErrorString ? h->nccl.GetErrorString(rc) : "?";
        throw Error(CDL_ERR_COMM, std::string(what) + ": " + s);
    }
}

// ncclDataType_t / ncclRedOp_t values (nccl.h): ncclUint64 = 5, ncclFloat64 = 8, ncclSum = 0
constexpr int NCCL_U64 = 5, NCCL_F64 = 8, NCCL_SUM = 0;

void allreduce_u64(cdl_handle h, unsigned long long* buf, size_t n) {
    if (h->world <= 1) return;
    nccl_check(h, h->nccl.AllReduce(buf, buf, n, NCCL_U64, NCCL_SUM, h->comm, h->st), "ncclAllReduce(u64)");
}
void allreduce_d(cdl_handle h, double* buf, size_t n) {
    if (h->world <= 1) return;
    nccl_check(h, h->nccl.AllReduce(buf, buf, n, NCCL_F64, NCCL_SUM, h->comm, h->st), "ncclAllReduce(f64)");
}
double allreduce_scalar_d(cdl_handle h, double v) {
    if (h->world <= 1) return v;
    DevBuf<double> b(1);
    b.upload(&v, 1, h->st);
    nccl_check(h, h->nccl.AllReduce(b.p, b.p, 1, NCCL_F64, NCCL_SUM, h->comm, h->st), "ncclAllReduce(f64)");
    double out = 0;
    b.download(&out, 1, h->st);
    return out;
}

// ---- host-side input normalisation -----------------------------------------------------------

struct RawCsc {
    std::vector<int64_t> p;
    std::vector<int32_t> i;
    std::vector<uint16_t> a, d;
};

inline bool is_na(double x) { return x != x; }

void check_count(double v, const char* what) {
    if (v != std::floor(v)) throw Error(CDL_ERR_INVALID, "A and/or D contain non-integer values.");
    if (v < 0) throw Error(CDL_ERR_INVALID, std::string("negative count in ") + what);
    if (v > 65535.0)
        throw Error(CDL_ERR_UNSUPPORTED, std::string("count above 65535 in ") + what + " (uint16 storage)");
}

void upload_raw(cdl_handle h, int64_t N, int64_t M, const RawCsc& r) {
    const int64_t nnz = (int64_t)r.i.size();
    DevBuf<int64_t> p((size_t)M + 1);
    DevBuf<int32_t> vi((size_t)nnz);
    DevBuf<uint16_t> a((size_t)nnz), d((size_t)nnz);
    p.upload(r.p.data(), (size_t)M + 1, h->st);
    if (nnz) {
        vi.upload(r.i.data(), (size_t)nnz, h->st);
        a.upload(r.a.data(), (size_t)nnz, h->st);
        d.upload(r.d.data(), (size_t)nnz, h->st);
    }
    cudaFree(h->donor.labels);
    h->donor.labels = nullptr;
    h->donor.M_total = 0;
    h->donor.cell_offset = 0;
    donor_build(h->donor, N, M, nnz, p.p, vi.p, a.p, d.p, h->st);
    h->has_data = true;
    retire_chains(h);
    h->have_sweep = false;
}

void require_data(cdl_handle h) {
    if (!h->has_data) throw Error(CDL_ERR_STATE, "no donor loaded (call cdl_load_* first)");
}

Chain& get_chain(cdl_handle h, int c) {
    if (c < 0 || c >= (int)h->chains.size() || !h->chains[(size_t)c]->done)
        throw Error(CDL_ERR_STATE, "no finished Gibbs chain with that index");
    return *h->chains[(size_t)c];
}

std::vector<uint32_t> config_masks(const double* Config, int64_t N, int K, bool require_binary) {
    std::vector<uint32_t> m((size_t)N, 0u);
    for (int k = 0; k < K; ++k)
        for (int64_t i = 0; i < N; ++i) {
            const double v = Config[i + (int64_t)k * N];
            if (v == 1.0) m[(size_t)i] |= 1u << k;
            else if (v != 0.0 && require_binary)
                throw Error(CDL_ERR_UNSUPPORTED, "Config must be binary (0/1) for this path");
        }
    return m;
}

void fill_logpsi(const double* Psi, int K, double* out) {
    for (int k = 0; k < KMAX; ++k) out[k] = 0.0;
    for (int k = 0; k < K; ++k) {
        const double p = Psi ? Psi[k] : 1.0 / K;
        if (!(p > 0.0)) throw Error(CDL_ERR_INVALID, "Psi must be positive");
        out[k] = std::log(p);
    }
}

void dic_from_trace(const std::vector<double>& ll, double logLik_post, double out[6]) {
    const size_t n = ll.size();
    double m = 0;
    for (double v : ll) m += v;
    m /= (double)n;
    double var = NAN;
    if (n > 1) {
        double s = 0;
        for (double v : ll) s += (v - m) * (v - m);
        var = s / (double)(n - 1);
    }
    const double pDS = -2 * m - (-2 * logLik_post);
    const double DICS = -2 * logLik_post + 2 * pDS;
    const double DICG = -2 * logLik_post + 2 * (2 * var);
    out[0] = DICG; out[1] = var; out[2] = -2 * m; out[3] = -2 * logLik_post; out[4] = DICG; out[5] = DICS;
}

constexpr int NCCL_U8 = 1, NCCL_MIN = 3;   // ncclUint8 / ncclMin (nccl.h)

void p2p_release(cdl_handle h) {
    for (void* q : h->p2p_opened) cudaIpcCloseMemHandle(q);
    h->p2p_opened.clear();
    cudaFree(h->p2p.arena); cudaFree(h->p2p.my_flags); cudaFree(h->p2p.error);
    h->p2p = P2PState();
    h->p2p_active = false;
}

// Collective over the ranks: make sure every rank owns a table arena of `elems` uint64 per buffer and has the
// peers' arenas and flag arrays mapped.  Falls back to the NCCL all-reduce (on ALL ranks) if any rank cannot
// map peer memory.  Returns whether the peer-memory exchange is active.
bool p2p_ensure(cdl_handle h, size_t elems) {
    if (h->world <= 1 || h->exchange_mode == CDL_EXCHANGE_NCCL || h->world - 1 > MAX_PEERS || !h->nccl.AllGather) {
        if (h->exchange_mode == CDL_EXCHANGE_PEER && h->world > 1)
            throw Error(CDL_ERR_COMM, "peer-memory exchange requested but not available (world size or NCCL symbols)");
        return false;
    }
    if (h->p2p_active && h->p2p.table_elems == elems) return true;
    p2p_release(h);
    P2PState& q = h->p2p;
    int ok = 1;
    struct Pack { cudaIpcMemHandle_t arena, flags; };
    Pack mine{};
    q.table_elems = elems; q.my_rank = h->rank; q.n_peer = h->world - 1;
    if (cudaMalloc(&q.arena, sizeof(unsigned long long) * 3 * elems) != cudaSuccess ||
        cudaMalloc(&q.my_flags, sizeof(unsigned long long) * (size_t)h->world) != cudaSuccess ||
        cudaMalloc(&q.error, sizeof(unsigned int)) != cudaSuccess) ok = 0;
    if (ok) {
        cudaMemsetAsync(q.arena, 0, sizeof(unsigned long long) * 3 * elems, h->st);
        cudaMemsetAsync(q.my_flags, 0, sizeof(unsigned long long) * (size_t)h->world, h->st);
        cudaMemsetAsync(q.error, 0, sizeof(unsigned int), h->st);
        if (cudaIpcGetMemHandle(&mine.arena, q.arena) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.flags, q.my_flags) != cudaSuccess) ok = 0;
    }
    cudaGetLastError();
    // exchange the handles with the communicator we already have
    std::vector<Pack> all((size_t)h->world);
    {
        DevBuf<uint8_t> snd(sizeof(Pack)), rcv(sizeof(Pack) * (size_t)h->world);
        snd.upload(reinterpret_cast<const uint8_t*>(&mine), sizeof(Pack), h->st);
        nccl_check(h, h->nccl.AllGather(snd.p, rcv.p, sizeof(Pack), NCCL_U8, h->comm, h->st), "ncclAllGather(ipc handles)");
        rcv.download(reinterpret_cast<uint8_t*>(all.data()), sizeof(Pack) * (size_t)h->world, h->st);
    }
    int slot = 0;
    for (int r = 0; r < h->world && ok; ++r) {
        if (r == h->rank) continue;
        void *pa = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pa, all[(size_t)r].arena, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        h->p2p_opened.push_back(pa);
        if (cudaIpcOpenMemHandle(&pf, all[(size_t)r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        h->p2p_opened.push_back(pf);
        q.peer_arena[slot] = (unsigned long long*)pa;
        q.peer_flags[slot] = (unsigned long long*)pf;
        q.peer_rank[slot] = r;
        ++slot;
    }
    cudaGetLastError();
    // every rank must take the same path
    {
        DevBuf<uint8_t> f(1);
        const uint8_t v = (uint8_t)ok;
        f.upload(&v, 1, h->st);
        nccl_check(h, h->nccl.AllReduce(f.p, f.p, 1, NCCL_U8, NCCL_MIN, h->comm, h->st), "ncclAllReduce(p2p ok)");
        uint8_t all_ok = 0;
        f.download(&all_ok, 1, h->st);
        ok = all_ok;
    }
    if (!ok) {
        p2p_release(h);
        if (h->exchange_mode == CDL_EXCHANGE_PEER)
            throw Error(CDL_ERR_COMM, "peer-memory exchange requested but cudaIpc mapping failed on some rank");
        return false;
    }
    q.epoch = 0;
    h->p2p_active = true;
    return true;
}

void one_sweep(cdl_handle h, Chain& ch, SweepArgs& a, uint32_t it, const cdl_replay* rp_dev, int64_t n_el,
               uint32_t rng_advance = 0) {
    a.it = it;
    a.rng_it = it + rng_advance;
    if (ch.tl.p) ch.dev.tl = ch.tl.p + (size_t)(a.rng_it & 1u) * 3 * TL_ROWS * TL_SLOTS;   // timeline: two sweeps, by parity
    const cdl_gibbs_params& gp = h->gp;
    a.learn_relax_next = (a.relax && !a.relax_fixed_set && ((double)(it + 1) > 0.1 * gp.min_iter + 5.0)) ? 1 : 0;
    if (rp_dev) {
        const int64_t row = (int64_t)it - 1;
        a.rp_u_assign = rp_dev->u_assign ? rp_dev->u_assign + row * a.M : nullptr;
        a.rp_u_config = rp_dev->u_config ? rp_dev->u_config + row * a.N * a.K : nullptr;
        a.rp_theta0 = rp_dev->theta0 ? rp_dev->theta0 + row : nullptr;
        a.rp_theta1 = rp_dev->theta1 ? rp_dev->theta1 + row * n_el : nullptr;
        a.rp_relax_next = (rp_dev->relax_rate && (int64_t)it < a.T) ? rp_dev->relax_rate + row + 1 : nullptr;
    }
    if (a.wise == 2) {
        // wise = element: per-entry theta1 (element.cu)
        a.p2p = nullptr;
        launch_el_cells(h->donor, a, ch.dev, h->st);
        launch_el_stats(h->donor, a, ch.dev, h->st);
        // cell shards: the read sums and the per-(variant, clone) log-odds sums over all ranks (every rank receives the
        // same bits, so the redundant table steps stay in lockstep)
        allreduce_u64(h, ch.dev.sab, (size_t)(a.N * a.K));
        allreduce_d(h, ch.dev.sd1, (size_t)(a.N * a.K));
        launch_table(h->donor, a, ch.dev, h->st);
        launch_el_theta(h->donor, a, ch.dev, h->st);
        h->launches += 4;
        return;
    }
    if (h->use_sell) launch_cells_sell(h->donor, a, ch.dev, h->sm_count, h->st);
    else launch_cells(h->donor, a, ch.dev, h->sm_count, h->st);
    auto capture = [&]() {       // test hook: this rank's table as the statistic kernel left it
        if (ch.dev.dbg_sab)
            CDL_CUDA(cudaMemcpyAsync(ch.dev.dbg_sab, ch.dev.sab, sizeof(unsigned long long) * (size_t)(a.N * a.K),
                                     cudaMemcpyDeviceToDevice, h->st));
    };
    if (h->p2p_active) {
        // statistic table into this sweep's half of the peer-visible arena; publish; the table kernel sums
        // the peers' halves while it reads them (no collective call).  With incremental statistics the previous
        // table kernel has left this rank's previous table in this half, for the statistic kernel to correct.
        P2PState& q = h->p2p;
        ++q.epoch;
        ch.dev.sab = q.arena + (size_t)(q.epoch & 1ull) * q.table_elems;
        a.p2p = &q;
        // (the last CTA of the kernel that completes the table publishes it to the peers)
        launch_stats(h->donor, a, ch.dev, h->sm_count, h->st);     // incremental correction or the full pass
        capture();
        launch_table(h->donor, a, ch.dev, h->st);
        h->launches += 3;
    } else {
        a.p2p = nullptr;
        launch_stats(h->donor, a, ch.dev, h->sm_count, h->st);     // incremental correction or the full pass
        capture();
        allreduce_u64(h, ch.dev.sab, (size_t)(a.N * a.K));
        launch_table(h->donor, a, ch.dev, h->st);
        h->launches += 3;   // k_cells / k_cells_sell, k_stats, k_table (+ memset nodes)
    }
}

double geweke_fraction(cdl_handle h, Chain& ch, int64_t n_rows, DevBuf<unsigned long long>& counts) {
    launch_geweke(ch.dev.prob_all, h->donor.M * h->K, n_rows, counts.p, h->st);
    if (h->world > 1) allreduce_u64(h, counts.p, 2);
    unsigned long long c[2];
    counts.download(c, 2, h->st);
    if (c[0] == 0) return NAN;
    return (double)c[1] / (double)c[0];
}

// CDL_TIMELINE=1: where the LAST sweep of a bench pass spent its time, from %globaltimer stamps the kernels left
// (common.cuh::tl_stamp).  Microseconds after the first CTA of the cell kernel started; "first" / "last" over the CTAs.
void print_timeline(cdl_handle h, Chain& ch) {
    const size_t per = (size_t)3 * TL_ROWS * TL_SLOTS;
    std::vector<unsigned long long> t(2 * per);
    ch.tl.download(t.data(), t.size(), h->st);
    auto at = [&](int sweep, int kern, int row, int slot) {
        return t[(size_t)sweep * per + ((size_t)kern * TL_ROWS + (size_t)row) * TL_SLOTS + (size_t)slot];
    };
    // the two halves hold the last two sweeps (by parity of the Philox iteration): the older one first
    unsigned long long first[2] = {~0ull, ~0ull};
    for (int sw = 0; sw < 2; ++sw)
        for (int r = 0; r < TL_ROWS; ++r) if (at(sw, 0, r, 0) && at(sw, 0, r, 0) < first[sw]) first[sw] = at(sw, 0, r, 0);
    if (first[0] == ~0ull && first[1] == ~0ull) return;
    int order2[2] = {0, 1};
    if (first[1] < first[0]) { order2[0] = 1; order2[1] = 0; }
    const unsigned long long t0 = std::min(first[0], first[1]);
    static const char* names[3][TL_SLOTS] = {
        {"cells: CTA start", "cells: previous kernel complete (griddepcontrol.wait)", "cells: weight table staged",
         "cells: last warp of the CTA done", nullptr, nullptr, nullptr, nullptr},
        {"stats: CTA start", "stats: cell kernel complete", nullptr, "stats: CTA done (table corrected / full pass done)",
         nullptr, nullptr, nullptr, nullptr},
        {"table: CTA start", "table: statistic kernel complete", "table: peers' tables published (flags)",
         "table: flips + block sums done (ticket)", "table: LAST block, grid totals reduced", "table: Gamma jobs done",
         "table: LAST block writes theta0 / relax rate", "table: Gamma candidates ready (before the wait)"}};
    static const int order[3][TL_SLOTS] = {{0, 1, 2, 3, -1, -1, -1, -1}, {0, 1, 3, -1, -1, -1, -1, -1}, {0, 7, 1, 2, 3, 4, 5, 6}};
    if (const char* e = getenv("CDL_TIMELINE")) {
        if (atoi(e) >= 2) {          // raw rows of the cell kernel: CTA, SM, microseconds from the wait to the CTA's last warp
            const int sw = order2[1];
            std::vector<int32_t> asg;
            std::vector<uint32_t> off;
            const Donor& d = h->donor;
            if (d.s_assign && d.s_off) {
                asg.resize((size_t)d.s_assign_grid * d.s_assign_warps);
                off.resize((size_t)d.s_slices + 1);
                cudaMemcpy(asg.data(), d.s_assign, 4 * asg.size(), cudaMemcpyDeviceToHost);
                cudaMemcpy(off.data(), d.s_off, 4 * off.size(), cudaMemcpyDeviceToHost);
            }
            for (int r = 0; r < TL_ROWS; ++r) {
                if (!at(sw, 0, r, 0)) continue;
                long long work = 0, longest = 0;
                if (!asg.empty() && r < d.s_assign_grid)
                    for (int w = 0; w < d.s_assign_warps; ++w) {
                        const int32_t sl = asg[(size_t)r * d.s_assign_warps + w];
                        if (sl >= 0) { const long long l = off[(size_t)sl + 1] - off[(size_t)sl]; work += l; longest = std::max(longest, l); }
                    }
                std::fprintf(stderr, "[cdl cta] %d sm %llu staged %.2f done %.2f slots %lld longest %lld\n", r, at(sw, 0, r, 4),
                             ((double)at(sw, 0, r, 2) - (double)t0) * 1e-3, ((double)at(sw, 0, r, 3) - (double)t0) * 1e-3, work, longest);
            }
        }
    }
    std::fprintf(stderr, "[cdl timeline r%d] last two sweeps, microseconds after the first cell-kernel CTA of the earlier one started; "
                         "first / last over the CTAs\n", h->rank);
    for (int q2 = 0; q2 < 2; ++q2) {
        const int sw = order2[q2];
        if (first[sw] == ~0ull) continue;
        for (int k = 0; k < 3; ++k)
            for (int q = 0; q < TL_SLOTS; ++q) {
                const int sl = order[k][q];
                if (sl < 0 || !names[k][sl]) continue;
                double lo = 1e30, hi = -1e30;
                int n = 0;
                for (int r = 0; r < TL_ROWS; ++r) {
                    const unsigned long long v = at(sw, k, r, sl);
                    if (!v || v < first[sw]) continue;            // (stamps of earlier sweeps: only the last block writes some slots)
                    const double us = ((double)v - (double)t0) * 1e-3;
                    lo = std::min(lo, us); hi = std::max(hi, us); ++n;
                }
                if (n) std::fprintf(stderr, "[cdl timeline r%d] sweep %c  %-58s first %9.2f  last %9.2f  (%d CTAs)\n", h->rank,
                                    q2 ? 'B' : 'A', names[k][sl], lo, hi, n);
            }
    }
}

double r_round3_percent(double frac) {   // round(x, 3) * 100 (R/clone_id.R:585,595)
    if (frac != frac) return frac;
    return std::nearbyint(frac * 1000.0) / 1000.0 * 100.0;
}

}  // namespace

// =================================================================================================
extern "C" {

int cdl_version(void) { return 100; }

const char* cdl_last_error(cdl_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cdl_create(int device, cdl_handle* out) {
    if (!out) return CDL_ERR_INVALID;
    *out = nullptr;
    try {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw Error(CDL_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") +
                                          cudaGetErrorString(e));
        if (device < 0 || device >= n) throw Error(CDL_ERR_INVALID, "device index out of range");
        CDL_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        CDL_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            throw Error(CDL_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only");
        std::unique_ptr<cdl_handle_s> h(new cdl_handle_s());
        h->device = device;
        h->sm_count = prop.multiProcessorCount;
        CDL_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
        *out = h.release();
        return CDL_OK;
    } catch (const Error& e) {
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return CDL_ERR_CUDA;
    }
}

int cdl_destroy(cdl_handle h) {
    if (!h) return CDL_OK;
    cudaSetDevice(h->device);
    p2p_release(h);
    if (h->comm && h->nccl.CommDestroy) h->nccl.CommDestroy(h->comm);
    h->chains.clear();
    h->spare.clear();
    donor_free(h->donor);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return CDL_OK;
}

int cdl_load_dense(cdl_handle h, int64_t N, int64_t M, const double* A, const double* D) {
    return guard(h, [&] {
        if (!A || !D || N <= 0 || M <= 0) throw Error(CDL_ERR_INVALID, "A and D must be non-empty matrices");
        RawCsc r;
        r.p.assign((size_t)M + 1, 0);
        for (int64_t j = 0; j < M; ++j) {
            for (int64_t i = 0; i < N; ++i) {
                const double dv = D[i + j * N], av = A[i + j * N];
                if (!is_na(av)) { if (av != std::floor(av)) throw Error(CDL_ERR_INVALID, "A and/or D contain non-integer values."); }
                if (is_na(dv)) {
                    // D missing (NA) with reads in A: the reference keeps A there (only D == 0 clears A, :380), then
                    // B1 = D - A is NA -> 0 (:391): alternative reads with no depth.  Refused rather than guessed.
                    if (!is_na(av) && av != 0.0)
                        throw Error(CDL_ERR_INVALID, "A has a count where D is missing (NA); the reference would "
                                                     "count those reads with zero depth");
                    continue;
                }
                check_count(dv, "D");
                if (dv == 0.0) continue;               // D == 0 -> NA (R/clone_id.R:380-381)
                double a_use = is_na(av) ? 0.0 : av;   // A NA where D > 0 -> 0 (:382)
                check_count(a_use, "A");
                if (a_use > dv) throw Error(CDL_ERR_INVALID, "A exceeds D for some entry");
                r.i.push_back((int32_t)i);
                r.a.push_back((uint16_t)a_use);
                r.d.push_back((uint16_t)dv);
            }
            r.p[(size_t)j + 1] = (int64_t)r.i.size();
        }
        upload_raw(h, N, M, r);
    });
}

int cdl_load_csc(cdl_handle h, int64_t N, int64_t M, const int64_t* d_p, const int32_t* d_i, const double* d_x,
                 const int64_t* a_p, const int32_t* a_i, const double* a_x) {
    return guard(h, [&] {
        if (!d_p || !a_p || N <= 0 || M <= 0) throw Error(CDL_ERR_INVALID, "bad sparse input");
        RawCsc r;
        r.p.assign((size_t)M + 1, 0);
        for (int64_t j = 0; j < M; ++j) {
            int64_t ea = a_p[j];
            const int64_t ea1 = a_p[j + 1];
            for (int64_t e = d_p[j]; e < d_p[j + 1]; ++e) {
                const int32_t i = d_i[e];
                const double dv = d_x[e];
                if (i < 0 || i >= N) throw Error(CDL_ERR_INVALID, "row index out of range in D");
                if (is_na(dv) || dv == 0.0) continue;
                check_count(dv, "D");
                // A entries at positions where D is zero or absent are dropped, as A[which(D == 0)] <- NA does
                // (R/clone_id.R:380); the dense loader and SparseCounts.from_sparse do the same
                while (ea < ea1 && a_i[ea] < i) ++ea;
                double av = 0.0;
                if (ea < ea1 && a_i[ea] == i) { av = is_na(a_x[ea]) ? 0.0 : a_x[ea]; ++ea; }
                check_count(av, "A");
                if (av > dv) throw Error(CDL_ERR_INVALID, "A exceeds D for some entry");
                r.i.push_back(i);
                r.a.push_back((uint16_t)av);
                r.d.push_back((uint16_t)dv);
            }
            r.p[(size_t)j + 1] = (int64_t)r.i.size();
        }
        upload_raw(h, N, M, r);
    });
}

int cdl_load_csc_u16(cdl_handle h, int64_t N, int64_t M, const int64_t* p, const int32_t* i, const uint16_t* a,
                     const uint16_t* d) {
    return guard(h, [&] {
        if (!p || N <= 0 || M <= 0) throw Error(CDL_ERR_INVALID, "bad sparse input");
        const int64_t nnz = p[M];
        if (p[0] != 0 || nnz < 0) throw Error(CDL_ERR_INVALID, "bad column pointers");
        DevBuf<int64_t> pd((size_t)M + 1);
        DevBuf<int32_t> vi((size_t)nnz);
        DevBuf<uint16_t> ad((size_t)nnz), dd((size_t)nnz);
        pd.upload(p, (size_t)M + 1, h->st);
        if (nnz) {
            vi.upload(i, (size_t)nnz, h->st);
            ad.upload(a, (size_t)nnz, h->st);
            dd.upload(d, (size_t)nnz, h->st);
        }
        cudaFree(h->donor.labels);
        h->donor.labels = nullptr;
        h->donor.M_total = 0;
        h->donor.cell_offset = 0;
        donor_build(h->donor, N, M, nnz, pd.p, vi.p, ad.p, dd.p, h->st);
        h->has_data = true;
        retire_chains(h);
        h->have_sweep = false;
    });
}

int cdl_synth_generate(cdl_handle h, int64_t N, int64_t M_total, int64_t cell_offset, int64_t M_local, int32_t K,
                       const double* Config, double coverage, uint64_t seed) {
    return guard(h, [&] {
        if (!Config || N <= 0 || M_local <= 0 || cell_offset < 0 || cell_offset + M_local > M_total)
            throw Error(CDL_ERR_INVALID, "bad synthetic donor shape");
        synth_generate(h->donor, N, M_total, cell_offset, M_local, K, Config, coverage, seed, h->st);
        h->has_data = true;
        retire_chains(h);
        h->have_sweep = false;
    });
}

int cdl_synth_labels(cdl_handle h, int32_t* labels_out) {
    return guard(h, [&] {
        require_data(h);
        if (!h->donor.labels) throw Error(CDL_ERR_STATE, "donor is not synthetic");
        CDL_CUDA(cudaMemcpyAsync(labels_out, h->donor.labels, 4 * (size_t)h->donor.M, cudaMemcpyDeviceToHost, h->st));
        CDL_CUDA(cudaStreamSynchronize(h->st));
    });
}

int cdl_export_csc_u16(cdl_handle h, int64_t* p, int32_t* i, uint16_t* a, uint16_t* d) {
    return guard(h, [&] {
        require_data(h);
        donor_export(h->donor, p, i, a, d, h->st);
    });
}

int cdl_data_info(cdl_handle h, int64_t* N, int64_t* M, int64_t* nnz, double* W_log) {
    return guard(h, [&] {
        require_data(h);
        if (N) *N = h->donor.N;
        if (M) *M = h->donor.M;
        if (nnz) *nnz = h->donor.nnz;
        if (W_log) *W_log = h->donor.W_log;
    });
}

int cdl_set_shard(cdl_handle h, int64_t cell_offset, int64_t M_total) {
    return guard(h, [&] {
        require_data(h);
        if (cell_offset < 0 || cell_offset + h->donor.M > M_total || M_total >= (int64_t)0xFFFFFFFFll)
            throw Error(CDL_ERR_INVALID, "bad shard bounds");
        h->donor.cell_offset = cell_offset;
        h->donor.M_total = M_total;
    });
}

int cdl_set_cells_layout(cdl_handle h, int32_t layout) {
    return guard(h, [&] {
        if (layout != CDL_CELLS_AUTO && layout != CDL_CELLS_ROWS && layout != CDL_CELLS_SLICES && layout != CDL_CELLS_SMALL)
            throw Error(CDL_ERR_INVALID, "unknown cells layout");
        h->cells_layout = layout;
    });
}

int cdl_set_exchange_mode(cdl_handle h, int32_t mode) {
    return guard(h, [&] {
        if (mode != CDL_EXCHANGE_AUTO && mode != CDL_EXCHANGE_NCCL && mode != CDL_EXCHANGE_PEER)
            throw Error(CDL_ERR_INVALID, "unknown exchange mode");
        h->exchange_mode = mode;
        if (mode == CDL_EXCHANGE_NCCL) p2p_release(h);
    });
}

int cdl_exchange_is_peer(cdl_handle h) { return (h && h->p2p_active) ? 1 : 0; }

int cdl_comm_unique_id(cdl_handle h, const char* nccl_lib_path, uint8_t* id128) {
    return guard(h, [&] {
        nccl_load(h, nccl_lib_path);
        nccl_check(h, h->nccl.GetUniqueId(id128), "ncclGetUniqueId");
    });
}

int cdl_comm_init(cdl_handle h, const char* nccl_lib_path, const uint8_t* id128, int32_t rank, int32_t world) {
    return guard(h, [&] {
        if (world < 1 || rank < 0 || rank >= world) throw Error(CDL_ERR_INVALID, "bad rank/world");
        nccl_load(h, nccl_lib_path);
        Id128 id;
        std::memcpy(id.b, id128, 128);
        nccl_check(h, h->nccl.CommInitRank(&h->comm, world, id, rank), "ncclCommInitRank");
        h->rank = rank;
        h->world = world;
    });
}

void cdl_gibbs_defaults(cdl_gibbs_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->relax_Config = 1;
    p->relax_rate_fixed_set = 0;
    p->relax_rate_prior[0] = 1; p->relax_rate_prior[1] = 9;
    p->keep_base_clone = 1;
    p->prior0[0] = 0.2; p->prior0[1] = 99.8;
    p->prior1[0] = 0.45; p->prior1[1] = 0.55;
    p->min_iter = 5000; p->max_iter = 20000;
    p->buin_frac = 0.5;
    p->wise = CDL_WISE_VARIANT;
    p->n_chain = 1;
    p->seed = 1;
    p->keep_prob_all = -1;
    p->trace_budget_bytes = (int64_t)48 << 30;
    p->check_every = 100;
    p->incremental_stats = -1;
    p->snapshot_fp32 = 1;
}

int cdl_gibbs_run(cdl_handle h, const double* Config, int32_t K, const double* Psi, const cdl_gibbs_params* params,
                  const cdl_replay* replay) {
    return guard(h, [&] {
        require_data(h);
        if (!Config || !params) throw Error(CDL_ERR_INVALID, "Config and params are required");
        const cdl_gibbs_params gp = *params;
        const Donor& d = h->donor;
        if (K < 1 || K > KMAX) throw Error(CDL_ERR_UNSUPPORTED, "number of clones must be in 1..32");
        if (gp.wise != CDL_WISE_GLOBAL && gp.wise != CDL_WISE_VARIANT && gp.wise != CDL_WISE_ELEMENT)
            throw Error(CDL_ERR_INVALID, "Input wise mode: unknown, while only supporting: element, variant, global");
        const bool element = gp.wise == CDL_WISE_ELEMENT;
        if (gp.max_iter < 2) throw Error(CDL_ERR_INVALID, "max_iter must be at least 2");
        if (gp.relax_Config && gp.relax_rate_fixed_set && (gp.relax_rate_fixed > 1 || gp.relax_rate_fixed < 0))
            throw Error(CDL_ERR_INVALID, "relax_rate_fixed needs to be NULL or in [0, 1].");
        if (gp.n_chain < 1) throw Error(CDL_ERR_INVALID, "n_chain must be >= 1");
        if (gp.chain_offset < 0 || gp.chain_offset + gp.n_chain > (1 << 23))
            throw Error(CDL_ERR_INVALID, "chain_offset + n_chain must stay below 2^23 (Philox stream word)");
        if (replay && gp.n_chain != 1) throw Error(CDL_ERR_INVALID, "replay mode runs a single chain");
        if (replay && replay->n_rows < gp.max_iter) throw Error(CDL_ERR_INVALID, "replay arrays shorter than max_iter");
        if (gp.acc_fp32) throw Error(CDL_ERR_UNSUPPORTED, "acc_fp32 is reserved");
        if (d.max_variant_depth >= (1ull << 32))
            throw Error(CDL_ERR_UNSUPPORTED, "a variant's total depth exceeds 2^32 (packed statistics)");

        retire_chains(h);
        if (h->spare.size() > (size_t)gp.n_chain) h->spare.resize((size_t)gp.n_chain);     // more than this run can take over
        h->gp = gp;
        h->K = K;
        // cell kernel: one lane per cell on the slice-major layout for large donors, row-split otherwise.
        // The choice depends on the WHOLE donor's cell count so that it is the same on every rank.
        h->use_sell = h->cells_layout == CDL_CELLS_SLICES ||
                      (h->cells_layout == CDL_CELLS_AUTO && d.M_total >= 16384);
        if (element) h->use_sell = false;
        if (h->use_sell) donor_build_sell(h->donor, h->st);
        p2p_ensure(h, (size_t)(d.N * K));
        const int64_t N = d.N, M = d.M, T = gp.max_iter;
        const int64_t n_el = gp.wise == CDL_WISE_VARIANT ? N : (element ? d.nnz : 1);
        if (element) {
            if ((double)T * (double)n_el * 8.0 > (double)gp.trace_budget_bytes)
                throw Error(CDL_ERR_UNSUPPORTED, "wise = \"element\": the max_iter x nnz theta1 trace exceeds trace_budget_bytes");
            donor_build_elem0(h->donor, h->st);
        }
        h->T = T;
        h->n_el = n_el;
        std::vector<uint32_t> masks = config_masks(Config, N, K, true);
        h->cfg0.alloc((size_t)N);
        h->cfg0.upload(masks.data(), (size_t)N, h->st);
        if (gp.prior1_matrix && gp.prior1_matrix_rows != 0 && gp.prior1_matrix_rows != n_el)
            throw Error(CDL_ERR_INVALID, "prior1 must be a matrix with one row per theta1 element (" +
                                             std::to_string(n_el) + " rows for this wise mode) and two columns");
        if (gp.prior1_matrix) {
            h->prior1_mat.alloc((size_t)n_el * 2);
            h->prior1_mat.upload(gp.prior1_matrix, (size_t)n_el * 2, h->st);
        } else {
            h->prior1_mat.release();
        }

        SweepArgs a{};
        a.N = N; a.M = M; a.M_total = d.M_total; a.cell_offset = d.cell_offset;
        a.K = K; a.wise = gp.wise; a.relax = gp.relax_Config ? 1 : 0; a.keep_base = gp.keep_base_clone ? 1 : 0;
        for (int q = 0; q < 2; ++q) {
            a.prior0[q] = gp.prior0[q]; a.prior1[q] = gp.prior1[q]; a.relax_prior[q] = gp.relax_rate_prior[q];
        }
        a.prior1_mat = gp.prior1_matrix ? h->prior1_mat.p : nullptr;
        a.relax_fixed = gp.relax_rate_fixed; a.relax_fixed_set = gp.relax_rate_fixed_set ? 1 : 0;
        fill_logpsi(Psi, K, a.logPsi);
        a.cfg0 = h->cfg0.p;
        {
            // base clone (column 1 of Config, all zero in a clonal tree): with keep_base_clone or without relaxation
            // its column never changes, so no variant ever contributes to its logit
            bool col0_zero = true;
            for (uint32_t m : masks) if (m & 1u) { col0_zero = false; break; }
            a.skip0 = (col0_zero && (!gp.relax_Config || gp.keep_base_clone)) ? 1 : 0;
        }
        a.W_log = allreduce_scalar_d(h, d.W_log);
        a.el_offset = 0;
        if (element && h->world > 1) {
            // every rank's covered-entry count -> this rank's first global element index
            std::vector<unsigned long long> cnt((size_t)h->world, 0ull);
            cnt[(size_t)h->rank] = (unsigned long long)d.nnz;
            DevBuf<unsigned long long> b((size_t)h->world);
            b.upload(cnt.data(), cnt.size(), h->st);
            allreduce_u64(h, b.p, cnt.size());
            b.download(cnt.data(), cnt.size(), h->st);
            unsigned long long tot = 0;
            for (int r = 0; r < h->world; ++r) { if (r == h->rank) a.el_offset = (int64_t)tot; tot += cnt[(size_t)r]; }
            if (tot >= (1ull << 32)) throw Error(CDL_ERR_UNSUPPORTED, "wise = \"element\": more than 2^32 covered entries");
        }
        if (h->p2p_active) {
            // The all-reduce above is also the barrier that makes this safe: every rank has returned from its previous
            // run (its last table kernel, which reads the peers' arenas, is complete).  A run with incremental
            // statistics leaves its carried table in the arena; every run starts from cleared halves.
            CDL_CUDA(cudaMemsetAsync(h->p2p.arena, 0, sizeof(unsigned long long) * 3 * h->p2p.table_elems, h->st));
            CDL_CUDA(cudaStreamSynchronize(h->st));
            allreduce_scalar_d(h, 0.0);      // nobody starts sweeping (and reading a peer's arena) before all are cleared
        }
        a.key = PhiloxKey{(uint32_t)(gp.seed & 0xffffffffu), (uint32_t)(gp.seed >> 32)};
        a.T = T;
        // incremental statistics: per-variant tables (the element path keeps fp64 sums); across GPUs only with the
        // peer-memory exchange, whose arena keeps each rank's own table apart from the sum (the NCCL all-reduce
        // overwrites it in place)
        const bool incr_possible = !element && !replay && (h->world == 1 || h->p2p_active);
        a.incremental = (gp.incremental_stats != 0 && incr_possible) ? 1 : 0;

        // replay arrays -> device
        DevBuf<double> r_ua, r_uc, r_t0, r_t1, r_rr;
        cdl_replay rp_dev{};
        const cdl_replay* rp = nullptr;
        if (replay) {
            const size_t R = (size_t)replay->n_rows;
            auto up = [&](DevBuf<double>& b, const double* src, size_t n) -> const double* {
                if (!src) return nullptr;
                b.alloc(n);
                b.upload(src, n, h->st);
                return b.p;
            };
            rp_dev.n_rows = replay->n_rows;
            rp_dev.u_assign = up(r_ua, replay->u_assign, R * (size_t)M);
            rp_dev.u_config = up(r_uc, replay->u_config, R * (size_t)(N * K));
            rp_dev.theta0 = up(r_t0, replay->theta0, R);
            rp_dev.theta1 = up(r_t1, replay->theta1, R * (size_t)n_el);
            rp_dev.relax_rate = up(r_rr, replay->relax_rate, R);
            rp = &rp_dev;
        }

        // exact mode?
        const double trace_bytes = (double)T * (double)M * (double)K * 8.0;
        bool exact = gp.keep_prob_all == 1 || (gp.keep_prob_all == -1 && trace_bytes <= (double)gp.trace_budget_bytes);
        if (gp.relabel && !exact)
            throw Error(CDL_ERR_UNSUPPORTED, "relabel needs the prob_all trace (keep_prob_all = 1, or a trace that fits "
                                             "trace_budget_bytes)");
        if (gp.relabel && h->world > 1)
            throw Error(CDL_ERR_UNSUPPORTED, "relabel runs on one GPU (the column match of every row is a reduction over all cells)");
        const int check_every = gp.check_every > 0 ? gp.check_every : 100;
        // host callback (verbose printing, user interrupt) every check_every sweeps; never from another thread
        auto notify = [&](int chain, int64_t it, double frac) {
            if (gp.progress && gp.progress(gp.progress_user, (int32_t)chain, (int32_t)it, frac) != 0)
                throw Error(CDL_ERR_INTERRUPTED, "interrupted by the progress callback");
        };
        const int64_t n_buin_planned = (int64_t)std::ceil((double)T * gp.buin_frac);
        DevBuf<unsigned long long> gcounts(2);

        // Streaming mode with the reference's stop rule.  Without prob_all the Geweke test (:580-601) and the
        // burn-in that depends on the final iteration (:604-606,645) are evaluated from column sums of prob and
        // prob^2 over SEGMENTS of trace rows.  The cut rows are the rows where a window of some check begins or
        // ends: floor(.1 n) (end of window A), ceil(.5 n) - 1 (row before window B) and ceil(buin_frac n) - 1
        // (row before the burn-in mean) for every check row n.  The cell kernel accumulates into one pair of running
        // sums; at a cut row they are stored as the segment (previous cut, this cut] and cleared.  A window is then a
        // union of whole segments, ADDED up -- never a difference of two long prefix sums, which loses a window of
        // tiny values behind a prefix that once held large ones (a cell that visited a clone early: found by the
        // saturated-probability test).  Window A only grows, so its segments are folded into one fp64 pair as the
        // checks pass.  About 2-3 segments per check are alive instead of n rows; kept as fp32 by default
        // (snapshot_fp32: each is the sum of its own 10-50 rows, so the rounding is relative to the window's own
        // scale; Z moves by ~1e-6 relative, the posterior means by < 1e-7) for half the memory.  Used when a stop
        // before max_iter is possible, the segments alive together fit the trace budget, and the run takes the plain
        // one-chain-at-a-time branch (wise = global / variant).
        const bool snap32 = gp.snapshot_fp32 != 0;
        struct Seg {
            int64_t r0 = 0, r1 = 0;              // rows r0 + 1 .. r1 (1-based trace rows)
            DevBuf<double> d1, d2; DevBuf<float> f1, f2;
            bool need_s2 = false, stored = false, merged = false;
            int64_t merge_at = 0;                // first check whose window A contains it (0 = none)
            int64_t free_at = 0;                 // last check that needs its buffers
            void release() { d1.release(); d2.release(); f1.release(); f2.release(); }
        };
        std::vector<Seg> segs;                  // in row order
        std::vector<int64_t> win_checks;
        auto rowsA = [](int64_t n) { return (int64_t)std::floor(0.1 * (double)n); };
        auto rowsB = [](int64_t n) { return (int64_t)std::ceil(0.5 * (double)n) - 1; };
        auto rowsU = [&](int64_t n) { return (int64_t)std::ceil((double)n * gp.buin_frac) - 1; };
        bool win = !exact && !element && gp.min_iter < T;
        auto plan_segments = [&]() {
            segs.clear();
            std::vector<int64_t> cuts;
            for (int64_t n : win_checks) {
                const int64_t r3[3] = {rowsA(n), rowsB(n), rowsU(n)};
                for (int q = 0; q < 3; ++q) if (r3[q] > 1 && r3[q] < T) cuts.push_back(r3[q]);   // rows <= 1 hold no sums
            }
            std::sort(cuts.begin(), cuts.end());
            cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
            int64_t prev = 1;
            for (int64_t r : cuts) {
                Seg sg;
                sg.r0 = prev; sg.r1 = r;
                for (int64_t n : win_checks) {
                    if (n < r) continue;
                    const bool inA = r <= rowsA(n), inB = sg.r0 >= rowsB(n), inU = sg.r0 >= rowsU(n);
                    if (inA && !sg.merge_at) sg.merge_at = n;
                    if (inA || inB) sg.need_s2 = true;
                    if (inB || inU) sg.free_at = std::max(sg.free_at, n);
                }
                sg.free_at = std::max(sg.free_at, sg.merge_at);
                segs.push_back(std::move(sg));
                prev = r;
            }
        };
        if (win) {
            for (int64_t n = ((gp.min_iter + check_every - 1) / check_every) * check_every; n <= T; n += check_every)
                win_checks.push_back(n);
            if (win_checks.empty() || win_checks.back() != T) win_checks.push_back(T);
            plan_segments();
            // most arrays alive at once: a segment lives from its cut row to the last check that needs it
            double peak = 0.0;
            for (const Seg& at : segs) {
                double live = 0.0;
                for (const Seg& q : segs)
                    if (q.r1 <= at.r1 && q.free_at >= at.r1) live += q.need_s2 ? 2.0 : 1.0;
                peak = std::max(peak, live);
            }
            // + the running pair, the window-A pair and the two temporaries of a check, all fp64
            if ((peak * (snap32 ? 0.5 : 1.0) + 6.0) * (double)M * (double)K * 8.0 > (double)gp.trace_budget_bytes) {
                win = false; segs.clear();
            }
        }
        DevBuf<double> winA1, winA2, winB1, winB2;      // window-A sums (folded), per-check temporaries
        DevBuf<const void*> seg_ptrs;

        // ---- a chain's life in three steps, each on whatever stream h->st currently is ----
        auto start_chain = [&](int c, SweepArgs& a) -> std::unique_ptr<Chain> {
            // a chain of the previous run, if the handle kept one: its buffers of the same size are taken over as they are
            // (every buffer is initialised below exactly as a new one would be), the others are reallocated
            std::unique_ptr<Chain> chp;
            if (!h->spare.empty()) {
                chp = std::move(h->spare.back());
                h->spare.pop_back();
                chp->dev = ChainDev{};
                chp->theta0_mean = 0.0; chp->relax_mean = 0.0;
                chp->info = cdl_gibbs_info{};
                chp->done = false;
            } else {
                chp.reset(new Chain());
            }
            Chain& ch = *chp;
            ch.scal.alloc(16); ch.scal.zero(h->st);
            ch.l1.alloc((size_t)(element ? 1 : n_el)); ch.l1c.alloc((size_t)(element ? 1 : n_el));
            if (element) {
                ch.el_l1.alloc((size_t)d.c_groups * GROUP); ch.el_l1.zero(h->st);
                ch.el_l1c.alloc((size_t)d.c_groups * GROUP); ch.el_l1c.zero(h->st);
                ch.sd1.alloc((size_t)(N * K)); ch.sd1.zero(h->st);
            } else {
                ch.el_l1.release(); ch.el_l1c.release(); ch.sd1.release();
            }
            ch.mask.alloc((size_t)N);
            ch.z.alloc((size_t)M); ch.z.zero(h->st);
            ch.sab.alloc((size_t)(N * K)); ch.sab.zero(h->st);
            ch.theta0_all.alloc((size_t)T); ch.theta0_all.zero(h->st);
            ch.theta1_all.alloc((size_t)(T * n_el)); ch.theta1_all.zero(h->st);
            ch.loglik_all.alloc((size_t)T); ch.loglik_all.zero(h->st);
            ch.relax_all.alloc((size_t)T);
            {
                std::vector<double> rr((size_t)T, (gp.relax_Config && gp.relax_rate_fixed_set) ? gp.relax_rate_fixed : 0.0);
                ch.relax_all.upload(rr.data(), (size_t)T, h->st);
                CDL_CUDA(cudaStreamSynchronize(h->st));
            }
            ch.config_all.alloc((size_t)(T * N * K)); ch.config_all.zero(h->st);
            // (a buffer this run does not use is released: a taken-over chain may hold one, and the fetch calls and the
            // sweep tell "kept" from "not kept" by the pointer)
            if (exact) { ch.prob_all.alloc((size_t)(T * M * K)); ch.prob_all.zero(h->st); ch.prob_acc.release(); }
            else { ch.prob_acc.alloc((size_t)(M * K)); ch.prob_acc.zero(h->st); ch.prob_all.release(); }
            if (win) { ch.prob_sq.alloc((size_t)(M * K)); ch.prob_sq.zero(h->st); } else ch.prob_sq.release();
            if (gp.keep_assign_all) { ch.assign_all.alloc((size_t)(T * M)); ch.assign_all.zero(h->st); } else ch.assign_all.release();
            const size_t tgrid = (size_t)((N + 7) / 8) + 1;     // k_table blocks: 256 / KP variants each, KP <= 32
            ch.part_d.alloc(tgrid * 4); ch.part_u.alloc(tgrid * 8);
            ch.ticket.alloc(1); ch.ticket.zero(h->st);
            ch.work.alloc(2); ch.work.zero(h->st);
            ch.stats_work.alloc((size_t)d.n_cb + 1); ch.stats_work.zero(h->st);       // + the statistic kernels' ticket
            ch.tab_g.alloc((size_t)N);
            ch.dev.scal = ch.scal.p; ch.dev.l1 = ch.l1.p; ch.dev.l1c = ch.l1c.p; ch.dev.mask = ch.mask.p;
            ch.dev.z = ch.z.p; ch.dev.sab = ch.sab.p;
            ch.dev.theta0_all = ch.theta0_all.p; ch.dev.theta1_all = ch.theta1_all.p;
            ch.dev.loglik_all = ch.loglik_all.p; ch.dev.relax_all = ch.relax_all.p;
            ch.dev.config_all = ch.config_all.p;
            ch.dev.prob_all = exact ? ch.prob_all.p : nullptr;
            ch.dev.assign_all = gp.keep_assign_all ? ch.assign_all.p : nullptr;
            ch.dev.prob_acc = nullptr; ch.dev.prob_sq = nullptr;
            ch.dev.part_d = ch.part_d.p; ch.dev.part_u = ch.part_u.p; ch.dev.ticket = ch.ticket.p;
            ch.dev.tab_g = ch.tab_g.p;
            ch.dev.el_l1 = ch.el_l1.p; ch.dev.el_l1c = ch.el_l1c.p; ch.dev.sd1 = ch.sd1.p;
            ch.dev.work = ch.work.p;
            ch.dev.stats_work = ch.stats_work.p;
            ch.dev.stats_ticket = ch.stats_work.p + d.n_cb;
            ch.chg_ctl.alloc(4); ch.chg_ctl.zero(h->st);
            ch.dev.chg_ctl = ch.chg_ctl.p; ch.dev.chg_list = nullptr; ch.dev.chg_old = nullptr;
            ch.dev.sab_zero = nullptr;
            if (a.incremental) {
                ch.chg_list.alloc((size_t)M); ch.chg_old.alloc((size_t)M);
                ch.dev.chg_list = ch.chg_list.p; ch.dev.chg_old = ch.chg_old.p;
                ch.sab_zero.alloc((size_t)(N * K)); ch.sab_zero.zero(h->st);
                ch.dev.sab_zero = ch.sab_zero.p;
                const unsigned int ctl0[4] = {0u, 1u, 1u, 0u};          // the first sweep takes the full pass
                ch.chg_ctl.upload(ctl0, 4, h->st);
                CDL_CUDA(cudaStreamSynchronize(h->st));
            } else {
                ch.chg_list.release(); ch.chg_old.release(); ch.sab_zero.release();
            }
            ch.dev.dbg_s34 = nullptr; ch.dev.dbg_sab = nullptr;
            ch.dev.tl = nullptr; ch.dev.tl_hist = nullptr;
            ch.tl.release();
            if (const char* e = getenv("CDL_TIMELINE")) {
                if (atoi(e) != 0) {
                    ch.tl.alloc((size_t)2 * 3 * TL_ROWS * TL_SLOTS + 1024); ch.tl.zero(h->st);
                    ch.dev.tl = ch.tl.p;
                    ch.dev.tl_hist = ch.tl.p + (size_t)2 * 3 * TL_ROWS * TL_SLOTS;
                }
            }
            if (h->debug_capture) {
                ch.dbg_s34.alloc((size_t)(2 * N + 5)); ch.dbg_s34.zero(h->st);
                ch.dbg_sab.alloc((size_t)(N * K)); ch.dbg_sab.zero(h->st);
                ch.dev.dbg_s34 = ch.dbg_s34.p; ch.dev.dbg_sab = ch.dbg_sab.p;
            } else {
                ch.dbg_s34.release(); ch.dbg_sab.release();
            }
            ch.dev.sell_scratch = nullptr; ch.dev.sell_tickets = nullptr;
            if (h->use_sell && d.s_slices < 64 * (int64_t)h->sm_count) {
                const int KPs = K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 16 ? 16 : 32));
                ch.sell_scratch.alloc((size_t)d.s_slices * 4 * 32 * KPs);
                ch.sell_tickets.alloc((size_t)d.s_slices); ch.sell_tickets.zero(h->st);
                ch.dev.sell_scratch = ch.sell_scratch.p; ch.dev.sell_tickets = ch.sell_tickets.p;
            } else {
                ch.sell_scratch.release(); ch.sell_tickets.release();
            }

            a.chain = (uint32_t)(gp.chain_offset + c);
            // initial draws (row 1)
            a.it = 1;
            a.rng_it = 1;
            a.learn_relax_next = (a.relax && !a.relax_fixed_set && (2.0 > 0.1 * gp.min_iter + 5.0)) ? 1 : 0;
            if (rp) {
                a.rp_theta0 = rp->theta0; a.rp_theta1 = rp->theta1;
                a.rp_relax_next = rp->relax_rate ? rp->relax_rate + 1 : nullptr;
            }
            launch_init_chain(a, ch.dev, h->st);
            if (element) launch_el_init(h->donor, a, ch.dev, h->st);
            return chp;
        };
        bool win_burn_ready = false;            // windows mode: winB1 holds the sum over the burn-in rows of the final it
        auto finish_chain = [&](Chain& ch, SweepArgs& a, int64_t it_final, double frac) {
            const int64_t itf = it_final;
            if (exact) frac = geweke_fraction(h, ch, itf, gcounts);
            else if (win) { /* frac is the windowed test at itf, from the run loop */ }
            else {
                // streaming mode keeps no prob_all, so the reference's rule (Geweke Z of every prob_all column,
                // R/clone_id.R:594-601) cannot be evaluated; the same Z test on the theta1 traces -- the model
                // parameters the assignments follow from -- is reported instead (informational: streaming runs
                // always do max_iter sweeps)
                launch_geweke(ch.theta1_all.p, n_el, itf, gcounts.p, h->st);
                unsigned long long cnt[2];
                gcounts.download(cnt, 2, h->st);
                frac = cnt[0] ? (double)cnt[1] / (double)cnt[0] : NAN;
            }
            ch.info.n_iter = (int32_t)itf;
            ch.info.percent_converged = r_round3_percent(frac);
            const int64_t n_buin = (exact || win) ? (int64_t)std::ceil((double)itf * gp.buin_frac) : n_buin_planned;
            ch.info.n_buin = (int32_t)n_buin;
            ch.info.n_element = (int32_t)n_el;
            ch.info.exact_trace = exact ? 1 : 0;
            ch.info.stop_rule = exact ? 0 : (win ? 1 : 2);
            ch.info.W_log = a.W_log;
            const int64_t r0 = n_buin - 1, r1 = itf;
            if (gp.relabel) relabel_trace(ch.prob_all.p, ch.config_all.p, M, N, K, n_buin, itf, h->st);   // (:624-644)
            // posterior means (:645-653)
            ch.prob_mean.alloc((size_t)(M * K));
            if (exact) launch_col_means_d(ch.prob_all.p, M * K, r0, r1, ch.prob_mean.p, h->st);
            else if (win) {
                if (!win_burn_ready) throw Error(CDL_ERR_STATE, "windowed stop rule: burn-in sums missing");
                launch_window_mean(winB1.p, nullptr, false, M * K, 1.0 / (double)(r1 - r0), ch.prob_mean.p, h->st);
            }
            else {
                // accumulator holds the sum over rows [n_buin_planned, T]
                CDL_CUDA(cudaMemcpyAsync(ch.prob_mean.p, ch.prob_acc.p, sizeof(double) * (size_t)(M * K),
                                         cudaMemcpyDeviceToDevice, h->st));
                launch_scale(ch.prob_mean.p, M * K, 1.0 / (double)(r1 - r0), h->st);
            }
            ch.config_prob.alloc((size_t)(N * K));
            launch_col_means_u8(ch.config_all.p, N * K, r0, r1, ch.config_prob.p, h->st);
            ch.theta1_mean.alloc((size_t)n_el);
            launch_col_means_d(ch.theta1_all.p, n_el, r0, r1, ch.theta1_mean.p, h->st);
            std::vector<double> t0((size_t)itf), rr((size_t)itf), ll((size_t)itf);
            ch.theta0_all.download(t0.data(), (size_t)itf, h->st);
            ch.relax_all.download(rr.data(), (size_t)itf, h->st);
            ch.loglik_all.download(ll.data(), (size_t)itf, h->st);
            double s0 = 0, sr = 0;
            for (int64_t r = r0; r < r1; ++r) { s0 += t0[(size_t)r]; sr += rr[(size_t)r]; }
            ch.theta0_mean = s0 / (double)(r1 - r0);
            ch.relax_mean = sr / (double)(r1 - r0);
            // logLik_post + DIC (:656-657)
            double llp = loglik_general(d, ch.config_prob.p, K, nullptr, ch.prob_mean.p, ch.theta0_mean,
                                        ch.theta1_mean.p, n_el, h->st);
            if (h->world > 1) llp = allreduce_scalar_d(h, llp);
            ch.info.logLik_post = llp;
            std::vector<double> llw(ll.begin() + r0, ll.begin() + r1);
            dic_from_trace(llw, llp, ch.info.dic);
            ch.done = true;
        };

        // Independent chains (the reference forks one R process per chain, R/clone_id.R:127-154) are interleaved
        // on several streams when they fit in memory together: a small donor's sweep is a handful of tiny
        // kernels, so chains overlap on the GPU instead of queueing behind one another.  Cell-sharded and replay
        // runs keep one chain at a time on the library's stream.
        const double per_chain_bytes = (exact ? trace_bytes : 0.0) + (double)T * (double)(N * K) +
                                       (double)T * (double)n_el * 8.0;
        const bool concurrent_ok = gp.n_chain > 1 && h->world == 1 && !rp &&
                                   per_chain_bytes * gp.n_chain <= (double)gp.trace_budget_bytes;
        // small donors: every chain is one CTA of one grid, check_every sweeps per launch (small.cu)
        const bool small = (h->cells_layout == CDL_CELLS_SMALL ||
                            (h->cells_layout == CDL_CELLS_AUTO && d.M_total < 16384)) &&
                           h->world == 1 && !rp && !element && gp.incremental_stats != 1 && small_path_fits(d, K, gp.wise) &&
                           per_chain_bytes * gp.n_chain <= (double)gp.trace_budget_bytes &&
                           (exact || !win);    // without the trace the stop rule needs the window snapshots (three-kernel path)
        if (small) { win = false; segs.clear(); a.incremental = 0; }
        const bool concurrent = concurrent_ok && !win;     // the windowed stop rule runs one chain at a time
        if (h->cells_layout == CDL_CELLS_SMALL && !small)
            throw Error(CDL_ERR_UNSUPPORTED, "the small-donor path does not apply (size, replay, sharding, element or incremental mode)");
        if (small) {
            // CDL_PHASES=1 (debugging switch, like CDL_TIMELINE): host seconds spent setting the chains up, in the sweep
            // launches and the convergence checks between them, and in the per-chain summaries, on stderr
            static const bool phases = [] { const char* e = getenv("CDL_PHASES"); return e && e[0] == '1'; }();
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto secs = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
                return std::chrono::duration<double>(y - x).count();
            };
            const auto ph0 = now();
            double ph_checks = 0.0;
            std::vector<std::unique_ptr<Chain>> run;
            std::vector<SweepArgs> args((size_t)gp.n_chain, a);
            std::vector<ChainDev> devs((size_t)gp.n_chain);
            for (int c = 0; c < gp.n_chain; ++c) {
                run.push_back(start_chain(c, args[(size_t)c]));
                devs[(size_t)c] = run.back()->dev;
                devs[(size_t)c].prob_acc = exact ? nullptr : run.back()->prob_acc.p;   // the kernel gates it by iteration
            }
            DevBuf<ChainDev> devs_d((size_t)gp.n_chain);
            DevBuf<int> active_d((size_t)gp.n_chain);
            devs_d.upload(devs.data(), devs.size(), h->st);
            std::vector<int> active((size_t)gp.n_chain, 1);
            std::vector<int64_t> it_final((size_t)gp.n_chain, T);
            std::vector<double> fr((size_t)gp.n_chain, NAN);
            int n_active = gp.n_chain;
            int64_t it = 2;
            DevBuf<const double*> gw_ptrs((size_t)gp.n_chain);
            DevBuf<unsigned long long> gw_counts(2 * (size_t)gp.n_chain);
            std::vector<const double*> gw_list;
            std::vector<int> gw_chain;
            std::vector<unsigned long long> gw_host;
            const auto ph1 = now();
            while (it <= T && n_active > 0) {
                // up to the next point at which the reference checks convergence (:580), at most 1000 sweeps a launch
                int64_t end = T;
                if (exact) {
                    int64_t chk = ((std::max<int64_t>(it, gp.min_iter) + check_every - 1) / check_every) * check_every;
                    end = std::min<int64_t>(T, chk);
                }
                end = std::min<int64_t>(end, it + 999);
                active_d.upload(active.data(), active.size(), h->st);
                cdl_small_run r{};
                r.n_chain = gp.n_chain; r.chain0 = gp.chain_offset; r.min_iter = gp.min_iter; r.exact = exact ? 1 : 0;
                r.n_buin_planned = n_buin_planned; r.it_begin = (uint32_t)it; r.it_end = (uint32_t)end;
                r.chains_dev = devs_d.p; r.active_dev = active_d.p;
                launch_small_chains(d, a, r, h->st);
                h->launches += 1;
                if (exact && end >= gp.min_iter && end % check_every == 0) {
                    if (phases) CDL_CUDA(cudaStreamSynchronize(h->st));
                    const auto pc0 = now();
                    // the test of every chain that is still running, in one launch and one read-back
                    gw_list.clear(); gw_chain.clear();
                    for (int c = 0; c < gp.n_chain; ++c)
                        if (active[(size_t)c]) { gw_list.push_back(run[(size_t)c]->dev.prob_all); gw_chain.push_back(c); }
                    gw_ptrs.upload(gw_list.data(), gw_list.size(), h->st);
                    gw_host.assign(2 * gw_list.size(), 0ull);
                    const bool batched = launch_geweke_batch(gw_ptrs.p, (int)gw_list.size(), M * K, end, gw_counts.p, h->st);
                    if (batched) gw_counts.download(gw_host.data(), gw_host.size(), h->st);
                    for (size_t q = 0; q < gw_chain.size(); ++q) {
                        const int c = gw_chain[q];
                        if (batched) fr[(size_t)c] = gw_host[2 * q] ? (double)gw_host[2 * q + 1] / (double)gw_host[2 * q] : NAN;
                        else fr[(size_t)c] = geweke_fraction(h, *run[(size_t)c], end, gcounts);      // wide traces: one by one
                        notify(gp.chain_offset + c, end, fr[(size_t)c]);
                        if (fr[(size_t)c] > 0.995) { it_final[(size_t)c] = end; active[(size_t)c] = 0; --n_active; }
                    }
                    ph_checks += secs(pc0, now());
                } else {
                    notify(gp.chain_offset, end, NAN);
                }
                it = end + 1;
            }
            CDL_CUDA(cudaStreamSynchronize(h->st));
            const auto ph2 = now();
            for (int c = 0; c < gp.n_chain; ++c) {
                finish_chain(*run[(size_t)c], args[(size_t)c], it_final[(size_t)c], fr[(size_t)c]);
                h->chains.push_back(std::move(run[(size_t)c]));
            }
            if (phases)
                std::fprintf(stderr, "[cdl phases] %d chains: set-up %.4f s, sweeps %.4f s (+ checks %.4f s), summaries %.4f s\n",
                             gp.n_chain, secs(ph0, ph1), secs(ph1, ph2) - ph_checks, ph_checks, secs(ph2, now()));
        } else if (!concurrent) {
            for (int c = 0; c < gp.n_chain; ++c) {
                std::unique_ptr<Chain> chp = start_chain(c, a);
                Chain& ch = *chp;
                int64_t it_final = T;
                double frac = NAN;
                if (win) {              // every chain walks the same segment schedule
                    plan_segments();
                    const size_t cols = (size_t)(M * K);
                    if (!winA1.p) { winA1.alloc(cols); winA2.alloc(cols); winB1.alloc(cols); winB2.alloc(cols); seg_ptrs.alloc(2 * segs.size() + 4); }
                    winA1.zero(h->st); winA2.zero(h->st);
                }
                size_t next_check = 0;
                win_burn_ready = false;
                // sum of the stored segments whose rows lie after `from_row`, plus the running pair -> winB1 (and winB2)
                auto sum_after = [&](int64_t from_row, bool squares) {
                    std::vector<const void*> p1, p2;
                    for (const Seg& q : segs) {
                        if (!q.stored || q.r0 < from_row) continue;
                        p1.push_back(snap32 ? (const void*)q.f1.p : (const void*)q.d1.p);
                        if (squares) {
                            if (!q.need_s2) throw Error(CDL_ERR_STATE, "windowed stop rule: a segment lacks its squares");
                            p2.push_back(snap32 ? (const void*)q.f2.p : (const void*)q.d2.p);
                        }
                    }
                    std::vector<const void*> all(p1);
                    all.insert(all.end(), p2.begin(), p2.end());
                    if (!all.empty()) {
                        CDL_CUDA(cudaMemcpyAsync(seg_ptrs.p, all.data(), sizeof(void*) * all.size(), cudaMemcpyHostToDevice, h->st));
                        CDL_CUDA(cudaStreamSynchronize(h->st));      // `all` is a stack vector
                    }
                    launch_segment_sum(seg_ptrs.p, (int)p1.size(), snap32, false, ch.prob_acc.p, M * K, winB1.p, h->st);
                    if (squares) launch_segment_sum(seg_ptrs.p + p1.size(), (int)p2.size(), snap32, true, ch.prob_sq.p, M * K, winB2.p, h->st);
                };
                size_t next_seg = 0;
                for (int64_t it = 2; it <= T; ++it) {
                    if (win) { ch.dev.prob_acc = ch.prob_acc.p; ch.dev.prob_sq = ch.prob_sq.p; }
                    else ch.dev.prob_acc = (!exact && it >= n_buin_planned) ? ch.prob_acc.p : nullptr;
                    one_sweep(h, ch, a, (uint32_t)it, rp, n_el);
                    if (exact && it >= gp.min_iter && it % check_every == 0) {
                        frac = geweke_fraction(h, ch, it, gcounts);
                        notify(gp.chain_offset + c, it, frac);
                        if (frac > 0.995) { it_final = it; break; }
                    } else if (it % check_every == 0 && !(win && it >= gp.min_iter)) {
                        notify(gp.chain_offset + c, it, NAN);
                    }
                    if (!win) continue;
                    if (next_seg < segs.size() && segs[next_seg].r1 == it) {
                        // a cut row: the running sums become the segment (previous cut, it] and start again from zero
                        Seg& sg = segs[next_seg++];
                        const size_t cols = (size_t)(M * K);
                        if (sg.free_at == 0) {          // rows no window of any check covers: just start again from zero
                            launch_segment_store(ch.prob_acc.p, ch.prob_sq.p, nullptr, nullptr, snap32, M * K, h->st);
                        } else if (snap32) {
                            sg.f1.alloc(cols);
                            if (sg.need_s2) sg.f2.alloc(cols);
                            launch_segment_store(ch.prob_acc.p, ch.prob_sq.p, sg.f1.p, sg.need_s2 ? sg.f2.p : nullptr, true, M * K, h->st);
                        } else {
                            sg.d1.alloc(cols);
                            if (sg.need_s2) sg.d2.alloc(cols);
                            launch_segment_store(ch.prob_acc.p, ch.prob_sq.p, sg.d1.p, sg.need_s2 ? sg.d2.p : nullptr, false, M * K, h->st);
                        }
                        sg.stored = sg.free_at != 0;
                    }
                    if (next_check < win_checks.size() && it == win_checks[next_check]) {
                        ++next_check;
                        // window A only grows: fold the segments it has reached into the fp64 pair
                        for (Seg& q : segs) {
                            if (!q.stored || q.merged || q.r1 > rowsA(it)) continue;
                            launch_segment_add(winA1.p, snap32 ? (const void*)q.f1.p : (const void*)q.d1.p, snap32, false, M * K, h->st);
                            launch_segment_add(winA2.p, snap32 ? (const void*)q.f2.p : (const void*)q.d2.p, snap32, true, M * K, h->st);
                            q.merged = true;
                        }
                        sum_after(rowsB(it), true);
                        launch_geweke_sums(winA1.p, winA2.p, nullptr, nullptr, false, winB1.p, winB2.p, M * K, it, gcounts.p, h->st);
                        if (h->world > 1) allreduce_u64(h, gcounts.p, 2);
                        unsigned long long cnt[2];
                        gcounts.download(cnt, 2, h->st);
                        frac = cnt[0] ? (double)cnt[1] / (double)cnt[0] : NAN;
                        const bool rule = it >= gp.min_iter && it % check_every == 0;     // (:580); else the final row
                        if (rule) notify(gp.chain_offset + c, it, frac);
                        if ((rule && frac > 0.995) || it == T) {
                            it_final = it;
                            sum_after(rowsU(it), false);          // burn-in rows ceil(it * buin_frac) .. it (:604,645)
                            win_burn_ready = true;
                            break;
                        }
                        for (Seg& q : segs)            // segments no later check needs
                            if (q.stored && q.free_at <= it) { q.release(); q.stored = false; }
                    }
                }
                finish_chain(ch, a, it_final, frac);
                h->chains.push_back(std::move(chp));
            }
        } else {
            struct Live { std::unique_ptr<Chain> ch; SweepArgs a; cudaStream_t st; int64_t it_final; double frac; bool active; };
            // the host thread issues ~5 launches per chain sweep, so beyond a few streams the run is bound by
            // launch rate, not by the GPU (measured on example_donor: 1 chain 24k, 4 chains 49k chain-iterations/s,
            // 16 streams no better); capturing a sweep in a CUDA graph needs the iteration state on the device
            // and is the next step
            const int n_streams = std::min(gp.n_chain, 4);
            std::vector<cudaStream_t> streams((size_t)n_streams, nullptr);
            cudaStream_t home = h->st;
            std::vector<Live> live;
            try {
                for (auto& s : streams) CDL_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
                CDL_CUDA(cudaStreamSynchronize(home));          // uploads made on the library's stream are done
                for (int c = 0; c < gp.n_chain; ++c) {
                    Live lv{nullptr, a, streams[(size_t)(c % n_streams)], T, NAN, true};
                    h->st = lv.st;
                    lv.ch = start_chain(c, lv.a);
                    live.push_back(std::move(lv));
                }
                int n_active = gp.n_chain;
                for (int64_t it = 2; it <= T && n_active > 0; ++it) {
                    for (Live& lv : live) {
                        if (!lv.active) continue;
                        h->st = lv.st;
                        lv.ch->dev.prob_acc = (!exact && it >= n_buin_planned) ? lv.ch->prob_acc.p : nullptr;
                        one_sweep(h, *lv.ch, lv.a, (uint32_t)it, nullptr, n_el);
                    }
                    if (exact && it >= gp.min_iter && it % check_every == 0) {
                        for (Live& lv : live) {
                            if (!lv.active) continue;
                            h->st = lv.st;
                            lv.frac = geweke_fraction(h, *lv.ch, it, gcounts);
                            notify((int)lv.a.chain, it, lv.frac);
                            if (lv.frac > 0.995) { lv.it_final = it; lv.active = false; --n_active; }
                        }
                    } else if (it % check_every == 0) {
                        notify(gp.chain_offset, it, NAN);
                    }
                }
                for (Live& lv : live) {
                    h->st = lv.st;
                    finish_chain(*lv.ch, lv.a, lv.it_final, lv.frac);
                    CDL_CUDA(cudaStreamSynchronize(lv.st));
                    h->chains.push_back(std::move(lv.ch));
                }
            } catch (...) {
                h->st = home;
                cudaDeviceSynchronize();
                live.clear();
                for (auto s : streams) if (s) cudaStreamDestroy(s);
                throw;
            }
            h->st = home;
            for (auto s : streams) if (s) cudaStreamDestroy(s);
        }
        if (h->p2p_active) {
            unsigned int perr = 0;
            CDL_CUDA(cudaMemcpyAsync(&perr, h->p2p.error, sizeof(perr), cudaMemcpyDeviceToHost, h->st));
            CDL_CUDA(cudaStreamSynchronize(h->st));
            if (perr) throw Error(CDL_ERR_COMM, "a peer rank's statistic table did not arrive (peer-memory exchange timed out)");
        }
        a.rp_u_assign = a.rp_u_config = a.rp_theta0 = a.rp_theta1 = a.rp_relax_next = nullptr;
        a.p2p = nullptr;
        h->sweep = a;
        h->have_sweep = true;
        CDL_CUDA(cudaStreamSynchronize(h->st));
    });
}

int cdl_gibbs_info_get(cdl_handle h, int32_t chain, cdl_gibbs_info* out) {
    return guard(h, [&] { *out = get_chain(h, chain).info; });
}

int cdl_incremental_full_passes(cdl_handle h, int32_t chain, uint32_t* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        unsigned int ctl[4] = {0, 0, 0, 0};
        ch.chg_ctl.download(ctl, 4, h->st);
        *out = ctl[3];
    });
}

int cdl_debug_capture(cdl_handle h, int32_t enable) {
    return guard(h, [&] { h->debug_capture = enable != 0; });
}

int cdl_debug_fetch(cdl_handle h, uint64_t* SA, uint64_t* SB, uint64_t* S3, uint64_t* S4, uint64_t* totals5) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, 0);
        if (!ch.dbg_sab.p) throw Error(CDL_ERR_STATE, "cdl_debug_capture was not enabled before the run");
        const int64_t N = h->donor.N, K = h->K;
        std::vector<unsigned long long> t((size_t)(N * K)), s((size_t)(2 * N + 5));
        ch.dbg_sab.download(t.data(), t.size(), h->st);
        ch.dbg_s34.download(s.data(), s.size(), h->st);
        for (int64_t q = 0; q < N * K; ++q) {
            if (SA) SA[q] = (uint64_t)(uint32_t)t[(size_t)q];
            if (SB) SB[q] = (uint64_t)(t[(size_t)q] >> 32);
        }
        for (int64_t i = 0; i < N; ++i) {
            if (S3) S3[i] = s[(size_t)i];
            if (S4) S4[i] = s[(size_t)(N + i)];
        }
        if (totals5) for (int q = 0; q < 5; ++q) totals5[q] = s[(size_t)(2 * N + q)];
    });
}

int cdl_fetch_prob(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        const int64_t M = h->donor.M, K = h->K;
        ScratchUse t(h, (size_t)(M * K));                            // [cell][clone] -> R's column-major M x K
        launch_transpose_scale(ch.prob_mean.p, M, K, 1.0, t.p, h->st);
        CDL_CUDA(cudaMemcpyAsync(out, t.p, sizeof(double) * (size_t)(M * K), cudaMemcpyDeviceToHost, h->st));
        CDL_CUDA(cudaStreamSynchronize(h->st));
    });
}

int cdl_fetch_config_prob(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        ch.config_prob.download(out, (size_t)(h->donor.N * h->K), h->st);
    });
}

int cdl_fetch_theta(cdl_handle h, int32_t chain, double* theta0, double* theta1_elems) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        if (theta0) *theta0 = ch.theta0_mean;
        if (theta1_elems) ch.theta1_mean.download(theta1_elems, (size_t)h->n_el, h->st);
    });
}

int cdl_fetch_theta_traces(cdl_handle h, int32_t chain, double* theta0_all, double* theta1_all) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        const size_t n = (size_t)ch.info.n_iter;
        if (theta0_all) ch.theta0_all.download(theta0_all, n, h->st);
        if (theta1_all) {
            // R shape n_iter x n_element column-major
            ScratchUse t(h, n * (size_t)h->n_el);
            launch_transpose_scale(ch.theta1_all.p, (int64_t)n, h->n_el, 1.0, t.p, h->st);
            CDL_CUDA(cudaMemcpyAsync(theta1_all, t.p, sizeof(double) * n * (size_t)h->n_el, cudaMemcpyDeviceToHost, h->st));
            CDL_CUDA(cudaStreamSynchronize(h->st));
        }
    });
}

int cdl_fetch_loglik_trace(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        ch.loglik_all.download(out, (size_t)ch.info.n_iter, h->st);
    });
}

int cdl_fetch_relax_rate(cdl_handle h, int32_t chain, double* mean_out, double* all_out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        if (mean_out) *mean_out = ch.relax_mean;
        if (all_out) ch.relax_all.download(all_out, (size_t)ch.info.n_iter, h->st);
    });
}

int cdl_fetch_prob_all(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        if (!ch.info.exact_trace) throw Error(CDL_ERR_STATE, "prob_all was not kept (streaming mode)");
        const int64_t M = h->donor.M, K = h->K, n = ch.info.n_iter;
        std::vector<double> row((size_t)(M * K));
        for (int64_t t = 0; t < n; ++t) {
            ch.prob_all.download(row.data(), row.size(), h->st, (size_t)(t * M * K));
            double* o = out + t * M * K;
            for (int64_t j = 0; j < M; ++j)
                for (int64_t k = 0; k < K; ++k) o[k * M + j] = row[(size_t)(j * K + k)];
        }
    });
}

int cdl_fetch_config_all(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        const size_t n = (size_t)ch.info.n_iter * (size_t)(h->donor.N * h->K);
        std::vector<uint8_t> t(n);
        ch.config_all.download(t.data(), n, h->st);
        for (size_t q = 0; q < n; ++q) out[q] = (double)t[q];
    });
}

int cdl_fetch_assign_all(cdl_handle h, int32_t chain, int32_t* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        if (!ch.assign_all.p) throw Error(CDL_ERR_STATE, "assign_all was not kept (keep_assign_all = 0)");
        const size_t n = (size_t)ch.info.n_iter * (size_t)h->donor.M;
        std::vector<uint8_t> t(n);
        ch.assign_all.download(t.data(), n, h->st);
        for (size_t q = 0; q < n; ++q) out[q] = (int32_t)t[q];
    });
}

int cdl_fetch_prob_variant(cdl_handle h, int32_t chain, double* out) {
    return guard(h, [&] {
        Chain& ch = get_chain(h, chain);
        const Donor& d = h->donor;
        const double bytes = (double)d.N * (double)d.M * 8.0;
        if (bytes > 16e9) throw Error(CDL_ERR_UNSUPPORTED, "dense prob_variant would exceed 16 GB");
        DevBuf<double> pv((size_t)(d.N * d.M));
        launch_prob_variant(d, h->gp.wise, ch.theta0_all.p, ch.theta1_all.p, h->n_el, ch.info.n_buin - 1,
                            ch.info.n_iter, pv.p, h->st);
        pv.download(out, (size_t)(d.N * d.M), h->st);
    });
}

int cdl_fetch_element_index(cdl_handle h, int64_t* out) {
    return guard(h, [&] {
        require_data(h);
        const Donor& d = h->donor;
        std::vector<int64_t> p((size_t)d.M + 1);
        std::vector<int32_t> vi((size_t)d.nnz);
        std::vector<uint16_t> a((size_t)d.nnz), dd((size_t)d.nnz);
        donor_export(d, p.data(), vi.data(), a.data(), dd.data(), h->st);
        for (int64_t j = 0; j < d.M; ++j)
            for (int64_t e = p[(size_t)j]; e < p[(size_t)j + 1]; ++e) out[e] = vi[(size_t)e] + j * d.N;
    });
}

int cdl_get_loglik(cdl_handle h, const double* Config, int32_t K, const int32_t* assign, const double* assign_prob,
                   double theta0, const double* theta1, int64_t n_theta1, double* out) {
    return guard(h, [&] {
        require_data(h);
        const Donor& d = h->donor;
        if (!Config || !theta1 || !out || K < 1) throw Error(CDL_ERR_INVALID, "bad arguments");
        if ((assign == nullptr) == (assign_prob == nullptr))
            throw Error(CDL_ERR_INVALID, "give exactly one of assign / assign_prob");
        if (n_theta1 != 1 && n_theta1 != d.N && n_theta1 != d.nnz)
            throw Error(CDL_ERR_INVALID, "theta1 must have 1, N or nnz values");
        if (n_theta1 != 1 && n_theta1 != d.N) donor_build_elem0(h->donor, h->st);
        DevBuf<double> cfg((size_t)(d.N * K)), th1((size_t)n_theta1), pr;
        DevBuf<int32_t> as;
        cfg.upload(Config, (size_t)(d.N * K), h->st);
        th1.upload(theta1, (size_t)n_theta1, h->st);
        std::vector<double> t;
        if (assign) {
            for (int64_t j = 0; j < d.M; ++j)
                if (assign[j] < 1 || assign[j] > K) throw Error(CDL_ERR_INVALID, "assignment label out of range");
            as.alloc((size_t)d.M);
            as.upload(assign, (size_t)d.M, h->st);
        } else {
            t.resize((size_t)(d.M * K));
            for (int64_t j = 0; j < d.M; ++j)
                for (int64_t k = 0; k < K; ++k) t[(size_t)(j * K + k)] = assign_prob[j + k * d.M];
            pr.alloc(t.size());
            pr.upload(t.data(), t.size(), h->st);
        }
        double v = loglik_general(d, cfg.p, K, as.p, pr.p, theta0, th1.p, n_theta1, h->st);
        if (h->world > 1) v = allreduce_scalar_d(h, v);
        *out = v;
    });
}

int cdl_em_run(cdl_handle h, const double* Config, int32_t K, const double* Psi, int32_t min_iter, int32_t max_iter,
               double logLik_threshold, const double* theta_init, uint64_t seed, double* theta_out, double* prob_out,
               double* logLik_out, int32_t* n_iter_out) {
    return guard(h, [&] {
        require_data(h);
        const Donor& d = h->donor;
        if (!Config || K < 1 || K > KMAX) throw Error(CDL_ERR_UNSUPPORTED, "number of clones must be in 1..32");
        if (d.max_cell_depth >= (1ull << 32)) throw Error(CDL_ERR_UNSUPPORTED, "a cell's total depth exceeds 2^32");
        const int64_t M = d.M;
        std::vector<uint32_t> masks = config_masks(Config, d.N, K, true);
        DevBuf<uint32_t> cfg((size_t)d.N), S3((size_t)(M * K)), S4((size_t)(M * K)), TA((size_t)M), TB((size_t)M);
        cfg.upload(masks.data(), (size_t)d.N, h->st);
        double logPsi[KMAX];
        fill_logpsi(Psi, K, logPsi);
        DevBuf<double> lp(KMAX), prob((size_t)(M * K)), sums(5), partials((size_t)(((M + 255) / 256) * 5));
        lp.upload(logPsi, KMAX, h->st);
        em_stats(d, cfg.p, K, S3.p, S4.p, TA.p, TB.p, h->sm_count, h->st);
        const double W_log = allreduce_scalar_d(h, d.W_log);
        double theta[2];
        if (theta_init) { theta[0] = theta_init[0]; theta[1] = theta_init[1]; }
        else {
            PhiloxKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
            theta[0] = 0.001 + philox_uniform(key, 0, 0, STREAM_EM_INIT, 0) * (0.25 - 0.001);
            theta[1] = 0.3 + philox_uniform(key, 0, 0, STREAM_EM_INIT, 1) * (0.6 - 0.3);
        }
        double logLik = 0.0, logLik_new = logLik + logLik_threshold * 2;
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            if (it > min_iter && (logLik_new - logLik) < logLik_threshold) break;
            logLik = logLik_new;
            em_iteration(M, K, S3.p, S4.p, TA.p, TB.p, lp.p, theta[0], theta[1], prob.p, sums.p, partials.p, h->st);
            double s[5];
            if (h->world > 1)
                nccl_check(h, h->nccl.AllReduce(sums.p, sums.p, 5, NCCL_F64, NCCL_SUM, h->comm, h->st), "ncclAllReduce");
            sums.download(s, 5, h->st);
            logLik_new = s[0] + W_log;
            theta[0] = (s[2] != s[2] || s[2] == 0.0) ? 0.02 : s[1] / s[2];
            theta[1] = (s[4] != s[4] || s[4] == 0.0) ? 0.75 : s[3] / s[4];
        }
        if (it > max_iter) it = max_iter;   // R's `it` after a completed for loop is the last value
        if (theta_out) { theta_out[0] = theta[0]; theta_out[1] = theta[1]; }
        if (logLik_out) *logLik_out = logLik;
        if (n_iter_out) *n_iter_out = it;
        if (prob_out) {
            std::vector<double> t((size_t)(M * K));
            prob.download(t.data(), t.size(), h->st);
            for (int64_t j = 0; j < M; ++j)
                for (int64_t k = 0; k < K; ++k) prob_out[j + k * M] = t[(size_t)(j * K + k)];
        }
    });
}

int cdl_bench_sweeps(cdl_handle h, int32_t warmup, int32_t n, float* ms_total, float* ms_cells, float* ms_stats,
                     float* ms_table, int64_t* launches) {
    return guard(h, [&] {
        if (!h->have_sweep || h->chains.empty()) throw Error(CDL_ERR_STATE, "run cdl_gibbs_run first");
        Chain& ch = *h->chains[0];
        SweepArgs a = h->sweep;
        a.chain = 0;
        const uint32_t it = (uint32_t)h->T;     // keep overwriting the last trace row
        ch.dev.prob_acc = nullptr; ch.dev.prob_sq = nullptr;
        cudaEvent_t e0, e1;
        CDL_CUDA(cudaEventCreate(&e0));
        CDL_CUDA(cudaEventCreate(&e1));
        // warm-up and timed sweeps go down the stream back to back; the start event sits between them
        uint32_t adv = 0;                        // fresh Philox streams for every bench sweep (the trace row is reused)
        for (int w = 0; w < warmup; ++w) one_sweep(h, ch, a, it, nullptr, h->n_el, ++adv);
        const int64_t l0 = h->launches;
        CDL_CUDA(cudaEventRecord(e0, h->st));
        for (int q = 0; q < n; ++q) one_sweep(h, ch, a, it, nullptr, h->n_el, ++adv);
        CDL_CUDA(cudaEventRecord(e1, h->st));
        CDL_CUDA(cudaEventSynchronize(e1));
        float tot = 0;
        CDL_CUDA(cudaEventElapsedTime(&tot, e0, e1));
        if (ms_total) *ms_total = tot;
        if (launches) *launches = h->launches - l0;
        if (ch.dev.tl && n > 0) {
            print_timeline(h, ch);
            // duration of every timed sweep: end of its table kernel minus the end of the previous one
            std::vector<unsigned long long> hist(1024);
            CDL_CUDA(cudaMemcpy(hist.data(), ch.dev.tl_hist, 8 * 1024, cudaMemcpyDeviceToHost));
            std::string line = "[cdl sweeps r" + std::to_string(h->rank) + "] us per sweep (warm-up, then timed):";
            const uint32_t first = it + adv - (uint32_t)(warmup + n) + 1;
            for (uint32_t q = first + 1; q <= it + adv && q - first < 200; ++q) {
                const unsigned long long a1 = hist[q & 1023u], a0 = hist[(q - 1) & 1023u];
                char buf[32];
                std::snprintf(buf, sizeof buf, "%s%.1f", (q - first == (uint32_t)warmup) ? " | " : " ", a1 && a0 ? ((double)a1 - (double)a0) * 1e-3 : -1.0);
                line += buf;
            }
            std::fprintf(stderr, "%s\n", line.c_str());
        }
        // per-kernel breakdown (separate pass so the events do not perturb the total)
        float tc = 0, ts = 0, tt = 0;
        if (ms_cells || ms_stats || ms_table) {
            cudaEvent_t ev[4];
            for (auto& e : ev) CDL_CUDA(cudaEventCreate(&e));
            for (int q = 0; q < n; ++q) {
                a.it = it;
                a.rng_it = it + (++adv);
                a.learn_relax_next = (a.relax && !a.relax_fixed_set && ((double)(it + 1) > 0.1 * h->gp.min_iter + 5.0)) ? 1 : 0;
                CDL_CUDA(cudaEventRecord(ev[0], h->st));
                if (h->use_sell) launch_cells_sell(h->donor, a, ch.dev, h->sm_count, h->st);
                else launch_cells(h->donor, a, ch.dev, h->sm_count, h->st);
                CDL_CUDA(cudaEventRecord(ev[1], h->st));
                if (h->p2p_active) {
                    P2PState& q = h->p2p;
                    ++q.epoch;
                    ch.dev.sab = q.arena + (size_t)(q.epoch & 1ull) * q.table_elems;
                    a.p2p = &q;
                    launch_stats(h->donor, a, ch.dev, h->sm_count, h->st);
                } else {
                    a.p2p = nullptr;
                    launch_stats(h->donor, a, ch.dev, h->sm_count, h->st);
                    allreduce_u64(h, ch.dev.sab, (size_t)(a.N * a.K));
                }
                CDL_CUDA(cudaEventRecord(ev[2], h->st));
                launch_table(h->donor, a, ch.dev, h->st);
                CDL_CUDA(cudaEventRecord(ev[3], h->st));
                CDL_CUDA(cudaEventSynchronize(ev[3]));
                float x;
                CDL_CUDA(cudaEventElapsedTime(&x, ev[0], ev[1])); tc += x;
                CDL_CUDA(cudaEventElapsedTime(&x, ev[1], ev[2])); ts += x;
                CDL_CUDA(cudaEventElapsedTime(&x, ev[2], ev[3])); tt += x;
            }
            for (auto& e : ev) cudaEventDestroy(e);
        }
        if (ms_cells) *ms_cells = tc;
        if (ms_stats) *ms_stats = ts;
        if (ms_table) *ms_table = tt;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

}  // extern "C"
