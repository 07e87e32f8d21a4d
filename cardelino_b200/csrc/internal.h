// Internal host-side declarations shared by the translation units of libcardelino_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <stdexcept>
#include "common.cuh"

namespace cdl {

constexpr int GROUP = 8;           // entries per 128-bit load of uint16 payload
constexpr int CELL_BLOCK = 65536;  // cells per variant-major block (local cell id fits uint16)
constexpr int KMAX = 32;
constexpr int MAX_PEERS = 7;        // other ranks of one NVSwitch box

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CDL_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::cdl::Error(2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                                      __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    } while (0)

// Launch of a sweep kernel with programmatic stream serialization (programmatic dependent launch): the CTAs of
// kernel n+1 are placed on the SMs while kernel n drains and block in pdl_enter() (common.cuh) until kernel n has
// completed and its writes are visible, which takes the launch latency and CTA start-up out of the gap between
// the three kernels of a sweep.  Every kernel launched this way calls pdl_enter() before its first global access.
// CDL_PDL=0 in the environment launches them the ordinary way.
bool pdl_enabled();
template <typename P>
inline void launch_dep(void (*kern)(const P), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const P& p) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CDL_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
}

// Device-resident donor: the same counts in two layouts.
//  (A) cell-major, each cell's entries padded to a multiple of GROUP with (variant 0, a 0, b 0):
//      c_rowgrp[M+1] group offsets; c_vidx (uint16 if N <= 65536 else uint32), c_a, c_b.
//      One pass over (A) gives the thin SpMM logLik  (R/clone_id.R:470-478).
//  (B) variant-major inside blocks of CELL_BLOCK cells, segments padded likewise:
//      v_seggrp[n_cb*N+1]; v_cell (uint16 local cell id), v_a, v_b.
//      One pass over (B) gives the per-variant x clone read sums that replace the S-lists
//      (R/clone_id.R:394-403,537-562).
struct Donor {
    int64_t N = 0, M = 0, nnz = 0;      // M = local cells of this rank
    int64_t M_total = 0, cell_offset = 0;
    bool vidx32 = false;
    uint32_t* c_rowgrp = nullptr;
    void* c_vidx = nullptr;
    uint16_t* c_a = nullptr;
    uint16_t* c_b = nullptr;
    int64_t c_groups = 0;
    int64_t* c_elem0 = nullptr;         // [M+1] element index (covered entries in cell-major order) of each cell's first entry; built on first use
    int n_cb = 0;
    uint32_t* v_seggrp = nullptr;
    int* v_plan = nullptr;              // k_stats launch plan (CTA -> cell block), built on first use
    int v_plan_ctas = 0;
    uint16_t* v_flush = nullptr;        // [n_cb*N] groups a lane may bin before a 16+16-bit bin could overflow (k_stats)
    uint16_t* v_cell = nullptr;
    uint16_t* v_a = nullptr;
    uint16_t* v_b = nullptr;
    int64_t v_groups = 0;
    //  (C) slice-major copy of (A) for large donors (cells_sell.cu): slices of 32 length-sorted cells,
    //      one lane per cell; built lazily by donor_build_sell.
    uint32_t* s_off = nullptr;          // [s_slices + 1] group-slot offsets
    uint32_t* s_perm = nullptr;         // [s_slices * 32] local cell id per lane (0xffffffff = empty)
    void* s_vidx = nullptr;
    uint16_t* s_a = nullptr;
    uint16_t* s_b = nullptr;
    int32_t* s_assign = nullptr;        // one-round shards: slice of every (CTA, warp) of the cell kernel, built on first use
    int s_assign_grid = 0, s_assign_warps = 0;
    int64_t s_slices = 0, s_slots = 0;
    int s_hi_shift = 0;                 // odd entries of a slice group store variant id << s_hi_shift (2 when N <= 13824)
    double W_log = 0.0;                 // local shard's sum(lchoose(D, A))
    int32_t* labels = nullptr;          // synthetic ground truth (1-based), optional
    uint64_t max_variant_depth = 0;     // max_i sum_j D_ij (must stay < 2^32 for packed statistics)
    uint64_t max_cell_depth = 0;
};

void donor_free(Donor& d);
void donor_free_sell(Donor& d);
void donor_build_sell(Donor& d, cudaStream_t st);
void donor_build_elem0(Donor& d, cudaStream_t st);
// raw CSC (cells = columns) already on the device -> both layouts.  Frees nothing of the input.
void donor_build(Donor& d, int64_t N, int64_t M, int64_t nnz, const int64_t* p_dev, const int32_t* i_dev,
                 const uint16_t* a_dev, const uint16_t* d_dev, cudaStream_t st);
void donor_export(const Donor& d, int64_t* p, int32_t* i, uint16_t* a, uint16_t* dd, cudaStream_t st);
void synth_generate(Donor& d, int64_t N, int64_t M_total, int64_t cell_offset, int64_t M_local, int K,
                    const double* Config_host, double coverage, uint64_t seed, cudaStream_t st);

// Per-chain sampler state (device pointers).
struct ChainDev {
    // current state
    double* scal;        // [16] device scalars, see SC_* below
    double* l1;          // [n_tab] log(theta1), n_tab = N (variant) or 1 (global)
    double* l1c;         // [n_tab] log(1 - theta1)
    uint32_t* mask;      // [N] Config_new as clone bit masks
    uint8_t* z;          // [M] current labels (0-based)
    unsigned long long* sab;  // [N*K] packed statistics: low 32 bits sum A, high 32 bits sum B
    // traces (row t = iteration t+1)
    double* theta0_all;  // [T]
    double* theta1_all;  // [T * n_el]
    double* loglik_all;  // [T]
    double* relax_all;   // [T]
    uint8_t* config_all; // [T * N*K], element i + k*N like R's column-major flattening
    double* prob_all;    // [T * M*K] device layout [t][cell][clone]  (exact mode) or nullptr
    uint8_t* assign_all; // [T * M] 1-based labels or nullptr
    double* prob_acc;    // [M*K] streaming accumulator (compact mode) or nullptr
    double* prob_sq;     // [M*K] running sum of squares beside it (streaming Geweke windows) or nullptr
    // reduction scratch
    double* part_d;      // [grid * 4]
    unsigned long long* part_u; // [grid * 8]
    unsigned int* ticket;
    double* el_l1;       // wise = element: ln theta1 of every entry of the row layout (padded positions are 0)
    double* el_l1c;      //                 ln(1 - theta1)
    double* sd1;         //                 [N*K] per-(variant, clone) sums of A ln th1 + B ln(1 - th1)
    double2* tab_g;      // (du, dv) table in global memory when it does not fit shared memory, else nullptr
    double* sell_scratch;        // [s_slices][4][32][KP] partial logits of split slices (small shards) or nullptr
    unsigned int* sell_tickets;  // [s_slices]
    // incremental statistics (opt-in): cells whose label changed in this sweep, and control words
    uint32_t* chg_list;          // [M] local cell ids
    uint8_t* chg_old;            // [M] their previous labels
    unsigned int* chg_ctl;       // [4]: changed count, last sweep took the full pass, force-full flag, full passes done
    unsigned long long* sab_zero;// [N*K] all zero at rest: where a full pass accumulates when the table is carried over
    unsigned long long* dbg_s34;  // test hook (cdl_debug_capture): [2N + 5] Beta shape increments of the last sweep, or nullptr
    unsigned long long* dbg_sab;  // test hook: [N*K] copy of the local statistic table taken before the table kernel, or nullptr
    unsigned int* work;  // dynamic work-item counter of k_cells_sell
    unsigned int* stats_work;    // [n_cb] dynamic unit counters of k_stats
    unsigned int* stats_ticket;  // "last CTA" ticket of the statistic kernels (zero between launches)
    unsigned long long* tl;      // device timeline of the last sweep (CDL_TIMELINE=1), else nullptr
    unsigned long long* tl_hist; // [1024] end of the table kernel of sweep (rng_it & 1023), else nullptr
};

enum {
    SC_THETA0 = 0, SC_L0 = 1, SC_L0C = 2, SC_RELAX = 3,   // relax rate used by the NEXT flip step
    SC_ODD1 = 4,         // log(1 - r) - log(r) of that relax rate: the prior log-odds of an entry whose input Config is 1 (an
                         // input 0 has exactly the negative); written wherever SC_RELAX is, so the table kernel's 160 000 threads do
                         // not each evaluate four logarithms
    SC_N
};

// Peer-memory exchange of the statistic table between the ranks of one NVSwitch box (cell-sharded runs).
// Every rank owns an arena [2][table_elems] (double-buffered by sweep parity) that the chain's k_stats fills,
// and a flag array [world] that the PEERS write.  the last CTA of the statistic kernel publishes "sweep `epoch` complete" into the peers'
// flag arrays; k_table waits for all flags and sums the peers' tables while it reads them -- the all-reduce
// is fused into the table kernel, no NCCL call per sweep.
struct P2PState {
    int n_peer = 0, my_rank = 0;
    unsigned long long* arena = nullptr;             // this rank's [3][table_elems]: its own table by sweep parity, then its
                                                     // copy of the SUM over all ranks (carried sum, gibbs.cu::TableParams)
    unsigned long long* my_flags = nullptr;          // [world], written by the peers: [r] = 2 e + c once rank r's table of sweep
                                                     // e is complete; c = 1 if it differs from its table of sweep e - 1
    unsigned long long* peer_arena[MAX_PEERS] = {};  // mapped with cudaIpcOpenMemHandle
    unsigned long long* peer_flags[MAX_PEERS] = {};
    int peer_rank[MAX_PEERS] = {};
    size_t table_elems = 0;
    unsigned long long epoch = 0;                    // sweeps since the arena was mapped (same on every rank)
    unsigned int* error = nullptr;
};

struct SweepArgs {
    int64_t N, M, M_total, cell_offset;
    int K, KP;
    int wise;            // 0 global, 1 variant, 2 element
    int relax;           // Config flips enabled
    int learn_relax_next;// draw relax rate for iteration it+1 at the end of this sweep
    int keep_base;
    double prior0[2], prior1[2], relax_prior[2];
    const double* prior1_mat;   // device [n_el*2] column-major or nullptr
    double relax_fixed;  int relax_fixed_set;
    double logPsi[KMAX];
    const uint32_t* cfg0;       // [N] input Config bit masks
    double W_log;
    PhiloxKey key; uint32_t chain; uint32_t it;   // it = 1-based trace row being filled
    uint32_t rng_it;                              // iteration in the Philox keys (== it except in the bench hook)
    int skip0;                                    // clone 1's Config column is all zero and cannot flip
    int incremental;                              // update SA/SB from the changed cells when they are few
    int64_t T;                                    // trace rows allocated
    // replay (device pointers to row `it-1`), nullptr = own draws
    const double* rp_u_assign; const double* rp_u_config;
    const double* rp_theta0; const double* rp_theta1; const double* rp_relax_next;
    const P2PState* p2p;        // non-null: statistic tables are exchanged over peer memory
    int64_t el_offset;          // wise = element on a cell shard: covered entries of the ranks before this one (Philox index of
                                // an entry = its position in the WHOLE donor's cell-major order, so draws match any GPU count)
};

struct KernelTimes { float cells = 0, stats = 0, table = 0; };

int cells_smem_bytes(int64_t N, int K);
void launch_init_chain(const SweepArgs& a, ChainDev& c, cudaStream_t st);
void launch_cells(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st);
void launch_cells_sell(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st);
void launch_stats(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st);
void launch_table(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st);
// small donors: whole sweeps of all chains in one launch (small.cu)
struct cdl_small_run {
    int n_chain, chain0, min_iter, exact;
    int64_t n_buin_planned;
    uint32_t it_begin, it_end;
    const ChainDev* chains_dev;
    const int* active_dev;
};
bool small_path_fits(const Donor& d, int K, int wise);
void launch_small_chains(const Donor& d, const SweepArgs& a, const cdl_small_run& r, cudaStream_t st);
// wise = element (element.cu)
void launch_el_init(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st);
void launch_el_cells(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st);
void launch_el_stats(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st);
void launch_el_theta(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st);

// post-processing
void launch_geweke(const double* prob_all, int64_t cols, int64_t n_rows, unsigned long long* counts2,
                   cudaStream_t st);  // counts2[0] = #non-NaN, counts2[1] = #|Z|<=2
// the same on n_traces traces of one shape in one launch; counts[2 t], counts[2 t + 1].  Small traces only: returns false
// without launching anything when they are too wide
bool launch_geweke_batch(const double* const* traces_dev, int n_traces, int64_t cols, int64_t n_rows,
                         unsigned long long* counts, cudaStream_t st);
// window snapshots (s1a .. s2b) are fp32 or fp64 arrays, the running sums s1n / s2n always fp64
void launch_geweke_sums(const void* s1a, const void* s2a, const void* s1b, const void* s2b, bool fp32, const double* s1n,
                        const double* s2n, int64_t cols, int64_t n_rows, unsigned long long* counts2, cudaStream_t st);
void launch_window_mean(const double* s1n, const void* s1u, bool fp32, int64_t n, double inv, double* out, cudaStream_t st);
// segment sums of the streaming stop rule: out = (S)acc then acc = 0 (out may be null: only clear); dst += src;
// out = extra + sum of the listed arrays.  Segment arrays are fp32 or fp64 (`fp32`), everything else fp64.
void launch_segment_store(double* acc1, double* acc2, void* out1, void* out2, bool fp32, int64_t n, cudaStream_t st);
// (`squares`: the array holds sums of squares, which fp32 segments keep as their square root)
void launch_segment_add(double* dst, const void* src, bool fp32, bool squares, int64_t n, cudaStream_t st);
void launch_segment_sum(const void* const* ptrs, int count, bool fp32, bool squares, const double* extra, int64_t n,
                        double* out, cudaStream_t st);
// R/clone_id.R:624-644 on the device trace (prob_all [T][cell][clone], Config_all [T][i + k*N]); rows n_buin..it (1-based)
void relabel_trace(double* prob_all, uint8_t* config_all, int64_t M, int64_t N, int K, int64_t n_buin, int64_t it,
                   cudaStream_t st);
void launch_col_means_d(const double* X, int64_t cols, int64_t r0, int64_t r1, double* out, cudaStream_t st);
void launch_transpose_scale(const double* in, int64_t R, int64_t C, double scale, double* out, cudaStream_t st);
void launch_scale(double* x, int64_t n, double s, cudaStream_t st);
void launch_col_means_u8(const uint8_t* X, int64_t cols, int64_t r0, int64_t r1, double* out, cudaStream_t st);
void launch_prob_variant(const Donor& d, int wise, const double* theta0_all, const double* theta1_all,
                         int64_t n_el, int64_t r0, int64_t r1, double* out_dense, cudaStream_t st);   // wise 2 needs d.c_elem0
double loglik_general(const Donor& d, const double* Config_dev, int K, const int32_t* assign_dev,
                      const double* prob_dev /* [M][K] device layout, or null */, double theta0,
                      const double* theta1_dev, int64_t n_theta1, cudaStream_t st);
// EM
void em_stats(const Donor& d, const uint32_t* cfg_mask, int K, uint32_t* S3, uint32_t* S4, uint32_t* TA,
              uint32_t* TB, int sm_count, cudaStream_t st);
void em_iteration(int64_t M, int K, const uint32_t* S3, const uint32_t* S4, const uint32_t* TA,
                  const uint32_t* TB, const double* logPsi_dev, double theta0, double theta1,
                  double* prob_out /* [M][K] */, double* sums5_dev, double* partials, cudaStream_t st);

}  // namespace cdl
