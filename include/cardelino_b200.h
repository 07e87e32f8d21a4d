/*
 * cardelino_b200 -- C ABI of the B200-native clone-assignment sampler.
 *
 * The reference (single-cell-genetics/cardelino, pure R) has NO native interface for this path:
 * there is no src/, no .Call, no useDynLib.  The boundary a maintainer would bind is
 * therefore new; every entry point below names the reference R code whose body it replaces.
 * R binds these through r/cardelino_b200_shim.c (.Call); Python binds them with ctypes
 * (cardelino_b200/_lib.py).  See INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices crossing the ABI are COLUMN-MAJOR doubles exactly as
 *     R stores them (variants x cells for A/D, variants x clones for Config, cells x clones for prob).
 *   - caller-owned inputs are never modified; outputs are caller-allocated.
 *   - every call returns 0 on success or a CDL_ERR_* code; cdl_last_error() gives the message.
 *     No exceptions, no longjmp cross the ABI.
 *   - all device state lives behind the opaque handle.  There is no CPU fallback: cdl_create fails
 *     with CDL_ERR_CUDA when no sm_100-class device is usable.
 *   - iteration numbering follows the reference: trace row 1 holds the initial theta draws, sweeps are
 *     it = 2..max_iter (R/clone_id.R:461-466).
 */
#ifndef CARDELINO_B200_H
#define CARDELINO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdl_handle_s* cdl_handle;

enum {
    CDL_OK = 0,
    CDL_ERR_INVALID = 1,     /* bad argument / shape / value (maps to the reference's stop() checks) */
    CDL_ERR_CUDA = 2,        /* CUDA runtime failure or no usable device */
    CDL_ERR_NOMEM = 3,
    CDL_ERR_STATE = 4,       /* call order violated (e.g. fetch before run) */
    CDL_ERR_UNSUPPORTED = 5, /* documented limit (K > 32, counts > 65535, ...) */
    CDL_ERR_COMM = 6,        /* NCCL failure */
    CDL_ERR_INTERRUPTED = 7  /* the progress callback asked to stop (R: user interrupt); no results are kept */
};

/* Progress callback of cdl_gibbs_run.  Called on the calling thread (never from a worker thread) for every chain
 * every check_every sweeps; `fraction` is the converged fraction of the reference's Geweke check when one was made
 * at this iteration (R/clone_id.R:580-592; the host prints "x% converged." from it like the reference's cat()),
 * NaN otherwise.  A non-zero return value stops the run with CDL_ERR_INTERRUPTED.  The library prints nothing itself
 * (two debugging switches write to stderr: CDL_TIMELINE=1, a device timeline from cdl_bench_sweeps, and CDL_PHASES=1, the
 * host seconds of the phases of a small-donor run). */
typedef int (*cdl_progress_fn)(void* user, int32_t chain, int32_t iteration, double fraction);

enum { CDL_WISE_GLOBAL = 0, CDL_WISE_VARIANT = 1, CDL_WISE_ELEMENT = 2 };

/* Parameters of clone_id_Gibbs (R/clone_id.R:351-356).  Plain data; fill with cdl_gibbs_defaults(). */
typedef struct cdl_gibbs_params {
    int32_t relax_Config;        /* R/clone_id.R:352; 0 => Config_new == Config (the reference errors) */
    int32_t relax_rate_fixed_set;/* 0 => relax_rate_fixed = NULL (learned, :442-444) */
    double  relax_rate_fixed;    /* in [0,1] (:437-439) */
    double  relax_rate_prior[2]; /* (:353) */
    int32_t keep_base_clone;     /* (:353,453-455) */
    double  prior0[2];           /* (:354) */
    double  prior1[2];           /* (:354) used for every element unless prior1_matrix is given */
    const double* prior1_matrix; /* optional n_element x 2 column-major (:418-424), may be NULL */
    int32_t min_iter, max_iter;  /* (:355) */
    double  buin_frac;           /* (:355) */
    int32_t wise;                /* CDL_WISE_* (:356,406-416) */
    int32_t n_chain;             /* chains run inside the library (replaces doMC fork, :127-163) */
    uint64_t seed;               /* Philox key; chain c uses (seed, c) */
    int32_t keep_assign_all;     /* also record assign_all (R allocates it at :429 but does not return it) */
    int32_t keep_prob_all;       /* 1: exact mode, keep prob_all on device (needed for relabel and for fetching
                                    the trace); 0: streaming mode: running column sums of prob and prob^2 plus
                                    snapshots at the rows where a Geweke window or the burn-in of a later check
                                    starts, which give the reference's stop rule (:580-601) and final-iteration
                                    burn-in (:604,645) without the trace; -1: decide from trace_budget_bytes */
    int64_t trace_budget_bytes;  /* upper bound for on-device traces when keep_prob_all == -1 */
    int32_t check_every;         /* convergence check period, reference value 100 (:580) */
    int32_t verbose;
    int32_t acc_fp32;            /* reserved, must be 0 (fp64 accumulation) */
    int32_t chain_offset;        /* chains are numbered chain_offset .. chain_offset + n_chain - 1 in the Philox keys, so
                                    chain groups placed on different GPUs / processes draw independent streams */
    int32_t incremental_stats;   /* when at most ~3 % of the cells changed label in a sweep, correct the previous
                                    sweep's SA/SB table from those cells instead of re-reading every count (exact:
                                    integer sums, bit-identical chains; falls back to the full pass sweep by sweep).
                                    -1 (default): on wherever it applies (wise = global / variant; one GPU or the
                                    peer-memory exchange; not replay, not the single-CTA small-donor path); 1: on,
                                    which also takes small donors off the single-CTA path; 0: always the full pass */
    int32_t relabel;             /* R/clone_id.R:624-644: match the clone columns of consecutive prob_all rows from
                                    n_buin on (colMatch; all permutations for K <= 5, greedy above) and permute
                                    prob_all / Config_all rows accordingly, on the device, before the summaries.
                                    Needs the exact trace (keep_prob_all = 1, or -1 with a trace that fits). */
    int64_t prior1_matrix_rows;  /* rows of prior1_matrix (checked against n_element; 0 = not given) */
    cdl_progress_fn progress;    /* may be NULL */
    void* progress_user;
    int32_t snapshot_fp32;       /* streaming stop rule: keep the window snapshots as fp32 (default 1; half the memory,
                                    Z changes by ~1e-6 relative) or fp64 (0) */
} cdl_gibbs_params;

/* Optional injected draws ("replay mode", north_star correctness check 2).  Any pointer may be NULL,
 * in which case the library's own Philox / Beta sampler supplies that draw.  Row t (0-based) of each
 * array feeds iteration it = t + 1 of the reference loop (row 0 = initial draws, :461-462).
 * Only n_chain == 1 is accepted in replay mode. */
typedef struct cdl_replay {
    int32_t n_rows;              /* rows available (>= max_iter) */
    const double* u_assign;      /* n_rows x M   row-major: uniform consumed by sample() for cell j (:493) */
    const double* u_config;      /* n_rows x N*K row-major: uniform for rbinom element i + k*N (:534) */
    const double* theta0;        /* n_rows: theta0_all[t] (:461,563) */
    const double* theta1;        /* n_rows x n_element row-major: theta1_all[t,] (:462,567) */
    const double* relax_rate;    /* n_rows: value of the rbeta at :504 for iteration t+1 (ignored rows allowed) */
} cdl_replay;

/* Result sizes of the last cdl_gibbs_run. */
typedef struct cdl_gibbs_info {
    int32_t n_iter;              /* final `it` of chain (R/clone_id.R:593-594) */
    int32_t n_buin;              /* ceiling(it * buin_frac) (:604) */
    int32_t n_element;           /* 1, N or nnz (:406-416) */
    int32_t exact_trace;         /* 1 if prob_all was kept */
    double  percent_converged;   /* (:595); streaming mode: the same test from running sums when a stop before
                                    max_iter is possible (min_iter < max_iter, snapshots within the trace budget),
                                    else the Geweke test on the theta1 traces */
    double  logLik_post;         /* (:656) */
    double  W_log;               /* (:388) */
    double  dic[6];              /* DIC, logLik_var, D_mean, D_post, DIC_Gelman, DIC_Spiegelhalter (:790-797) */
    int32_t stop_rule;           /* how the reference's stop rule (:580-592) was applied: 0 = on the stored prob_all;
                                    1 = the same test from running sums and window snapshots (streaming mode);
                                    2 = not at all: max_iter sweeps (min_iter == max_iter, snapshots over the trace
                                    budget, the single-CTA path without a trace, wise = "element" without a trace) */
} cdl_gibbs_info;

int  cdl_version(void);
/* message of the last failing call on this handle (or of the last failing cdl_create when h == NULL) */
const char* cdl_last_error(cdl_handle h);

int cdl_create(int device, cdl_handle* out);
int cdl_destroy(cdl_handle h);

/* ---- data: replaces as.matrix() + the NA normalisation + S-lists (R/clone_id.R:121-122,380-403) ----
 * Dense column-major N x M doubles; NA (NaN) or 0 in D means "not covered"; A is NA/ignored there.
 * Errors mirror the reference checks (non-integer values, R/clone_id.R:93-96). */
int cdl_load_dense(cdl_handle h, int64_t N, int64_t M, const double* A, const double* D);
/* Two dgCMatrix-style CSC matrices (cells are columns), as load_cellSNP_vcf returns them
 * (R/read_data.R:293-294): p has M+1 entries, i are 0-based variant rows sorted within a column. */
int cdl_load_csc(cdl_handle h, int64_t N, int64_t M,
                 const int64_t* d_p, const int32_t* d_i, const double* d_x,
                 const int64_t* a_p, const int32_t* a_i, const double* a_x);
/* Already-merged compact form: one pattern (D > 0 everywhere), uint16 counts. */
int cdl_load_csc_u16(cdl_handle h, int64_t N, int64_t M, const int64_t* p, const int32_t* i,
                     const uint16_t* a, const uint16_t* d);
/* Synthetic donor generated on the device with sim_read_count's distributions
 * (R/reads_simulator.R:53-152,174-212): coverage ~ Bernoulli(cov), depths from an empirical table,
 * theta0 ~ Beta(0.2, 99.8) per entry, theta1 ~ Beta(0.45, 0.55) per variant, labels ~ Multinomial(1/K).
 * cell_offset/M_local select this rank's cell shard of an M_total-cell donor (same data at any rank count).
 * Config (N x K column-major, binary) is the generating tree; true labels can be fetched afterwards. */
int cdl_synth_generate(cdl_handle h, int64_t N, int64_t M_total, int64_t cell_offset, int64_t M_local,
                       int32_t K, const double* Config, double coverage, uint64_t seed);
int cdl_synth_labels(cdl_handle h, int32_t* labels_out /* M_local, 1-based */);
/* Export the resident counts in the cdl_load_csc_u16 layout (sizes from cdl_data_info). */
int cdl_export_csc_u16(cdl_handle h, int64_t* p, int32_t* i, uint16_t* a, uint16_t* d);
int cdl_data_info(cdl_handle h, int64_t* N, int64_t* M, int64_t* nnz, double* W_log);
/* Declare the loaded cells to be cells [cell_offset, cell_offset + M) of an M_total-cell donor (cell-sharded
 * multi-GPU runs): the Philox stream of a cell is keyed by its GLOBAL id, so draws do not depend on the
 * number of ranks.  Call after cdl_load_*. */
int cdl_set_shard(cdl_handle h, int64_t cell_offset, int64_t M_total);

/* Which cell kernel the Gibbs sweep uses: ROWS = several lanes per cell on the row layout (small donors),
 * SLICES = one lane per cell on the slice-major layout (large donors), AUTO = SLICES when the whole donor
 * has >= 16384 cells.  SMALL = a small donor's whole sweep in one CTA, many sweeps per launch, one CTA per chain
 * (single GPU, no replay; AUTO takes it whenever the donor fits: <= 65536 cells, <= 2^18 covered entries, N*K <=
 * 16384).  Results agree to rounding (the per-cell summation order differs). */
enum { CDL_CELLS_AUTO = 0, CDL_CELLS_ROWS = 1, CDL_CELLS_SLICES = 2, CDL_CELLS_SMALL = 3 };
int cdl_set_cells_layout(cdl_handle h, int32_t layout);

/* ---- multi-GPU (cells sharded over ranks; one all-reduce of the N x K statistic table per sweep) ----
 * nccl_lib_path: libnccl.so.2 to dlopen (the host passes torch's bundled copy); unique_id: the 128-byte
 * ncclUniqueId broadcast by the host plumbing.  cdl_comm_unique_id fills it on rank 0. */
int cdl_comm_unique_id(cdl_handle h, const char* nccl_lib_path, uint8_t* id128);
int cdl_comm_init(cdl_handle h, const char* nccl_lib_path, const uint8_t* id128, int32_t rank, int32_t world);
/* How the per-sweep N x K statistic table is summed over the ranks: PEER = every rank's table lives in a
 * cudaIpc-shared arena and the table kernel sums the peers' tables over NVLink while it reads them (flags in
 * peer memory order the sweeps; no collective call per sweep); NCCL = one ncclAllReduce per sweep; AUTO = PEER
 * when every rank can map its peers (one NVSwitch box, <= 8 ranks, one rank per GPU), else NCCL.  Must be set
 * identically on every rank before cdl_gibbs_run.  Integer sums: both give bit-identical chains. */
enum { CDL_EXCHANGE_AUTO = 0, CDL_EXCHANGE_NCCL = 1, CDL_EXCHANGE_PEER = 2 };
int cdl_set_exchange_mode(cdl_handle h, int32_t mode);
int cdl_exchange_is_peer(cdl_handle h);   /* 1 if the last cdl_gibbs_run used the peer-memory exchange */

/* ---- clone_id_Gibbs (R/clone_id.R:351-675) ---- */
void cdl_gibbs_defaults(cdl_gibbs_params* p);
/* Config: N x K column-major (binary); Psi: K or NULL (uniform, :357-359). */
int cdl_gibbs_run(cdl_handle h, const double* Config, int32_t K, const double* Psi,
                  const cdl_gibbs_params* params, const cdl_replay* replay /* may be NULL */);
int cdl_gibbs_info_get(cdl_handle h, int32_t chain, cdl_gibbs_info* out);
/* incremental_stats = 1: how many sweeps of the chain (bench sweeps included) took the full statistic pass */
int cdl_incremental_full_passes(cdl_handle h, int32_t chain, uint32_t* out);
/* Fetch fields of the reference return list (:662-673); `chain` selects the chain.  All column-major
 * doubles with R's shapes, EXCEPT the three big traces (prob_all, Config_all, assign_all), which are
 * returned ROW-major (n_iter rows, R's column order within a row: cell j + clone k * M, variant
 * i + clone k * N) so that one iteration is contiguous; the R shim transposes.  Trace fetches return the
 * first n_iter rows. */
int cdl_fetch_prob(cdl_handle h, int32_t chain, double* out /* M x K */);
int cdl_fetch_config_prob(cdl_handle h, int32_t chain, double* out /* N x K */);
int cdl_fetch_theta(cdl_handle h, int32_t chain, double* theta0, double* theta1_elems /* n_element means */);
int cdl_fetch_theta_traces(cdl_handle h, int32_t chain, double* theta0_all /* n_iter */,
                           double* theta1_all /* n_iter x n_element */);
int cdl_fetch_loglik_trace(cdl_handle h, int32_t chain, double* out /* n_iter */);
int cdl_fetch_relax_rate(cdl_handle h, int32_t chain, double* mean_out, double* all_out /* n_iter or NULL */);
int cdl_fetch_prob_all(cdl_handle h, int32_t chain, double* out /* n_iter x M*K */);
int cdl_fetch_config_all(cdl_handle h, int32_t chain, double* out /* n_iter x N*K */);
int cdl_fetch_assign_all(cdl_handle h, int32_t chain, int32_t* out /* n_iter x M, 1-based, row 1 zeros */);
int cdl_fetch_prob_variant(cdl_handle h, int32_t chain, double* out /* N x M dense, NaN-free (:606-622) */);
/* indices (0-based, column-major i + j*N) of the covered entries, in theta1_all column order for wise=element */
int cdl_fetch_element_index(cdl_handle h, int64_t* out /* nnz */);

/* ---- test hooks (tests/ only) ----
 * cdl_debug_capture(h, 1): every following sweep keeps a copy of this rank's SA/SB statistic table as it stands
 * between the statistic kernel and the table kernel (before any exchange with other ranks), and the table kernel
 * records the integers it adds to the Beta priors: S3[i], S4[i] per variant (R/clone_id.R:557-562) and the totals
 * S1, S2, S3, S4, diff1 (:546-556,502).  cdl_debug_fetch returns those of the last sweep of chain 0. */
int cdl_debug_capture(cdl_handle h, int32_t enable);
int cdl_debug_fetch(cdl_handle h, uint64_t* SA /* N x K row-major [i*K+k] */, uint64_t* SB, uint64_t* S3 /* N */,
                    uint64_t* S4 /* N */, uint64_t* totals5);

/* ---- get_logLik (R/clone_id.R:734-752) ----
 * Config N x K doubles (may be fractional); exactly one of assign (1-based labels, M) / assign_prob
 * (M x K) is non-NULL; theta1 has n_theta1 in {1, N, nnz} values (nnz: one per covered entry in the order of
 * cdl_fetch_element_index, wise = "element"). */
int cdl_get_loglik(cdl_handle h, const double* Config, int32_t K, const int32_t* assign,
                   const double* assign_prob, double theta0, const double* theta1, int64_t n_theta1,
                   double* out);

/* ---- clone_id_EM (R/clone_id.R:240-326) ----
 * theta_init: the two runif() starts of :278 (NULL => drawn from Philox(seed)). */
int cdl_em_run(cdl_handle h, const double* Config, int32_t K, const double* Psi, int32_t min_iter,
               int32_t max_iter, double logLik_threshold, const double* theta_init, uint64_t seed,
               double* theta_out /* 2 */, double* prob_out /* M x K */, double* logLik_out,
               int32_t* n_iter_out);

/* ---- measurement hooks (bench.py): run `n` sweeps of chain state on the resident data and report
 * device times measured with CUDA events on the library's stream. ---- */
int cdl_bench_sweeps(cdl_handle h, int32_t warmup, int32_t n, float* ms_total, float* ms_cells,
                     float* ms_stats, float* ms_table, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
