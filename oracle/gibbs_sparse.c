/*
 * TEST INFRASTRUCTURE ONLY -- sparse C restatement (fp64, OpenMP) of cardelino's Gibbs sweep.
 *
 * Checker and CPU baseline for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; nothing under cardelino_b200/ does.
 * PARITY UNPINNED, like oracle/cardelino_oracle.py: the reference is pure R and R is not installed; this file is
 * pinned against the NumPy restatement (tests/test_oracle.py::test_c_oracle_*), which follows the R source op
 * for op.  It uses the sufficient-statistic form of SURVEY.md section 7 so that it can run the BASELINE C3/C4
 * shapes the dense R formulation cannot hold:
 *
 *   orc_sweep_cells   logLik -> softmax -> sample()            R/clone_id.R:470-497
 *                     + SA[i,k] = sum_{j: z_j = k} A_ij, SB    (replaces the S-lists, :394-403,537-562)
 *   orc_sweep_table   Config flips, theta statistics, logLik   R/clone_id.R:519-535,546-562,574-577 (-> :734-752)
 *   orc_run           T sweeps with its own RNG (timing leg)   R/clone_id.R:466-593
 *   orc_synth         sim_read_count-shaped donor              R/reads_simulator.R:53-152,174-212
 *
 * Counts come as one CSC pattern over cells (D > 0 everywhere): p[M+1], i[nnz] (variant, ascending inside a
 * cell), a[nnz], d[nnz] as uint16.  Config rows are clone bit masks.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KMAX 32

int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* number of OpenMP threads of the following calls.  bench.py sets it to the host's core count explicitly: under
 * torch.distributed.run the launcher exports OMP_NUM_THREADS=1, which would time the port on one core. */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* R's revsort (src/main/sort.c): heapsort a[] into descending order, permuting ib[] alongside */
static void revsort(double* a, int* ib, int n) {
    int l, j, ir, i, ii;
    double ra;
    if (n <= 1) return;
    a--; ib--;
    l = (n >> 1) + 1;
    ir = n;
    for (;;) {
        if (l > 1) { l = l - 1; ra = a[l]; ii = ib[l]; }
        else {
            ra = a[ir]; ii = ib[ir];
            a[ir] = a[1]; ib[ir] = ib[1];
            if (--ir == 1) { a[1] = ra; ib[1] = ii; return; }
        }
        i = l; j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) { a[i] = a[j]; ib[i] = ib[j]; j += (i = j); }
            else j = ir + 1;
        }
        a[i] = ra; ib[i] = ii;
    }
}

/* sample(seq_len(K), 1, replace = TRUE, prob = pr) with its uniform injected (base R ProbSampleReplace) */
static int r_sample_one(const double* pr, int K, double u) {
    double p[KMAX], s = 0.0;
    int perm[KMAX], k, j;
    for (k = 0; k < K; ++k) s += pr[k];
    for (k = 0; k < K; ++k) { p[k] = pr[k] / s; perm[k] = k; }
    revsort(p, perm, K);
    for (k = 1; k < K; ++k) p[k] += p[k - 1];
    for (j = 0; j < K - 1; ++j) if (u <= p[j]) break;
    return perm[j];
}

/* one cell pass.  theta1 has n_t1 in {1, N} values.  prob (M x K row-major) may be NULL. */
void orc_sweep_cells(int64_t N, int64_t M, int K, const int64_t* p, const int32_t* vi, const uint16_t* a,
                     const uint16_t* d, const uint32_t* mask, double theta0, const double* theta1, int64_t n_t1,
                     const double* logPsi, const double* u_assign, double* prob, uint8_t* z, uint64_t* SA,
                     uint64_t* SB) {
    const double l0 = log(theta0), l0c = log(1.0 - theta0);
    double* du = (double*)malloc(sizeof(double) * (size_t)N);
    double* dv = (double*)malloc(sizeof(double) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) {
        const double t = theta1[n_t1 == 1 ? 0 : i];
        du[i] = log(t) - l0;
        dv[i] = log(1.0 - t) - l0c;
    }
    memset(SA, 0, sizeof(uint64_t) * (size_t)(N * K));
    memset(SB, 0, sizeof(uint64_t) * (size_t)(N * K));
#pragma omp parallel
    {
        uint64_t* sa = (uint64_t*)calloc((size_t)(N * K), sizeof(uint64_t));
        uint64_t* sb = (uint64_t*)calloc((size_t)(N * K), sizeof(uint64_t));
#pragma omp for schedule(dynamic, 256)
        for (int64_t j = 0; j < M; ++j) {
            double L[KMAX], pr[KMAX];
            for (int k = 0; k < K; ++k) L[k] = 0.0;
            for (int64_t e = p[j]; e < p[j + 1]; ++e) {
                const int64_t i = vi[e];
                const double del = (double)a[e] * du[i] + (double)(d[e] - a[e]) * dv[i];
                uint32_t m = mask[i];
                while (m) { const int k = __builtin_ctz(m); L[k] += del; m &= m - 1; }
            }
            double mx = -INFINITY, sm = 0.0;
            for (int k = 0; k < K; ++k) { L[k] += logPsi[k]; if (L[k] > mx) mx = L[k]; }
            for (int k = 0; k < K; ++k) { pr[k] = exp(L[k] - mx); sm += pr[k]; }
            for (int k = 0; k < K; ++k) pr[k] /= sm;
            if (prob) for (int k = 0; k < K; ++k) prob[j * K + k] = pr[k];
            const int zj = r_sample_one(pr, K, u_assign[j]);
            z[j] = (uint8_t)zj;
            for (int64_t e = p[j]; e < p[j + 1]; ++e) {
                sa[(int64_t)vi[e] * K + zj] += a[e];
                sb[(int64_t)vi[e] * K + zj] += (uint64_t)(d[e] - a[e]);
            }
        }
#pragma omp critical
        for (int64_t t = 0; t < N * K; ++t) { SA[t] += sa[t]; SB[t] += sb[t]; }
        free(sa); free(sb);
    }
    free(du); free(dv);
}

/* rbinom(1, 1, p) with its uniform injected (base R nmath/rbinom.c, size-1 inversion) */
static int rbinom1(double p, double u) {
    if (p == 0.0) return 0;
    if (p == 1.0) return 1;
    const double pm = p < 1.0 - p ? p : 1.0 - p, q = 1.0 - pm;
    int ix = (u < q) ? 0 : 1;
    if (p > 0.5) ix = 1 - ix;
    return ix;
}

/* N x K step with the OLD theta: Config flips (u_config indexed i + k*N like R's column-major rbinom),
 * sums for the theta draws, logLik trace.  stats_out: S1, S2 (scalars), then S3[N], S4[N]. */
void orc_sweep_table(int64_t N, int K, const uint64_t* SA, const uint64_t* SB, const uint32_t* cfg0, int relax,
                     int keep_base, double relax_rate, double theta0, const double* theta1, int64_t n_t1,
                     const double* u_config, double W_log, uint32_t* mask_new, double* stats_out,
                     double* loglik_out, int64_t* diff1_out) {
    const double l0 = log(theta0), l0c = log(1.0 - theta0);
    const double odd1 = log(1.0 - relax_rate) - log(relax_rate), odd0 = log(relax_rate) - log(1.0 - relax_rate);
    double S1 = 0, S2 = 0, ll = 0;
    int64_t diff1 = 0;
    for (int64_t i = 0; i < N; ++i) {
        const double t = theta1[n_t1 == 1 ? 0 : i];
        const double l1 = log(t), l1c = log(1.0 - t);
        double s3 = 0, s4 = 0;
        uint32_t nm = 0;
        for (int k = 0; k < K; ++k) {
            const double sa = (double)SA[i * K + k], sb = (double)SB[i * K + k];
            const int cbit = (cfg0[i] >> k) & 1;
            int bit = cbit;
            if (relax) {
                double prior = (keep_base && k == 0) ? (cbit ? INFINITY : -INFINITY) : (cbit ? odd1 : odd0);
                double x = sa * (l1 - l0) + sb * (l1c - l0c) + prior;
                x = x > 50.0 ? 50.0 : (x < -50.0 ? -50.0 : x);
                const double ex = exp(x);
                bit = rbinom1(ex / (ex + 1.0), u_config[i + (int64_t)k * N]);
            }
            nm |= (uint32_t)bit << k;
            if (bit) { s3 += sa; s4 += sb; ll += sa * l1 + sb * l1c; }
            else { S1 += sa; S2 += sb; ll += sa * l0 + sb * l0c; }
            if (k >= 1) diff1 += (bit != cbit);
        }
        mask_new[i] = nm;
        stats_out[2 + i] = s3;
        stats_out[2 + N + i] = s4;
    }
    stats_out[0] = S1; stats_out[1] = S2;
    *loglik_out = ll + W_log;
    *diff1_out = diff1;
}

/* ---- self-contained runner for timing: splitmix64 uniforms, Marsaglia-Tsang gammas ---- */
static inline uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double unif(uint64_t* s) { return ((double)(splitmix(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double gamma_mt(double a, uint64_t* s) {
    double boost = 1.0;
    if (a < 1.0) { boost = pow(unif(s), 1.0 / a); a += 1.0; }
    const double dd = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
    for (;;) {
        const double x = sqrt(-2.0 * log(unif(s))) * cos(6.283185307179586 * unif(s));
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = unif(s);
        if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + dd * (1.0 - v + log(v))) return boost * dd * v;
    }
}
static double beta_draw(double a, double b, uint64_t* s) {
    const double x = gamma_mt(a, s), y = gamma_mt(b, s);
    double t = x / (x + y);
    if (!(t > 1e-300)) t = 1e-300;
    if (t > 1.0 - 1.2e-16) t = 1.0 - 1.2e-16;
    return t;
}

/* T sweeps (wise = variant, relax_Config = TRUE with a learned rate, defaults of R/clone_id.R:351-356); returns the
 * seconds spent in the sweeps.  z_out (M) and theta1_out (N) receive the final state. */
double orc_run(int64_t N, int64_t M, int K, const int64_t* p, const int32_t* vi, const uint16_t* a, const uint16_t* d,
               const uint32_t* cfg0, int T, uint64_t seed, uint8_t* z_out, double* theta1_out) {
    uint64_t s = seed * 0x2545F4914F6CDD1Dull + 1;
    double* theta1 = (double*)malloc(sizeof(double) * (size_t)N);
    double* stats = (double*)malloc(sizeof(double) * (size_t)(2 + 2 * N));
    double* ua = (double*)malloc(sizeof(double) * (size_t)M);
    double* uc = (double*)malloc(sizeof(double) * (size_t)(N * K));
    uint64_t* SA = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N * K));
    uint64_t* SB = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N * K));
    uint32_t* mask = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
    uint32_t* mask2 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
    double logPsi[KMAX];
    for (int k = 0; k < K; ++k) logPsi[k] = -log((double)K);
    double theta0 = beta_draw(0.2, 99.8, &s), relax = 0.1, ll;
    int64_t diff1 = 0;
    for (int64_t i = 0; i < N; ++i) { theta1[i] = beta_draw(0.45, 0.55, &s); mask[i] = cfg0[i]; }
    double t0 = 0.0, t1 = 0.0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#endif
    for (int it = 2; it <= T + 1; ++it) {
        for (int64_t j = 0; j < M; ++j) ua[j] = unif(&s);
        for (int64_t q = 0; q < N * K; ++q) uc[q] = unif(&s);
        orc_sweep_cells(N, M, K, p, vi, a, d, mask, theta0, theta1, N, logPsi, ua, NULL, z_out, SA, SB);
        if (it > 2) relax = beta_draw(1.0 + (double)diff1, 9.0 + (double)(N * (K - 1) - diff1), &s);
        orc_sweep_table(N, K, SA, SB, cfg0, 1, 1, relax, theta0, theta1, N, uc, 0.0, mask2, stats, &ll, &diff1);
        memcpy(mask, mask2, sizeof(uint32_t) * (size_t)N);
        theta0 = beta_draw(0.2 + stats[0], 99.8 + stats[1], &s);
        for (int64_t i = 0; i < N; ++i) theta1[i] = beta_draw(0.45 + stats[2 + i], 0.55 + stats[2 + N + i], &s);
    }
#ifdef _OPENMP
    t1 = omp_get_wtime();
#endif
    if (theta1_out) memcpy(theta1_out, theta1, sizeof(double) * (size_t)N);
    free(theta1); free(stats); free(ua); free(uc); free(SA); free(SB); free(mask); free(mask2);
    return t1 - t0;
}

/* sim_read_count-shaped donor (R/reads_simulator.R:53-152,174-212): labels ~ Multinomial(1/K), coverage ~
 * Bernoulli(cov) per entry, depth from `depth_q[256]` (quantiles of the packaged D_input), theta1 ~ Beta(0.45,
 * 0.55) per variant, theta0 ~ Beta(0.2, 99.8) per entry.  Call once with vi == NULL to get the column pointers p
 * (returns nnz), then again with the arrays allocated.  Same counts both times (per-cell seeds). */
int64_t orc_synth(int64_t N, int64_t M, int K, double cov, const uint32_t* cfg, const uint8_t* depth_q, uint64_t seed,
                  int64_t* p, int32_t* vi, uint16_t* a, uint16_t* d, int32_t* labels) {
    double* th1 = (double*)malloc(sizeof(double) * (size_t)N);
    uint64_t s0 = seed ^ 0xabcdef1234567ull;
    for (int64_t i = 0; i < N; ++i) th1[i] = beta_draw(0.45, 0.55, &s0);
    const double l1m = log(1.0 - cov);
    if (!vi) {
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < M; ++j) {
            uint64_t s = (seed + 1) * 0x9e3779b97f4a7c15ull + (uint64_t)j * 0xD1B54A32D192ED03ull;
            (void)unif(&s);
            int64_t i = -1, n = 0;
            for (;;) {
                i += 1 + (int64_t)floor(log(unif(&s)) / l1m);
                if (i >= N) break;
                (void)unif(&s); (void)unif(&s); (void)unif(&s); (void)splitmix(&s);
                ++n;
            }
            p[j + 1] = n;
        }
        p[0] = 0;
        for (int64_t j = 0; j < M; ++j) p[j + 1] += p[j];
        free(th1);
        return p[M];
    }
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < M; ++j) {
        uint64_t s = (seed + 1) * 0x9e3779b97f4a7c15ull + (uint64_t)j * 0xD1B54A32D192ED03ull;
        int lab = (int)(unif(&s) * K);
        if (lab >= K) lab = K - 1;
        if (labels) labels[j] = lab + 1;
        int64_t i = -1, o = p[j];
        for (;;) {
            i += 1 + (int64_t)floor(log(unif(&s)) / l1m);
            if (i >= N) break;
            const uint32_t dep = depth_q[(uint32_t)(unif(&s) * 256.0) & 255u];
            const double ub = unif(&s), ut = unif(&s);
            double th = th1[i];
            uint64_t st = splitmix(&s) ^ (uint64_t)(ut * 1e18);          /* consumed for every entry: both passes agree */
            if (!((cfg[i] >> lab) & 1u)) th = beta_draw(0.2, 99.8, &st);
            /* binomial by inversion */
            double q = 1.0 - th, f = pow(q, (double)dep), u = ub;
            uint32_t ix = 0;
            while (ix < dep && u >= f) { u -= f; f *= (th / q) * (double)(dep - ix) / (double)(ix + 1); ++ix; }
            vi[o] = (int32_t)i; a[o] = (uint16_t)ix; d[o] = (uint16_t)dep;
            ++o;
        }
    }
    free(th1);
    return p[M];
}
