#!/usr/bin/env Rscript
# Records a GENUINE reference run for replay-mode parity (north_star check 2).  Needs R + cardelino; not
# runnable in the build image.  It wraps the three RNG consumers of clone_id_Gibbs -- sample() (:493),
# rbinom() (:534) and rbeta() (:461-462,504,563,567) -- so that the uniforms and Beta draws are written out,
# then runs the stock function.  sample(K, 1, prob=) and rbinom(n, 1, p) each consume exactly one
# unif_rand() per call / per element with 0 < p < 1, so the recorded uniforms are re-drawn here with the
# same seed and handed to BOTH implementations.
#
#   Rscript tools/record_reference.R out.npz-dir
#
# Writes CSVs: u_assign (T x M), u_config (T x N*K), theta0, theta1 (T x n_element), relax_rate, assign_all,
# Config_all, prob_all, logLik.  Load them in Python and pass as the `replay` dict of
# cardelino_b200.clone_id_Gibbs; tests/test_gpu_parity.py::test_replay_mode shows the comparison.
suppressMessages(library(cardelino))
args <- commandArgs(trailingOnly = TRUE)
outdir <- if (length(args)) args[1] else "reference_recording"
dir.create(outdir, showWarnings = FALSE)
data(example_donor)
T <- 200
rec <- new.env(); rec$ua <- list(); rec$uc <- list(); rec$beta <- list()
ns <- asNamespace("cardelino")
gibbs <- get("clone_id_Gibbs", ns)
body_txt <- deparse(body(gibbs))
body_txt <- gsub("sample(seq_len(K), 1,", ".rec_sample(seq_len(K), 1,", body_txt, fixed = TRUE)
body_txt <- gsub("stats::rbinom(", ".rec_rbinom(", body_txt, fixed = TRUE)
body_txt <- gsub("stats::rbeta(", ".rec_rbeta(", body_txt, fixed = TRUE)
body(gibbs) <- parse(text = paste(body_txt, collapse = "\n"))[[1]]
environment(gibbs) <- list2env(list(
    .rec_sample = function(x, size, replace, prob) {
        u <- stats::runif(1); rec$ua[[length(rec$ua) + 1]] <- u
        p <- prob / sum(prob); o <- order(-p); cum <- cumsum(p[o])   # NB: revsort tie order differs from order()
        x[o[min(which(u <= cum), length(x))]]
    },
    .rec_rbinom = function(n, size, prob) {
        u <- stats::runif(n); rec$uc[[length(rec$uc) + 1]] <- u
        as.numeric(ifelse(prob > 0.5, u < prob, u >= 1 - prob))
    },
    .rec_rbeta = function(n, a, b) { v <- stats::rbeta(n, a, b); rec$beta[[length(rec$beta) + 1]] <- v; v }),
    parent = ns)
set.seed(7)
res <- gibbs(A_clone, D_clone, tree$Z, min_iter = T, max_iter = T, relax_Config = TRUE)
write.csv(matrix(unlist(rec$ua), ncol = ncol(A_clone), byrow = TRUE), file.path(outdir, "u_assign.csv"), row.names = FALSE)
write.csv(do.call(rbind, rec$uc), file.path(outdir, "u_config.csv"), row.names = FALSE)
write.csv(res$theta0_all, file.path(outdir, "theta0.csv"), row.names = FALSE)
write.csv(res$theta1_all, file.path(outdir, "theta1.csv"), row.names = FALSE)
write.csv(res$relax_rate_all, file.path(outdir, "relax_rate.csv"), row.names = FALSE)
write.csv(res$prob_all, file.path(outdir, "prob_all.csv"), row.names = FALSE)
write.csv(res$Config_all, file.path(outdir, "Config_all.csv"), row.names = FALSE)
write.csv(res$logLik, file.path(outdir, "logLik.csv"), row.names = FALSE)
