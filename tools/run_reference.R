#!/usr/bin/env Rscript
# Times the genuine reference (R cardelino) for the BASELINE metric: Gibbs iterations/sec with
# n_chain = n_proc = all cores.  Not runnable in the build image (no R); for users who have R.
#   Rscript tools/run_reference.R <cells> <variants> <clones> <iters>
suppressMessages(library(cardelino))
args <- as.numeric(commandArgs(trailingOnly = TRUE))
M <- if (length(args) > 0) args[1] else 2000; N <- if (length(args) > 1) args[2] else 2000
K <- if (length(args) > 2) args[3] else 8;    T <- if (length(args) > 3) args[4] else 20
data(simulation_input)
set.seed(20261019)
Config <- matrix(rbinom(N * K, 1, 0.3), N, K); Config[, 1] <- 0
rownames(Config) <- paste0("v", seq_len(N)); colnames(Config) <- paste0("Clone", seq_len(K))
D <- sample_seq_depth(D_input, n_cells = M, n_sites = N)
sim <- sim_read_count(Config, D, cell_num = M)
cores <- parallel::detectCores()
t0 <- Sys.time()
res <- clone_id(sim$A_sim, sim$D_sim, Config = Config, min_iter = T, max_iter = T,
                n_chain = cores, n_proc = cores, verbose = FALSE)
dt <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
cat(sprintf('{"impl": "reference-R", "cells": %d, "variants": %d, "clones": %d, "chains": %d, "iterations_per_s": %.6g}\n',
            M, N, K, cores, cores * (T - 1) / dt))
