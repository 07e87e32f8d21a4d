# Replacement bodies for clone_id_Gibbs / clone_id_EM (R/clone_id.R:351-675, :240-326) that call the
# B200 library through the .Call shim (r/cardelino_b200_shim.c).  Signatures, defaults, stop() messages
# and the returned list are the reference's; clone_id() itself (R/clone_id.R:78-186) stays as it is except
# that the doMC branch (:127-154) must go: a forked worker cannot use the parent's CUDA context, so
# n_chain is handed to the library (argument n_chain below) and n_proc is ignored.
# NOT run in the build image (no R there); the shim underneath is executed by tests/test_r_shim.py.
#
# Extra arguments (all optional, defaults keep the reference's behaviour):
#   n_chain  chains run inside the library (replaces the doMC fork)
#   light    NULL (default): TRUE when the dense parts of the reference's return list could not be held in R --
#            N * M > 2e8 (theta1, prob_variant) or max_iter * M * K doubles above 2 GB (prob_all) -- else FALSE.
#            light = TRUE leaves out prob_all, Config_all, prob_variant and the dense theta1 (theta1_elements holds
#            the per-variant / global / per-entry means instead); the Geweke stop rule then runs inside the library
#            from running sums (same rule, same stop iteration; `stop_rule` in the result says how it was applied).
#            relabel = TRUE needs the prob_all trace, i.e. light = FALSE.

.b200_env <- new.env()
.b200_handle <- function() {
    if (is.null(.b200_env$h)) .b200_env$h <- .Call("b200_create", 0L)
    .b200_env$h
}

.b200_load <- function(h, A, D) {
    if (methods::is(D, "dgCMatrix")) {
        A <- methods::as(A, "dgCMatrix")
        .Call("b200_load_dgc", h, dim(D), D@p, D@i, D@x, A@p, A@i, A@x)
    } else {
        storage.mode(A) <- "double"; storage.mode(D) <- "double"
        .Call("b200_load_dense", h, A, D)
    }
}

clone_id_Gibbs <- function(A, D, Config, Psi = NULL,
    relax_Config = TRUE, relax_rate_fixed = NULL,
    relax_rate_prior = c(1, 9), keep_base_clone = TRUE,
    prior0 = c(0.2, 99.8), prior1 = c(0.45, 0.55),
    min_iter = 5000, max_iter = 20000, buin_frac = 0.5,
    wise = "variant", relabel = FALSE, verbose = TRUE, n_chain = 1, light = NULL) {
    if (is.null(Psi)) Psi <- rep(1 / ncol(Config), ncol(Config))
    if (dim(A)[1] != dim(D)[1] || dim(A)[2] != dim(D)[2] ||
        dim(A)[1] != dim(Config)[1] || dim(Config)[2] != length(Psi)) {
        stop("A and D must have the same size;\n ",
             "A and Config must have the same number of variants;\n",
             "Config and Psi must have the same number of clones")
    }
    if (sum(c("element", "variant", "global") == wise) == 0) {
        stop("Input wise mode: ", wise, ", while only supporting: element, variant, global")
    }
    if (!is.null(relax_Config) && relax_Config != FALSE && !is.null(relax_rate_fixed) &&
        (relax_rate_fixed > 1 || relax_rate_fixed < 0)) {
        stop("relax_rate_fixed needs to be NULL or in [0, 1].")
    }
    if (!is.matrix(prior1) && length(prior1) != 2) {
        stop("prior1 need to be a matrix of n_element x 2")
    }
    N <- nrow(Config); M <- ncol(D); K <- ncol(Config)
    if (is.null(light)) {
        light <- !relabel && (as.double(N) * M > 2e8 || as.double(max_iter) * M * K * 8 > 2^31)
    }
    if (light && relabel) stop("relabel needs the prob_all trace: use light = FALSE")
    h <- .b200_handle()
    .b200_load(h, A, D)
    storage.mode(Config) <- "double"
    if (is.matrix(prior1)) storage.mode(prior1) <- "double" else prior1 <- as.double(prior1)
    params <- list(relax_Config = !is.null(relax_Config) && relax_Config != FALSE,
                   relax_rate_fixed = relax_rate_fixed, relax_rate_prior = as.double(relax_rate_prior),
                   keep_base_clone = keep_base_clone, prior0 = as.double(prior0), prior1 = prior1,
                   min_iter = as.integer(min_iter), max_iter = as.integer(max_iter), buin_frac = buin_frac,
                   wise = wise, n_chain = as.integer(n_chain),
                   seed = sample.int(.Machine$integer.max, 1),   # ties the Philox key to set.seed()
                   relabel = isTRUE(relabel),    # R/clone_id.R:624-644, done on the device before the summaries
                   verbose = isTRUE(verbose),    # "x% converged." at every check, printed by the shim (Rprintf)
                   light = isTRUE(light))
    # user interrupts are polled every 100 sweeps inside the call; an interrupt ends it with an R error and leaves
    # the handle usable
    chains <- .Call("b200_gibbs", h, Config, as.double(Psi), params)
    finish <- function(ch) {
        it <- ch$n_iter
        if (is.na(ch$percent_converged) || ch$percent_converged <= 95) {
            warning("Only ", ch$percent_converged, "% of parameters converged. ",
                    "Consider increasing `max_iter`.", call. = FALSE)
        } else print(paste("Converged in", it, "iterations."))
        prob <- ch$prob; rownames(prob) <- colnames(A); colnames(prob) <- colnames(Config)
        Config_prob <- Config; Config_prob[, ] <- ch$Config_prob
        dic <- as.list(ch$DIC); names(dic) <- c("DIC", "logLik_var", "D_mean", "D_post", "DIC_Gelman", "DIC_Spiegelhalter")
        cat(paste("DIC:", round(dic$DIC, 2), "D_mean:", round(dic$D_mean, 2), "D_post:", round(dic$D_post, 2),
                  "logLik_var:", round(dic$logLik_var, 2), "\n"))
        if (light) {
            return(list(theta0 = ch$theta0, theta1_elements = ch$theta1_elements, theta0_all = ch$theta0_all,
                        theta1_all = ch$theta1_all, logLik = ch$logLik, prob = prob, relax_rate = ch$relax_rate,
                        Config_prob = Config_prob, relax_rate_all = ch$relax_rate_all, DIC = dic,
                        stop_rule = ch$stop_rule))
        }
        theta1 <- matrix(NA, N, M)
        element <- if (wise == "element") ch$element else seq_len(N * M)     # idx_mat, R/clone_id.R:406-416
        theta1[element] <- ch$theta1_elements                    # recycling, R/clone_id.R:469,653
        pv <- ch$prob_variant; dimnames(pv) <- dimnames(A)
        list(theta0 = ch$theta0, theta1 = theta1, theta0_all = ch$theta0_all, theta1_all = ch$theta1_all,
             element = element, logLik = ch$logLik, prob_all = t(ch$prob_all_t), prob = prob,
             prob_variant = pv, relax_rate = ch$relax_rate, Config_prob = Config_prob,
             Config_all = t(ch$Config_all_t), relax_rate_all = ch$relax_rate_all, DIC = dic)
    }
    outs <- lapply(chains, finish)
    if (n_chain == 1) outs[[1]] else outs          # clone_id() merges the list with colMatch (:165-181)
}

clone_id_EM <- function(A, D, Config, Psi = NULL, min_iter = 10, max_iter = 1000,
    logLik_threshold = 1e-5, verbose = TRUE) {
    if (is.null(Psi)) Psi <- rep(1 / ncol(Config), ncol(Config))
    if (dim(A)[1] != dim(D)[1] || dim(A)[2] != dim(D)[2] ||
        dim(A)[1] != dim(Config)[1] || dim(Config)[2] != length(Psi)) {
        stop("A and D must have the same size;\n",
             "A and Config must have the same number of variants;\n",
             "Config and Psi must have the same number of clones")
    }
    h <- .b200_handle()
    .b200_load(h, A, D)
    storage.mode(Config) <- "double"
    theta_init <- c(stats::runif(1, 0.001, 0.25), stats::runif(1, 0.3, 0.6))     # R/clone_id.R:278
    out <- .Call("b200_em", h, Config, as.double(Psi), as.integer(min_iter), as.integer(max_iter),
                 logLik_threshold, theta_init)
    if (verbose) print(paste("Total iterations:", out$n_iter))
    prob <- out$prob; rownames(prob) <- colnames(A); colnames(prob) <- colnames(Config)
    list(theta = out$theta, prob = prob, logLik = out$logLik)
}
