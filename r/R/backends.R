# GPU back-ends for mfBiclust: the bodies of the exported functions on the factorization hot path, rewritten over
# the .Call shim in src/r_shim.c (libmfbiclust_cuda.so, include/mfbiclust.h).  Signatures, return types, messages
# and error behaviour are those of the reference (R/bicluster.R:70-216, 218-289, 528-539; R/bcv.R:31-98, 124-239;
# R/BiclusterStrategy.R:442-456, 508-528), so addStrat(), BiclusterStrategy(), the Shiny app and the testthat file
# stay untouched.  Everything random (runif inits, restart(), sample() folds) is drawn HERE, in R, exactly where
# the reference draws it; the library never draws random numbers.  There is no CPU fallback: without a B200
# `.mfb_ctx()` stops with the library's error, which BiclusterStrategy() turns into its usual warning.
#
# NAMESPACE additions: useDynLib(mfBiclust, .registration = TRUE); the importFrom(NMF, .fcnnls), nipals, MASS and
# EBImage imports of these functions are no longer needed.
#
# R is not installed where this repository is built: this file is reviewed, not executed, there; its .Call
# targets and argument counts are checked against src/r_shim.c by tests/test_r_shim.py.

.mfb_ctx <- local({
    ctx <- NULL
    function() {
        if (is.null(ctx)) ctx <<- .Call(C_mfb_context, as.integer(getOption("mfBiclust.gpu", 0L)))
        ctx
    }
})

.mfb_upload <- function(A) {
    storage.mode(A) <- "double"
    if (identical(getOption("mfBiclust.precision", "f64"), "f32")) .Call(C_mfb_upload_f32, .mfb_ctx(), A)
    else .Call(C_mfb_upload, .mfb_ctx(), A)
}

.normalisedW <- function(m, k) {                                   # R/bicluster.R:95-101
    W <- matrix(runif(m * k), nrow = m, ncol = k)
    frob <- apply(W, 2, function(x) sqrt(sum(x ^ 2)))
    sweep(W, 2, frob, "/")
}

.print_track <- function(trace, converged) {                        # R/bicluster.R:181-201, same text
    for (i in seq_len(ncol(trace))) {
        last <- i == ncol(trace)
        if (i %% 10 == 0 || (last && converged)) {
            cat("Track:\tIter\tNorm\t\trms(dW)\t\trms(dH)\n")
            cat(sprintf("\t%d\t%f\t%f\t%f\n", i, trace[1, i], trace[2, i], trace[3, i]))
        }
    }
    if (converged) message("Converged!")
}

als_nmf <- function(A, k, reps = 4L, maxIter = 100L, eps_conv = 1e-4, verbose = FALSE, ...) {
    m <- nrow(A); n <- ncol(A)
    if (any(A < 0)) stop("Negative values are not allowed")                                      # :76
    if (eps_conv <= 0) stop("ALS-NMF: Invalid argument 'eps_conv' - value should be positive")   # :86
    Ad <- .mfb_upload(A)                                         # once for all replicates
    st <- .Call(C_mfb_stats, Ad)                                 # c(min, max, sumsq, nNA, sum)
    solutions <- lapply(seq_len(reps), function(rep) {
        W <- .normalisedW(m, k)                                                                  # :95-101
        Hold <- matrix(runif(k * n), nrow = k, ncol = n)                                          # :104
        s <- .Call(C_mfb_als_begin, .mfb_ctx(), Ad, as.integer(k), W, Hold, st[2])               # residNormOld <- max(A)
        nrestart <- 0L; i <- 0L; converged <- FALSE
        repeat {
            r <- .Call(C_mfb_als_iterate, s, as.integer(maxIter - i), eps_conv, FALSE)
            i <- r[2]; converged <- r[3] == 1L
            # NMF::.fcnnls consumed r[4] uniforms through max.col(ties.method = "random") in the reference: skip them,
            # so that restart() and the next replicate draw from the stream position they have in the reference
            if (r[4] > 0L) invisible(runif(r[4]))
            if (r[1] == 0L || i >= maxIter) break                # status 0: converged or maxIter reached
            nrestart <- nrestart + 1L                                                            # :124, :138, :153
            if (nrestart >= 10L) {
                warning("als_nmf: Factorization failed too many times [Computation stopped after the 9th restart]")
                break
            }
            .Call(C_mfb_als_restart, s, .normalisedW(m, k))      # restart(): :107-116
        }
        res <- .Call(C_mfb_als_result, s, m, n, as.integer(k), i)
        if (verbose) .print_track(res[[4]], converged)
        list(obj = res[[3]], W = res[[1]], H = res[[2]])
    })
    solution.best <- which.min(vapply(solutions, function(x) x$obj, numeric(1)))                  # :207
    res <- new("genericFactorization", W = solutions[[solution.best]]$W, H = solutions[[solution.best]]$H)
    new("genericFit", fit = res, method = "als-nmf")
}

# What NMF::nmf(x = A, rank = k, method = "snmf/r", beta, eta, rng = .Random.seed) computes (NMF 0.21.0 .nmf_snmf,
# version "R"), on the device.  Seeding is NMF's "random" method (rnmf: basis, then coef, runif(max = max(A))), drawn
# here; the session's RNG is put back on exit as NMF::nmf does.  Signals NMF's own warning / error texts, which the
# handlers of snmf() below match on.
.nmf_snmf <- function(A, k, beta = 0.01, eta = -1, maxIter = 20000L, bi_conv = c(0, 10), eps_conv = 1e-4, ...) {
    m <- nrow(A); n <- ncol(A)
    seed <- .Random.seed
    on.exit(assign(".Random.seed", seed, envir = .GlobalEnv))
    Ad <- .Call(C_mfb_upload, .mfb_ctx(), A)
    W <- matrix(runif(m * k, max = max(A)), m, k)
    invisible(runif(k * n, max = max(A)))
    s <- .Call(C_mfb_snmf_begin, .mfb_ctx(), Ad, as.integer(k), W, beta, eta)
    nrestart <- 0L; i <- 0L
    repeat {
        r <- .Call(C_mfb_snmf_iterate, s, as.integer(maxIter - i), as.numeric(bi_conv), eps_conv)
        i <- r[2]
        if (r[1] == 1) stop("system is computationally singular")          # solve() inside .fcnnls
        if (r[1] == 0 || i >= maxIter) break
        nrestart <- nrestart + 1L                                           # zero row in H
        if (nrestart >= 10L) {
            warning("NMF::snmf - Too many restarts due to too big 'beta' value [Computation stopped after the 9th restart]")
            break
        }
        .Call(C_mfb_snmf_restart, s, matrix(runif(m * k), m, k))
    }
    res <- .Call(C_mfb_als_result, s, m, n, as.integer(k), 0L)
    fit <- new("genericFit", fit = new("genericFactorization", W = res[[1]], H = res[[2]]), method = "snmf/r")
    attr(fit, "parameters") <- list(beta = beta, eta = eta)
    attr(fit, "niter") <- i
    attr(fit, "residuals") <- r[4]
    fit
}

# R/bicluster.R:341-441 with the NMF::nmf call (:361-367) replaced by .nmf_snmf; the beta / eta search is the
# reference's own
snmf <- function(A, k, verbose = TRUE, ...) {
    eta <- list(...)$eta
    beta <- list(...)$beta
    args <- list(...)
    args <- c(maxIter = args$maxIter, bi_conv = list(args$bi_conv), eps_conv = args$eps_conv)
    args <- args[!vapply(args, is.null, logical(1))]
    res <- NULL
    if (is.null(beta)) beta <- mean(A)                                                            # :351
    if (is.null(eta)) eta <- mean(A)                                                              # :352
    if (eta < 0 || beta < .Machine$double.eps) stop("eta must be >= 0 and beta must be > 0")      # :353-355
    while (eta >= 0 && beta >= .Machine$double.eps && !inherits(res, "genericFit")) {             # :357
        res <- tryCatch({
            res <- do.call(.nmf_snmf, c(list(A = A, k = k, beta = beta, eta = eta), args))
            if (verbose) {
                cat(paste("method:", res@method, "\n"))
                cat(paste("parameters:\nbeta:", beta, "\neta:", eta, "\n"))
            }
            res
        },
        warning = function(w) {                                                                   # :378-394
            if (grepl("too big 'beta' value", w$message, fixed = TRUE)) {
                beta <<- beta / 2
                if (verbose) cat(paste("mfBiclust::snmf: Restarting with beta = ", beta, "\n"))
            } else {
                warning(w)
            }
            NULL
        },
        error = function(e) {                                                                     # :395-429
            if (grepl("system is computationally singular", e[[1]], fixed = TRUE)) {
                eta <<- if (eta > .Machine$double.eps) eta / 2 else 0
                if (verbose) cat(paste("mfBiclust::snmf: Restarting with eta = ", eta, "\n"))
            } else {
                stop(e)
            }
            NULL
        })
    }
    if (inherits(res, "genericFit")) {                                                            # :432-435
        res@method <- "snmf"
        return(res)
    }
    res <- new("genericFactorization", W = matrix(data = numeric()), H = matrix(numeric()))       # :436-440
    new("genericFit", fit = res, method = "snmf")
}

svd_pca <- function(A, k, ...) {                                   # R/bicluster.R:528-539
    Ad <- .Call(C_mfb_upload, .mfb_ctx(), A)
    res <- .Call(C_mfb_svd_pca, .mfb_ctx(), Ad, as.integer(k), nrow(A), ncol(A))
    W <- res[[1]]; H <- res[[2]]
    dimnames(W) <- list(rownames(A), paste0("PC", seq_len(ncol(W))))   # prcomp's names
    dimnames(H) <- list(paste0("PC", seq_len(nrow(H))), colnames(A))
    new("genericFit", fit = new("genericFactorization", H = H, W = W), method = "svd-pca")
}

nipals_pca_nocatch <- function(A, k, ...) {                        # R/bicluster.R:218-242
    args <- list(...)
    center <- if (is.null(args$center)) TRUE else args$center      # nipals::nipals defaults
    scale <- if (is.null(args$scale)) TRUE else args$scale
    maxiter <- if (is.null(args$maxiter)) 500L else args$maxiter
    gramschmidt <- if (is.null(args$gramschmidt)) TRUE else args$gramschmidt
    if (any(colSums(!is.na(A)) == 0) || any(rowSums(!is.na(A)) == 0))
        stop("At least one column or row is all NAs")              # nipals' own check
    Ad <- .Call(C_mfb_upload, .mfb_ctx(), A)                       # NaN = NA: the library's has.na branch
    np <- .Call(C_mfb_nipals, .mfb_ctx(), Ad, nrow(A), ncol(A), as.integer(k), center, scale, as.integer(maxiter),
                1e-6, gramschmidt)
    if (np[[5]] == 3L) warning(paste("Stopping after", maxiter, "iterations for a principal component"))
    new("genericFit", fit = new("genericFactorization", W = np[[1]], H = t(np[[2]])), method = "nipals-pca")
}

nipals_pca <- function(A, k, cleanParam = 0, verbose = TRUE, ...) {   # R/bicluster.R:269-289
    cleanRes <- clean(A, cleanParam, dimsRemain = TRUE)
    list(m = cleanRes$obj, genericFit = nipals_pca_nocatch(cleanRes$obj, k, ...),
         indexRemaining = cleanRes$dimsRemain)
}

# bcvGivenKs for several fold draws at once (folds are drawn by the caller with sample(), R/bcv.R:170-174)
.bcv_cells <- function(Yd, rowFolds, colFolds, ks, holdouts, na = FALSE) {
    nf <- ncol(rowFolds)
    err <- if (na) {                                              # any(is.na(Y)): NIPALS-based pca, R/bcv.R:177-181
        .Call(C_mfb_bcv_cells_na, .mfb_ctx(), Yd, rowFolds, colFolds, as.integer(holdouts), as.integer(ks),
              0L, as.integer(nf * holdouts ^ 2))
    } else {
        .Call(C_mfb_bcv_cells_multi, .mfb_ctx(), Yd, rowFolds, colFolds, as.integer(holdouts), as.integer(ks),
              0L, as.integer(nf * holdouts ^ 2))                   # nks x cells
    }
    vapply(seq_len(nf), FUN.VALUE = numeric(length(ks)), FUN = function(f) {
        rowSums(err[, (f - 1L) * holdouts ^ 2 + seq_len(holdouts ^ 2), drop = FALSE])            # :231-233
    })
}

.draw_folds <- function(n, p, holdouts) {                           # R/bcv.R:170-174, same two sample() calls
    list(rows = sample((seq_len(n) %% holdouts) + 1L, size = n), cols = sample((seq_len(p) %% holdouts) + 1L, size = p))
}

bcvGivenKs <- function(Y, ks, holdouts = 10L, Yd = .Call(C_mfb_upload, .mfb_ctx(), Y)) {
    f <- .draw_folds(nrow(Y), ncol(Y), holdouts)
    rcvs <- .bcv_cells(Yd, matrix(f$rows, ncol = 1L), matrix(f$cols, ncol = 1L), ks, holdouts, na = anyNA(Y))[, 1L]
    names(rcvs) <- as.character(ks)
    rcvs
}

bcv <- function(Y, ks = seq_len(min(nrow(Y), ncol(Y)) - 1), holdouts = 10L, interactive = TRUE) {   # R/bcv.R:124-159
    if (!inherits(Y, "matrix")) Y <- as.matrix(Y)
    if (any(is.na(Y)) && interactive) {
        message(paste("Since some elements of the matrix are NA, BCV will run", "a much slower iterative algorithm. Continue?"))
        if (menu(c("Yes", "No")) != 1) stop("User choice")
    }
    minDim <- min(ncol(Y), nrow(Y))
    ks <- kLimiter(holdouts, minDim, ks)
    if (length(ks) > 1 || is.numeric(ks)) {
        if (any(ks < 1) || any(!is.wholenumber(ks))) stop(paste("ks must be a vector of positive whole numbers"))
    } else {
        stop(paste("ks must be a vector of positive whole numbers"))
    }
    if (holdouts > minDim) {
        holdouts <- minDim
        warning(paste("Using", holdouts, "holdouts to ensure non-zero holdout", "size."))
    }
    result.list <- bcvGivenKs(Y, ks, holdouts)
    names(result.list) <- as.character(ks)
    result.list
}

auto_bcv <- function(Y, ks = seq_len(min(nrow(Y), ncol(Y)) - 1), holdouts = 3L, maxIter = 100L, tol = (10 ^ -4),
                     bestOnly = FALSE, verbose = FALSE, interactive = TRUE) {                      # R/bcv.R:31-98
    if (!inherits(Y, "matrix")) Y <- as.matrix(Y)
    na <- anyNA(Y)
    if (na && interactive) {
        message(paste("Since some elements of the matrix are NA, BCV will run", "a much slower iterative algorithm. Continue?"))
        if (menu(c("Yes", "No")) != 1) stop("User choice")
    }
    ks <- kLimiter(holdouts, min(ncol(Y), nrow(Y)), ks)
    Yd <- .Call(C_mfb_upload, .mfb_ctx(), Y)                       # resident for every bcv() call
    distr <- rep(1L, each = length(ks)); names(distr) <- as.character(ks)
    distrOld <- distr
    i <- 0L; converged <- FALSE
    n <- nrow(Y); p <- ncol(Y)
    # The fold draws do not depend on the data: the iterations the stopping rule provably cannot do without
    # (the rule needs ((D+t)/(S+t) - (D+t-1)/(S+t-1))^2 < tol for the leading count D) are drawn in order and their
    # cells computed as ONE batch; the rule is then replayed iteration by iteration exactly as the reference does.
    # Draws made for iterations after the one that converges are the only difference in RNG consumption.
    earliest <- function() {
        S <- sum(distr); D <- max(distr)
        for (t in seq_len(maxIter - i)) if (((D + t) / (S + t) - (D + t - 1) / (S + t - 1)) ^ 2 < tol) return(t)
        maxIter - i
    }
    while (!converged && i < maxIter) {
        nb <- max(1L, earliest())
        folds <- lapply(seq_len(nb), function(b) .draw_folds(n, p, holdouts))
        vals <- .bcv_cells(Yd, vapply(folds, function(f) f$rows, integer(n)), vapply(folds, function(f) f$cols, integer(p)),
                           ks, holdouts, na = na)                  # length(ks) x nb
        for (b in seq_len(nb)) {
            res <- vals[, b]; names(res) <- as.character(ks)
            bcvRes <- names(which.min(res))
            distr[bcvRes] <- distr[bcvRes] + 1L
            resid <- unlist(mapply(function(d, dOld) (d / sum(distr) - dOld / sum(distrOld)) ^ 2, d = distr, dOld = distrOld))
            if (verbose) {
                cat(paste("Iteration", i + 1L, "\n"))
                cat(paste("Current best:", names(which.max(distr)), "\n"))
                cat("BCV result distribution:\n")
                message(paste0(do.call(paste, as.list(c(as.list(names(distr)), sep = "\t"))), "\n",
                               do.call(paste, as.list(c(as.list(distr - 1L), sep = "\t")))))
            }
            converged <- all(resid < tol)
            distrOld <- distr
            i <- i + 1L
            if (converged) break
        }
    }
    if (i == maxIter) warning("BCV results did not converge after", maxIter, "iterations")
    distr <- distr - 1
    med <- as.numeric(names(which.max(distr)))
    if (bestOnly) med else list(best = med, counts = distr)
}

otsuHelper <- function(matrix) .Call(C_mfb_otsu, .mfb_ctx(), matrix)     # R/BiclusterStrategy.R:442-456

threshold <- function(m, th, MARGIN = 2) {                                # R/BiclusterStrategy.R:508-528
    if (MARGIN == 1) {
        if (length(th) != nrow(m)) stop("Length of th must equal nrow(m).")
    } else {
        if (length(th) != ncol(m)) stop("Length of th must equal ncol(m)")
    }
    mat <- .Call(C_mfb_threshold, .mfb_ctx(), m, as.numeric(th), as.integer(MARGIN))
    dimnames(mat) <- dimnames(m)
    mat
}

# ---- post-processing (R/helperFunctions.R:150-346): the pairwise intersections and the union come from the device,
# ---- the greedy loop of filter.biclust is the reference's own
sizes <- function(biclusterRows, biclusterCols)                            # R/helperFunctions.R:303-308 (index lists: host)
    vapply(seq_along(biclusterRows), FUN.VALUE = numeric(1),
           FUN = function(b) length(biclusterRows[[b]]) * length(biclusterCols[[b]]))

union.biclust <- function(RowxBiclust, BiclustxCol, choose = rep(TRUE, ncol(RowxBiclust))) {   # :336-345
    if (ncol(RowxBiclust) == 0) return(matrix(0, nrow = nrow(RowxBiclust), ncol = ncol(BiclustxCol)))
    storage.mode(RowxBiclust) <- "logical"; storage.mode(BiclustxCol) <- "logical"
    .Call(C_mfb_bicluster_union, .mfb_ctx(), RowxBiclust, BiclustxCol, as.logical(choose))
}

filter.biclust <- function(rowxBicluster, biclusterxCol, max = NULL, overlap = 0.25) {         # :150-206
    if (ncol(rowxBicluster) != nrow(biclusterxCol))
        stop(paste0("RowxBicluster must have the same number of columns as", "BiclusterxCol"))
    k <- ncol(rowxBicluster)
    if (k == 0) {
        chosen <- rep(FALSE, ncol(rowxBicluster))
    } else if (k == 1) {
        chosen <- rep(TRUE, ncol(rowxBicluster))
    } else {
        storage.mode(rowxBicluster) <- "logical"; storage.mode(biclusterxCol) <- "logical"
        so <- .Call(C_mfb_bicluster_overlap, .mfb_ctx(), rowxBicluster, biclusterxCol)
        sizes <- so[[1]]; names(sizes) <- seq_len(k)
        overlaps <- so[[2]]
        chosen <- rep(FALSE, each = k)
        pool <- rep(TRUE, each = k)
        pool[sizes == 0 | sizes == nrow(rowxBicluster) * ncol(biclusterxCol)] <- FALSE
        while (all(sum(chosen) < max) && sum(pool) > 0) {
            chooseMe <- as.numeric(names(which.max(sizes[which(pool)])))
            chosen[chooseMe] <- TRUE
            pool[overlaps[, chooseMe] > overlap] <- FALSE
            pool[chooseMe] <- FALSE
        }
    }
    biclustered <- union.biclust(rowxBicluster, biclusterxCol, chosen)
    list(RowxBicluster = rowxBicluster[, chosen, drop = FALSE], BiclusterxCol = biclusterxCol[chosen, , drop = FALSE],
         biclustered = biclustered, chosen = chosen)
}
