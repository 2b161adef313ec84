# User-facing samplers.  Argument lists, defaults and the returned list are those of bmmmcmc 1.0
# (reference R/utils.R:23-107); the initial state is drawn here with R's RNG exactly as the reference
# does, then handed to the CUDA back end through the Rcpp host (src/host.cpp).
#
# Back-end settings that have no place in the reference's argument lists are R options:
#   options(bmm.seed = NULL)              Philox key; NULL draws it from R's RNG (set.seed() reproducible)
#   options(bmm.device = 0L)              CUDA device
#   options(bmm.stephens_fixed = FALSE)   TRUE: corrected Stephens relabelling (inverse permutation for the
#                                         column re-ordering, log p in the online cost, running-mean Q)
#                                         instead of the reference's (stephens.cpp:56,79,85,92)

.bmm_defaults <- function(nsamples, burnin, burnrelabel, alpha, clamp = TRUE) {
    if (is.null(burnin)) burnin <- round(0.1 * nsamples)
    if (clamp && burnrelabel > burnin) burnrelabel <- round(0.1 * burnin)
    list(burnin = burnin, burnrelabel = burnrelabel, alpha = if (is.null(alpha)) 0 else alpha)
}

.bmm_initial_weights <- function(K) {
    w <- exp(stats::runif(K))
    w / sum(w)
}

gibbs_dp <- function(data, nsamples, alpha = NULL, a = 1, b = 1, beta = 0.5, gamma = 0.5,
                     burnin = NULL, relabel = FALSE, burnrelabel = 50, maxK = 30, debug = FALSE) {
    d <- .bmm_defaults(nsamples, burnin, burnrelabel, alpha)
    collapsed_gibbs_dp_cpp(data, nsamples, d$alpha, beta, gamma, a, b, d$burnin, relabel,
                           d$burnrelabel, maxK, debug)
}

gibbs_collapsed <- function(data, nsamples, K, alpha = NULL, beta = 0.5, gamma = 0.5, a = 1, b = 1,
                            burnin = NULL, relabel = FALSE, burnrelabel = 50, debug = FALSE) {
    d <- .bmm_defaults(nsamples, burnin, burnrelabel, alpha)
    start <- sample(1:K, nrow(data), replace = TRUE)
    collapsed_gibbs_cpp(data, start, nsamples, K, d$alpha, beta, gamma, a, b, d$burnin, relabel,
                        d$burnrelabel, debug)
}

gibbs_full <- function(data, nsamples, K, alpha = NULL, beta = 0.5, gamma = 0.5, a = 1, b = 1,
                       burnin = NULL, relabel = FALSE, burnrelabel = 50, debug = FALSE) {
    d <- .bmm_defaults(nsamples, burnin, burnrelabel, alpha)
    w0 <- .bmm_initial_weights(K)
    theta0 <- matrix(stats::runif(K * ncol(data)), nrow = K, ncol = ncol(data))
    gibbs_cpp(data, w0, theta0, nsamples, K, d$alpha, beta, gamma, a, b, d$burnin, relabel,
              d$burnrelabel, debug)
}

# The reference wrapper does not clamp burnrelabel for this sampler (R/utils.R:95-107); kept.
gibbs_stickbreaking <- function(data, nsamples, maxK, alpha = NULL, beta = 0.5, gamma = 0.5, a = 1, b = 1,
                                burnin = NULL, relabel = FALSE, burnrelabel = 50, debug = FALSE) {
    d <- .bmm_defaults(nsamples, burnin, burnrelabel, alpha, clamp = FALSE)
    w0 <- .bmm_initial_weights(maxK)
    theta0 <- matrix(stats::runif(maxK * ncol(data)), nrow = maxK, ncol = ncol(data))
    gibbs_stickbreaking_cpp(data, w0, theta0, nsamples, maxK, d$alpha, beta, gamma, a, b, d$burnin,
                            relabel, d$burnrelabel, debug)
}

#' Posterior predictive distribution of new observations
#'
#' Not in the reference (its TODO file lists "Implement predictive distribution"): for every row of
#' \code{newdata}, the log of the posterior predictive probability averaged over the kept draws of an
#' uncollapsed sampler, and the cluster responsibilities averaged over the draws.
#'
#' @param obj The list returned by \code{gibbs_full} or \code{gibbs_stickbreaking}.
#' @param newdata An M x P 0/1 integer matrix.
#' @return A list with \code{log_pred} (length M) and \code{membership} (M x K).
#' @export
predict_gibbs <- function(obj, newdata) {
    if (is.null(obj$pi)) stop("predict_gibbs needs the pi history of gibbs_full / gibbs_stickbreaking")
    newdata <- as.matrix(newdata)
    storage.mode(newdata) <- "integer"
    predictive_cpp(newdata, obj$theta, obj$pi)
}
