# plot_gibbs(obj, theta = TRUE, z = TRUE, pi = TRUE, heights = NULL, cluster_threshold = 0.1):
# trace plots of a fitted sampler object, same arguments as the reference (R/utils.R:114-209).
# Pure R post-processing of the returned list (theta K x P x S, z S x N, pi S x K); nothing here
# touches the GPU.  Clusters whose mean allocation share is below `cluster_threshold` are dropped.

plot_gibbs <- function(obj, theta = TRUE, z = TRUE, pi = TRUE, heights = NULL, cluster_threshold = 0.1) {
    if (!requireNamespace("ggplot2", quietly = TRUE)) stop("plot_gibbs needs ggplot2")
    S <- nrow(obj$z)
    K <- dim(obj$theta)[1]
    share <- vapply(seq_len(K), function(k) mean(obj$z == k), numeric(1))
    keep <- which(share > cluster_threshold)
    panels <- list()
    if (theta) {
        P <- dim(obj$theta)[2]
        long <- do.call(rbind, lapply(keep, function(k) do.call(rbind, lapply(seq_len(P), function(d)
            data.frame(sample = seq_len(S), cluster = factor(k), variable = factor(d), value = obj$theta[k, d, ])))))
        panels$theta <- ggplot2::ggplot(long, ggplot2::aes(sample, value, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::facet_wrap(~variable) + ggplot2::ylim(0, 1) +
            ggplot2::labs(title = "theta", y = NULL) + ggplot2::theme_bw()
    }
    if (z) {
        long <- do.call(rbind, lapply(keep, function(k)
            data.frame(sample = seq_len(S), cluster = factor(k), value = rowMeans(obj$z == k))))
        panels$z <- ggplot2::ggplot(long, ggplot2::aes(sample, value, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::ylim(0, 1) + ggplot2::labs(title = "z (share of observations)", y = NULL) +
            ggplot2::theme_bw()
    }
    if (pi && !is.null(obj$pi)) {
        long <- do.call(rbind, lapply(keep, function(k)
            data.frame(sample = seq_len(S), cluster = factor(k), value = obj$pi[, k])))
        panels$pi <- ggplot2::ggplot(long, ggplot2::aes(sample, value, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::ylim(0, 1) + ggplot2::labs(title = "pi", y = NULL) + ggplot2::theme_bw()
    }
    if (length(panels) == 1) return(panels[[1]])
    if (!requireNamespace("gridExtra", quietly = TRUE)) return(panels)
    if (is.null(heights)) heights <- rep(1, length(panels))
    gridExtra::grid.arrange(grobs = panels, ncol = 1, heights = heights)
}
