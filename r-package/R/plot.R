# plot_gibbs: trace plots of a fitted sampler object.  Same arguments, defaults and panels as the reference
# (R/utils.R:114-209): pi (off by default), proportion of observations per cluster, theta per variable.
# Pure R post-processing of the returned list (theta K x P x S, z S x N, pi S x K); nothing here touches
# the GPU.  As in the reference the first retained sample is not drawn and a cluster is shown at a sample
# only where it holds more than `cluster_threshold` of the observations.

.bmm_cluster_props <- function(z, K, cluster_labels) {
    S <- nrow(z)
    counts <- vapply(seq_len(K), function(k) rowSums(z == k), numeric(S))
    if (!is.matrix(counts)) counts <- matrix(counts, nrow = S)
    props <- counts / rowSums(counts)
    long <- data.frame(sample = rep(seq_len(S), times = K),
                       cluster = factor(rep(cluster_labels, each = S), levels = cluster_labels),
                       n = as.vector(counts), prop = as.vector(props))
    long[long$sample != 1 & long$n > 0, , drop = FALSE]
}

plot_gibbs <- function(obj, theta = TRUE, z = TRUE, pi = FALSE, heights = NULL, cluster_threshold = 0.1,
                       cluster_labels = NULL, theta_labels = NULL, theta_to_display = NULL) {
    if (!requireNamespace("ggplot2", quietly = TRUE)) stop("plot_gibbs needs ggplot2")
    if (!requireNamespace("gridExtra", quietly = TRUE)) stop("plot_gibbs needs gridExtra")
    theta_raw <- obj$theta
    z_raw <- obj$z
    K <- dim(theta_raw)[1]
    P <- dim(theta_raw)[2]
    S <- dim(theta_raw)[3]
    panels <- list()

    if (pi) {
        pim <- obj$pi
        if (is.null(pim)) stop("this sampler returns no pi")
        S <- nrow(pim)
        K <- ncol(pim)
        if (is.null(cluster_labels)) cluster_labels <- seq_len(K)
        long <- data.frame(sample = rep(seq_len(S), times = K),
                           cluster = factor(rep(cluster_labels, each = S), levels = cluster_labels),
                           value = as.vector(pim))
        panels[[length(panels) + 1]] <- ggplot2::ggplot(long, ggplot2::aes(x = sample, y = value, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::theme_bw() + ggplot2::labs(x = "Sample", y = "Pi") +
            ggplot2::scale_colour_discrete("Cluster")
    }
    if (is.null(cluster_labels)) cluster_labels <- seq_len(K)

    shown <- NULL          # (sample, cluster) pairs above the threshold; drives the z legend and the theta panel
    if (z || theta) {
        props <- .bmm_cluster_props(z_raw, K, cluster_labels)
        shown <- props[props$prop > cluster_threshold, c("sample", "cluster"), drop = FALSE]
        shown_levels <- unique(as.character(shown$cluster))
    }
    if (z) {
        zp <- props
        zp$cluster <- factor(as.character(zp$cluster), levels = shown_levels)
        panels[[length(panels) + 1]] <- ggplot2::ggplot(zp, ggplot2::aes(x = sample, y = prop, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::theme_bw() + ggplot2::ylim(0, 1) +
            ggplot2::labs(x = "Sample", y = "Proportion in cluster") +
            ggplot2::scale_colour_discrete("Cluster", guide = "none", drop = FALSE)
    }
    if (theta) {
        if (is.null(theta_labels)) theta_labels <- seq_len(P)
        vars <- seq_len(P)
        if (!is.null(theta_to_display)) vars <- match(theta_to_display, theta_labels)
        vars <- vars[!is.na(vars)]
        kidx <- match(as.character(shown$cluster), as.character(cluster_labels))
        long <- do.call(rbind, lapply(vars, function(d)
            data.frame(sample = shown$sample, cluster = factor(as.character(shown$cluster), levels = shown_levels),
                       theta_var = theta_labels[d], value = theta_raw[cbind(kidx, d, shown$sample)])))
        shown_vars <- if (is.null(theta_to_display)) theta_labels else theta_to_display
        long$theta_var <- factor(long$theta_var, levels = shown_vars)
        panels[[length(panels) + 1]] <- ggplot2::ggplot(long, ggplot2::aes(x = sample, y = value, colour = cluster)) +
            ggplot2::geom_line() + ggplot2::facet_wrap(~theta_var) + ggplot2::theme_bw() + ggplot2::ylim(0, 1) +
            ggplot2::labs(x = "Sample", y = "Theta") +
            ggplot2::scale_colour_discrete("Cluster", guide = "none", drop = FALSE)
    }
    gridExtra::grid.arrange(gridExtra::arrangeGrob(grobs = panels, ncol = 1, heights = heights))
}

# histogram of the sampled concentration parameter (un-exported helper, reference R/utils.R:211-215)
plot_alpha <- function(obj) {
    ggplot2::ggplot(data.frame(alpha = as.vector(obj$alpha)), ggplot2::aes(alpha)) +
        ggplot2::geom_histogram(binwidth = 0.1, colour = "black", fill = "white") + ggplot2::xlim(0, 5)
}
