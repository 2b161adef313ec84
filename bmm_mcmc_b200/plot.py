"""plot_gibbs for the Python host: same arguments, defaults and panels as the reference's R function
(R/utils.R:114-209) -- pi traces (off by default), proportion of observations per cluster, theta traces per variable.
Pure post-processing of the returned list (theta K x P x S, z S x N, pi S x K); nothing here touches the GPU.

As in the reference the first retained sample is not drawn, and a cluster is shown at a sample only where it holds
more than `cluster_threshold` of the observations.  The panel data are returned as pandas DataFrames (long format, the
frames the reference pipes into ggplot2); when matplotlib is importable the figure is drawn as well.
"""
import numpy as np


def _one_chain(obj):
    z = np.asarray(obj["z"])
    if z.ndim == 3:
        if z.shape[0] != 1:
            raise ValueError("plot_gibbs draws one chain: index the chain first (e.g. {k: v[0] for k, v in obj.items()})")
        return {k: np.asarray(v)[0] for k, v in obj.items()}
    return {k: np.asarray(v) for k, v in obj.items()}


def plot_gibbs(obj, theta=True, z=True, pi=False, heights=None, cluster_threshold=0.1, cluster_labels=None,
               theta_labels=None, theta_to_display=None, draw=True):
    import pandas as pd
    obj = _one_chain(obj)
    theta_raw, z_raw = obj["theta"], obj["z"]
    K, P, S = theta_raw.shape
    panels = {}
    if pi:
        if "pi" not in obj:
            raise ValueError("this sampler returns no pi")
        pim = obj["pi"]
        S, K = pim.shape
        if cluster_labels is None:
            cluster_labels = list(range(1, K + 1))
        panels["pi"] = pd.DataFrame({"sample": np.tile(np.arange(1, S + 1), K),
                                     "cluster": np.repeat(np.asarray(cluster_labels, dtype=object), S),
                                     "value": pim.T.ravel()})
    if cluster_labels is None:
        cluster_labels = list(range(1, K + 1))
    labels = np.asarray(cluster_labels, dtype=object)
    shown = None
    if z or theta:
        counts = np.stack([(z_raw == k + 1).sum(1) for k in range(K)], axis=1)          # S x K
        props = counts / counts.sum(1, keepdims=True)
        frame = pd.DataFrame({"sample": np.tile(np.arange(1, S + 1), K), "cluster": np.repeat(labels, S),
                              "n": counts.T.ravel(), "prop": props.T.ravel()})
        frame = frame[(frame["sample"] != 1) & (frame["n"] > 0)].reset_index(drop=True)
        shown = frame.loc[frame["prop"] > cluster_threshold, ["sample", "cluster"]].reset_index(drop=True)
        shown_levels = list(pd.unique(shown["cluster"]))
    if z:
        panels["z"] = frame[frame["cluster"].isin(shown_levels)].reset_index(drop=True)
    if theta:
        if theta_labels is None:
            theta_labels = list(range(1, P + 1))
        tl = list(theta_labels)
        vars_ = list(range(P)) if theta_to_display is None else [tl.index(v) for v in theta_to_display if v in tl]
        lab_to_k = {lab: k for k, lab in enumerate(cluster_labels)}
        kidx = shown["cluster"].map(lab_to_k).to_numpy(dtype=int)
        sidx = shown["sample"].to_numpy(dtype=int) - 1
        rows = [pd.DataFrame({"sample": shown["sample"], "cluster": shown["cluster"], "theta_var": tl[d],
                              "value": theta_raw[kidx, d, sidx]}) for d in vars_]
        panels["theta"] = pd.concat(rows, ignore_index=True) if rows else pd.DataFrame(columns=["sample", "cluster", "theta_var", "value"])
    if draw:
        try:
            import matplotlib.pyplot as plt
        except Exception:
            return panels
        order = [k for k in ("pi", "z", "theta") if k in panels]
        fig, axes = plt.subplots(len(order), 1, squeeze=False,
                                 gridspec_kw={"height_ratios": heights} if heights is not None else None)
        for ax, name in zip(axes[:, 0], order):
            df = panels[name]
            ycol = "prop" if name == "z" else "value"
            keys = ["cluster"] if name != "theta" else ["cluster", "theta_var"]
            for key, grp in df.groupby(keys, sort=False):
                ax.plot(grp["sample"], grp[ycol], label=str(key))
            ax.set_xlabel("Sample")
            ax.set_ylabel({"pi": "Pi", "z": "Proportion in cluster", "theta": "Theta"}[name])
            if name != "pi":
                ax.set_ylim(0, 1)
        panels["figure"] = fig
    return panels


def plot_alpha(obj, draw=True):
    """Histogram data of the sampled concentration parameter (reference R/utils.R:211-215: bin width 0.1 on [0, 5])."""
    alpha = np.asarray(_one_chain(obj)["alpha"]).ravel()
    hist, edges = np.histogram(alpha, bins=np.arange(0.0, 5.0 + 1e-9, 0.1))
    return hist, edges
