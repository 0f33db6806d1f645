"""Dev probe (CPU, numpy): block-Jacobi macro-F1 BCA with commits applied `lag` batches late, against the
sequential oracle.  Used to decide whether a pipelined (lagged) commit keeps the 1e-4 contract."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcolumns_b200.synth import dense_probs
from oracle import oracle as orc


def f1(tp, fp, fn, n, eps=1e-9):
    return 2 * (tp / n) / ((tp + fp) / n + tp / n + fn / n + eps)


def run(eta, k, nb, lag, sweeps=12, seed=0, tol=1e-6):
    n, m = eta.shape
    E = eta.astype(np.float64)
    idx = np.argsort(-eta, axis=1, kind="stable")[:, :k]
    P = np.zeros((n, m), dtype=bool)
    np.put_along_axis(P, idx, True, axis=1)
    rng = np.random.default_rng(seed)
    colsum = E.sum(0)
    def state():
        tp = (E * P).sum(0); fp = ((1 - E) * P).sum(0)
        return tp, fp, colsum - tp
    tp, fp, fn = state()
    utils = [f1(tp, fp, fn, n).mean()]
    B = (n + nb - 1) // nb
    for s in range(sweeps):
        order = rng.permutation(n)
        pend = []  # pending deltas
        for b in range(nb):
            rows = order[b * B:(b + 1) * B]
            while len(pend) > lag:
                d = pend.pop(0); tp = tp + d[0]; fp = fp + d[1]; fn = fn + d[2]
            e = E[rows]; p = P[rows]
            t0 = tp - e * p; f0 = fp - (1 - e) * p; g0 = fn - e * (~p)
            gain = f1(t0 + e, f0 + (1 - e), g0, n) - f1(t0, f0, g0 + e, n)
            new = np.zeros_like(p)
            np.put_along_axis(new, np.argpartition(-gain, k, axis=1)[:, :k], True, axis=1)
            dt = (e * new).sum(0) - (e * p).sum(0)
            df = ((1 - e) * new).sum(0) - ((1 - e) * p).sum(0)
            pend.append((dt, df, -dt))
            P[rows] = new
        for d in pend:
            tp = tp + d[0]; fp = fp + d[1]; fn = fn + d[2]
        tp, fp, fn = state()
        utils.append(f1(tp, fp, fn, n).mean())
        if utils[-1] - utils[-2] < tol:
            break
    return utils


if __name__ == "__main__":
    n, m, k = int(sys.argv[1]), int(sys.argv[2]), 5
    eta = dense_probs(n, m, seed=1003)
    _, meta = orc.predict_using_bc_with_0approx(eta, "f1", k, seed=0, skip_tn=True)
    ref = meta["utilities"][-1]
    print("sequential", len(meta["utilities"]), ref)
    for nb, lag in [(8, 0), (8, 1), (16, 1), (16, 2), (4, 0), (32, 3), (6, 1)]:
        u = run(eta, k, nb, lag)
        print(f"nb={nb} lag={lag}: sweeps={len(u)-1} final={u[-1]:.7f} d={u[-1]-ref:+.2e} first={u[1]:.6f} monotone={all(np.diff(u) > -1e-12)}")
