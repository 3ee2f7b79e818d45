"""numpy model of K1k (gaselect_b200/csrc/simpls_kspace.cu): the reference's SIMPLS recurrences
(src/PLSSimpls.cpp:106-166) carried by vectors over the ROWS, with K = Z Z' of the globally centred selected columns,
inside PLSEvaluator::estSEP's loops (src/PLSEvaluator.cpp:228-343).  Test infrastructure: the CUDA kernel follows
exactly these formulas; tests/test_kspace_model.py checks them against the oracle."""
import numpy as np


def fit_residuals(K, yg, train, tests, A):
    """Residuals y - yhat on each held-out row set of `tests` for 1..A components of the fit on rows `train`."""
    n = K.shape[0]
    m = np.zeros(n, bool)
    m[train] = True
    ym = yg[m].mean()
    ycl = yg - ym                                 # response centred on the training rows (src/PLSSimpls.cpp:28-49)
    yc = np.where(m, ycl, 0.0)
    t_all = K @ yc                                # scores of every row: Z S_0, S_0 = X_c'y_c (:126)
    U = []
    pred = np.zeros(n)
    out = []
    for a in range(A):
        mu = t_all[m].mean()                      # column-centring of X applied to t = X S (:135)
        tcu = np.where(m, t_all - mu, 0.0)
        rtn = 1.0 / np.sqrt((tcu * tcu).sum())    # :138
        q = (yc * tcu).sum() * rtn                # :148
        pred = pred + (t_all - mu) * (q * rtn)    # cumulative coefficients S q / |t| (:152-156), src/PLS.cpp:34-47
        out.append([ycl[te] - pred[te] for te in tests])
        if a + 1 == A:
            break
        tc = tcu * rtn
        w = K @ tc                                # Z v for the loading v = X't (:145-147)
        c = [Uj @ tc for Uj in U]                 # V_j'v (:57-63)
        u = w - sum((cj * Uj for cj, Uj in zip(c, U)), np.zeros(n))
        vv = tc @ w - sum(cj * cj for cj in c)    # |v - sum_j (V_j'v) V_j|^2
        vs = tc @ t_all                           # v'S
        U.append(u / np.sqrt(vv))
        t_all = t_all - u * (vs / vv)             # S <- S - v (v'S)/(v'v) (:70-88, :160)
    return out


def evaluate(X, y, cols, seg, R, O, I, maxNComp, sdfact):
    """PLSEvaluator::evaluate for one subset; `seg` = the reference's segmentation (train, test, train, test, ...)."""
    Z = (X - X.mean(axis=0))[:, np.asarray(cols, dtype=np.int64)]
    yg = y - y.mean()
    K = Z @ Z.T
    A = min(maxNComp, Z.shape[1])
    total, s = 0.0, 0
    for _ in range(R):
        pooled = []
        for _ in range(O):
            msep = np.zeros((I, A))
            for f in range(I):
                res = fit_residuals(K, yg, seg[s], [seg[s + 1]], A)
                s += 2
                msep[f] = [(res[a][0] ** 2).sum() / len(res[a][0]) for a in range(A)]
            mean, sd = msep.mean(axis=0), msep.std(axis=0, ddof=1)
            lo = int(np.argmin(mean))
            cutoff = mean[lo] + sd[lo] * sdfact   # one-SE rule (src/PLSEvaluator.cpp:284-308)
            if lo == 0:
                opt = 1
            else:
                opt = 0
                while opt < lo and mean[opt] > cutoff:
                    opt += 1
                if opt <= lo:
                    opt += 1
            pooled.append(fit_residuals(K, yg, seg[s], [seg[s + 1]], opt)[opt - 1][0])
            s += 2
        total += np.concatenate(pooled).std(ddof=1)  # :323-335
    return -total / R
