"""CPU oracle for the many-models forecast path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``dss-ml-at-scale_b200``) never does; it fails loudly without its CUDA library.

PARITY STATUS: **parity unpinned** for the per-series model arithmetic.
The reference's per-group model is statsmodels ``SARIMAX`` tuned by ``hyperopt``
(``group_apply/02_Fine_Grained_Demand_Forecasting.py:435-481``); neither
library (nor pyspark, nor a JVM) exists in this image, no versions are pinned
anywhere in the reference, and the reference's ``fmin`` call is unseeded
(``02:467-469``) so its output is not reproducible even against itself.  What
IS pinned against the reference's own code (executed from ``/root/reference``
by ``tests/golden/make_reference_fixtures.py``): the calendar dummies of
``add_exo_variables`` (``02:343-358``), ``split_train_score_data``
(``02:372-380``) and the synthetic-demand calendar (``01-data-generator.py:
135-181``).  The model itself is the p=d=q=0 corner of the reference's search
space (regression of ``y`` on a design matrix; SARIMAX(0,0,0)+exog MLE == OLS)
with the richer design that ``BASELINE.json.north_star`` names.

Everything here is float64 and deliberately simple (loops over dates use
``datetime``), i.e. an independent restatement of the spec that the CUDA path
is compared against.

Model spec (shared with the engine)
--------------------------------------------------------
For one group (one series):
 1. sort by Date, re-index onto the regular grid min..max at ``freq``; gaps
    become NaN (``02:422-423``).
 2. ``mode="holdout"`` (reference semantics): the last ``horizon`` grid rows
    are held out (``02:372-380``), the model is fit on the observed rows of
    the first ``T - horizon`` rows and evaluated on *every* grid row
    (``02:484-488``): fitted values for the train dates, forecast for the
    hold-out dates.  ``mode="future"``: fit on all ``T`` rows, forecast the
    ``horizon`` rows after the end.
 3. design row x_t (P = 16), t = grid index, d_t = grid date, Tf = #fit rows:
      [1, (t-(Tf-1)/2)/Tf, sqrt(t/Tf), dow(d_t)==1..6 (Mon = baseline),
       sin/cos(2*pi*k*days_since_2000(d_t)/365.25) k=1,2,
       covid(d_t), christmas(d_t), new_year(d_t)]      (``02:351-356``)
 4. the shared calendar Gram  G = X_fit^T X_fit  is Cholesky-factored in column
    order; a column whose pivot is <= CAL_TOL * G_jj is aliased and dropped
    (like R's lm).  W = L^-T (zero rows/cols for dropped columns) whitens the
    design: A = X W, A_fit^T A_fit = I on the retained columns.
 5. per series, in the whitened basis: b = sum_obs a_t (y_t), G_i = sum_obs a_t
    a_t^T; in-order Cholesky of G_i, a column whose pivot is <= PIVOT_TOL *
    G_i[j,j] is dropped (gamma_j = 0).  For a fully observed series G_i = I
    and gamma = b.  Non-finite y_t == missing.
 6. prediction  yhat_t = a_t . gamma  for the requested rows.
 status: 0 ok, 1 no observed fit rows (outputs NaN), 2 ok but some whitened
 column was dropped for this series' mask.
"""
from __future__ import annotations

import datetime as _dt
import math

import numpy as np

P = 16                    # design columns (one MMA tile edge)
CAL_TOL = 1e-10           # aliasing threshold on the float64 calendar Gram
PIVOT_TOL = 1e-3          # per-series relative pivot threshold (== MMF_PIVOT_TOL in include/mmf.h)
COVID_BREAKPOINT = _dt.date(2020, 3, 1)   # 02:351
FOURIER_EPOCH = _dt.date(2000, 1, 1)
COLUMN_NAMES = (
    "intercept", "lin", "sqrt",
    "dow1", "dow2", "dow3", "dow4", "dow5", "dow6",
    "sin1", "cos1", "sin2", "cos2",
    "covid", "christmas", "new_year",
)
FREQ_DAYS = {"D": 1, "W-MON": 7}


# ----------------------------------------------------------------------------
# calendar  (02:343-358, 02:422-423)
# ----------------------------------------------------------------------------
def to_date(d) -> _dt.date:
    if isinstance(d, _dt.datetime):
        return d.date()
    if isinstance(d, _dt.date):
        return d
    # numpy datetime64 / pandas Timestamp
    return _dt.date.fromisoformat(str(np.datetime64(d, "D")))


def calendar_grid(start, n: int, freq: str) -> list:
    """``n`` dates from ``start`` on the regular grid of ``freq`` (asfreq, 02:423)."""
    step = FREQ_DAYS[freq]
    start = to_date(start)
    if freq == "W-MON" and start.weekday() != 0:
        raise ValueError("W-MON grid must start on a Monday")
    return [start + _dt.timedelta(days=step * i) for i in range(n)]


def exo_variables(dates) -> np.ndarray:
    """covid / christmas / new_year 0-1 regressors, verbatim logic of
    ``add_exo_variables`` (02:345-356): covid = ts >= 2020-03-01,
    christmas = ISO week in [51,52], new_year = ISO week in [1,4]."""
    out = np.zeros((len(dates), 3))
    for i, d in enumerate(dates):
        d = to_date(d)
        week = d.isocalendar()[1]
        out[i, 0] = 1.0 if d >= COVID_BREAKPOINT else 0.0
        out[i, 1] = 1.0 if 51 <= week <= 52 else 0.0
        out[i, 2] = 1.0 if 1 <= week <= 4 else 0.0
    return out


def design_matrix(dates, t_fit: int, design: str = "trend_season_exog") -> np.ndarray:
    """X[len(dates), P] float64 (spec item 3).  ``design="exog_only"`` is the
    reference's literal regressor set [covid, christmas, new_year] with no
    intercept (SARIMAX trend=None, 02:441-449), zero-padded to P columns."""
    n = len(dates)
    X = np.zeros((n, P))
    exo = exo_variables(dates)
    if design == "exog_only":
        X[:, 0:3] = exo
        return X
    if design != "trend_season_exog":
        raise ValueError(design)
    tf = float(t_fit)
    for t, d in enumerate(dates):
        d = to_date(d)
        X[t, 0] = 1.0
        X[t, 1] = (t - (tf - 1.0) / 2.0) / tf
        X[t, 2] = math.sqrt(t / tf)
        wd = d.weekday()
        if wd >= 1:
            X[t, 2 + wd] = 1.0
        tau = (d - FOURIER_EPOCH).days / 365.25
        X[t, 9] = math.sin(2.0 * math.pi * tau)
        X[t, 10] = math.cos(2.0 * math.pi * tau)
        X[t, 11] = math.sin(4.0 * math.pi * tau)
        X[t, 12] = math.cos(4.0 * math.pi * tau)
    X[:, 13:16] = exo
    return X


def design_has_constant(design: str) -> bool:
    return design == "trend_season_exog"


# ----------------------------------------------------------------------------
# split (02:372-380)
# ----------------------------------------------------------------------------
def split_train_score_data(data, forecast_horizon: int):
    """First ``len-horizon`` rows train, last ``horizon`` rows score."""
    n = len(data)
    is_history = np.array([True] * (n - forecast_horizon) + [False] * forecast_horizon)
    if hasattr(data, "iloc"):
        return data.iloc[is_history], data.iloc[~is_history]
    return data[is_history], data[~is_history]


# ----------------------------------------------------------------------------
# whitening of the shared calendar design (spec item 4)
# ----------------------------------------------------------------------------
def whiten(X_fit: np.ndarray):
    """Returns (W [P,P], kept [P] bool) with A = X @ W orthonormal on the fit
    rows for kept columns and exactly zero for aliased ones."""
    G = X_fit.T @ X_fit
    p = G.shape[0]
    L = np.zeros((p, p))
    kept = np.zeros(p, dtype=bool)
    for j in range(p):
        d = G[j, j] - np.dot(L[j, :j], L[j, :j])
        if G[j, j] <= 0.0 or d <= CAL_TOL * G[j, j]:
            continue                      # aliased: row/column j of L stays zero
        kept[j] = True
        L[j, j] = math.sqrt(d)
        for i in range(j + 1, p):
            L[i, j] = (G[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    # W = L^-T on the kept set
    idx = np.flatnonzero(kept)
    W = np.zeros((p, p))
    if idx.size:
        Lk = L[np.ix_(idx, idx)]
        W[np.ix_(idx, idx)] = np.linalg.inv(Lk).T
    return W, kept


# ----------------------------------------------------------------------------
# per-series solve in the whitened basis (spec item 5)
# ----------------------------------------------------------------------------
def solve_series(y_fit: np.ndarray, A_fit: np.ndarray):
    """Returns (gamma[P], status, min_pivot_ratio)."""
    p = A_fit.shape[1]
    obs = np.isfinite(y_fit)
    if not obs.any():
        return np.full(p, np.nan), 1, 0.0
    Ao = A_fit[obs]
    b = Ao.T @ y_fit[obs]
    G = Ao.T @ Ao
    L = np.zeros((p, p))
    kept = np.zeros(p, dtype=bool)
    min_ratio = 1.0
    dropped = False
    for j in range(p):
        gjj = G[j, j]
        if gjj <= 0.0:
            continue                      # globally aliased (zero) column
        d = gjj - np.dot(L[j, :j], L[j, :j])
        if d <= PIVOT_TOL * gjj:
            dropped = True
            continue
        min_ratio = min(min_ratio, d / gjj)
        kept[j] = True
        L[j, j] = math.sqrt(d)
        for i in range(j + 1, p):
            L[i, j] = (G[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    idx = np.flatnonzero(kept)
    gamma = np.zeros(p)
    if idx.size:
        Lk = L[np.ix_(idx, idx)]
        z = np.linalg.solve(Lk, b[idx])
        gamma[idx] = np.linalg.solve(Lk.T, z)
    return gamma, (2 if dropped else 0), min_ratio


def fit_forecast_packed(y: np.ndarray, X: np.ndarray, t_fit: int, pred_start: int,
                        n_pred: int, return_gamma: bool = False):
    """Vectorised float64 oracle on packed series.

    y [N, >=t_fit] (NaN = missing; columns beyond t_fit ignored),
    X [>=pred_start+n_pred, P] raw design rows on the shared calendar.
    Returns pred [N, n_pred] float64, status [N] int32 (and gamma [N,P], min
    pivot ratio [N] when asked).
    """
    y = np.asarray(y, dtype=np.float64)[:, :t_fit]
    X = np.asarray(X, dtype=np.float64)
    W, _ = whiten(X[:t_fit])
    A = X @ W
    A_fit, A_pred = A[:t_fit], A[pred_start:pred_start + n_pred]
    n = y.shape[0]
    gamma = np.zeros((n, A.shape[1]))
    status = np.zeros(n, dtype=np.int32)
    ratio = np.ones(n)
    full = np.isfinite(y).all(axis=1)
    if full.any():
        gamma[full] = y[full] @ A_fit          # G_i = I  =>  gamma = b
    for i in np.flatnonzero(~full):
        gamma[i], status[i], ratio[i] = solve_series(y[i], A_fit)
    pred = gamma @ A_pred.T
    if return_gamma:
        return pred, status, gamma, ratio
    return pred, status


def lstsq_reference(y_fit: np.ndarray, X_fit: np.ndarray, X_pred: np.ndarray) -> np.ndarray:
    """Independent route (numpy.linalg.lstsq, minimum-norm) used only to check
    the whitened-Cholesky route on full-rank problems."""
    obs = np.isfinite(y_fit)
    beta, *_ = np.linalg.lstsq(X_fit[obs], y_fit[obs], rcond=None)
    return X_pred @ beta


def beta_from_gamma(gamma: np.ndarray, X_fit: np.ndarray) -> np.ndarray:
    """Coefficients on the raw design columns: beta = W gamma."""
    W, _ = whiten(X_fit)
    return gamma @ W.T


# ----------------------------------------------------------------------------
# the per-group UDF, same skeleton as build_tune_and_score_model (02:417-494)
# ----------------------------------------------------------------------------
def build_tune_and_score_model(sku_pdf, *, keys=("Product", "SKU"), date_col="Date",
                               value_col="Demand", freq="W-MON", horizon=40,
                               mode="holdout", design="trend_season_exog", null_keys_on_gaps=False):
    """One group's rows in -> one group's rows out (``tuning_schema``, 02:498-506):
    keys..., Date, Demand, Demand_Fitted for every grid date.  ``null_keys_on_gaps=True`` reproduces a detail of the
    reference's output assembly (02:490): the key columns are read from the re-indexed frame, so the rows asfreq()
    inserted for missing dates carry NaN keys."""
    import pandas as pd

    # 02:422-423  sort + regular grid (NaN for gaps)
    pdf = sku_pdf.sort_values(date_col)
    dates_in = [to_date(d) for d in pdf[date_col].tolist()]
    step = FREQ_DAYS[freq]
    d0, d1 = dates_in[0], dates_in[-1]
    T = (d1 - d0).days // step + 1
    grid = calendar_grid(d0, T, freq)
    y = np.full(T, np.nan)
    present = np.zeros(T, dtype=bool)
    vals = pdf[value_col].to_numpy(dtype=np.float64)
    for d, v in zip(dates_in, vals):
        off = (d - d0).days
        if off % step:
            continue                                   # off-grid rows vanish under asfreq
        y[off // step] = v
        present[off // step] = True
    key_vals = [pdf[k].iloc[0] for k in keys]          # 02:428-429

    if mode == "holdout":                              # 02:430, 372-380
        t_fit = T - horizon
        if t_fit <= 0:
            raise ValueError("series shorter than the forecast horizon")
        dates_all, pred_start, n_pred = grid, 0, T
    elif mode == "future":
        t_fit = T
        dates_all = calendar_grid(d0, T + horizon, freq)
        pred_start, n_pred = T, horizon
    else:
        raise ValueError(mode)

    X = design_matrix(dates_all, t_fit, design)
    pred, status = fit_forecast_packed(y[None, :], X, t_fit, pred_start, n_pred)
    out_dates = dates_all[pred_start:pred_start + n_pred]
    demand = y if mode == "holdout" else np.full(horizon, np.nan)
    out = {k: [v] * n_pred for k, v in zip(keys, key_vals)}
    if null_keys_on_gaps and mode == "holdout":
        out = {k: [v if p else None for v, p in zip(col, present)] for k, col in out.items()}
    out[date_col] = out_dates
    out[value_col] = demand.astype(np.float32)
    out[value_col + "_Fitted"] = pred[0].astype(np.float32)   # 02:490-494
    return pd.DataFrame(out)


def fanout_apply(pdf, func, keys):
    """Stand-in for ``groupBy(keys).applyInPandas(func, schema)`` (02:523-528):
    one Python call per group, results concatenated."""
    import pandas as pd

    parts = [func(g.copy()) for _, g in pdf.groupby(list(keys), sort=True)]
    return pd.concat(parts, ignore_index=True)


# ----------------------------------------------------------------------------
# per-series model selection (the engine's analogue of the reference's hyperopt loop, 02:435-488)
# ----------------------------------------------------------------------------
def select_forecast_packed(y: np.ndarray, X: np.ndarray, t_fit: int, n_hold: int, candidates, pred_start: int,
                           n_pred: int):
    """Candidates = nested models on the first ``m`` whitened columns (``m`` in ``candidates``, ascending).
    Spec: fit the full whitened model on the observed rows of [0,t_fit); candidate m keeps
    the first m whitened coefficients (for a gap-free series that IS the least-squares fit of the sub-model, the
    basis being orthonormal on the calendar); score = MSE over the observed held-out rows [t_fit,t_fit+n_hold);
    the first minimum wins (no observed held-out row: the last candidate); predict with the winner.
    Returns pred [N,n_pred], choice [N] (0 for empty series), mse [N], status [N]."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    _, status, gamma, _ = fit_forecast_packed(y[:, :t_fit], X, t_fit, 0, 1, return_gamma=True)
    W, _ = whiten(X[:t_fit])
    A = X @ W
    A_hold = A[t_fit:t_fit + n_hold]
    n = y.shape[0]
    pred = np.full((n, n_pred), np.nan)
    choice = np.zeros(n, dtype=np.int32)
    mse = np.full(n, np.nan)
    for i in range(n):
        if status[i] == 1:
            continue
        yh = y[i, t_fit:t_fit + n_hold]
        obs = np.isfinite(yh)
        best_m, best = candidates[-1], np.nan
        if obs.any():
            scores = []
            for m in candidates:
                e = yh[obs] - A_hold[obs][:, :m] @ gamma[i, :m]
                scores.append(float(np.mean(e * e)))
            k = int(np.argmin(scores))                   # first minimum
            best_m, best = candidates[k], scores[k]
        choice[i], mse[i] = best_m, best
        pred[i] = A[pred_start:pred_start + n_pred, :best_m] @ gamma[i, :best_m]
    return pred, choice, mse, status


# ----------------------------------------------------------------------------
# C restatement (oracle/mmf_oracle_c.c, pthreads): the fair multi-core CPU baseline
# ----------------------------------------------------------------------------
def fit_forecast_packed_c(y: np.ndarray, X: np.ndarray, t_fit: int, pred_start: int, n_pred: int, n_threads: int = 0,
                          return_threads: bool = False, out=None, status=None, prepared=None):
    """Same contract as :func:`fit_forecast_packed`, computed by ``libmmf_oracle.so`` (float64 accumulation, one
    pass per series, all host cores).  ``y`` float32 [N, ld].  ``out`` / ``status`` / ``prepared`` (the whitened
    design from a previous call's ``prepared=`` dict) let a timed loop reuse its buffers instead of paying
    page faults and the calendar whitening on every pass."""
    import ctypes as C
    import os

    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmmf_oracle.so"))
    y = np.ascontiguousarray(y, dtype=np.float32)
    if prepared is not None and "A" in prepared:
        A, kept32 = prepared["A"], prepared["kept32"]
    else:
        W, kept = whiten(np.asarray(X, dtype=np.float64)[:t_fit])
        A = np.ascontiguousarray(np.asarray(X, dtype=np.float64) @ W)
        kept32 = np.ascontiguousarray(kept.astype(np.int32))
        if prepared is not None:
            prepared["A"], prepared["kept32"] = A, kept32
    n = y.shape[0]
    out = np.empty((n, n_pred), dtype=np.float64) if out is None else out
    status = np.empty(n, dtype=np.int32) if status is None else status
    lib.mmf_oracle_fit_forecast.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
    lib.mmf_oracle_fit_forecast.restype = C.c_int
    used = lib.mmf_oracle_fit_forecast(y.ctypes.data, n, y.strides[0] // 4, t_fit, A.ctypes.data, kept32.ctypes.data,
                                       pred_start, n_pred, out.ctypes.data, status.ctypes.data, n_threads)
    if return_threads:
        return out, status, int(used)
    return out, status


def numa_local_sample(y: np.ndarray, n_threads: int = 0) -> np.ndarray:
    """Copy of the float32 sample ``y`` [n, ld] whose pages were first touched by the pinned worker threads that
    :func:`fit_forecast_packed_c` later assigns to the same rows (``mmf_oracle_alloc_local``): the timed CPU baseline
    then reads socket-local memory on every box instead of wherever the calling thread happened to allocate."""
    import ctypes as C
    import os

    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmmf_oracle.so"))
    lib.mmf_oracle_alloc_local.argtypes = [C.c_int64, C.c_int64, C.c_int32]
    lib.mmf_oracle_alloc_local.restype = C.c_void_p
    lib.mmf_oracle_free_local.argtypes = [C.c_void_p]
    y = np.asarray(y, dtype=np.float32)
    n, ld = y.shape
    addr = lib.mmf_oracle_alloc_local(n, ld, n_threads)
    if not addr:
        raise MemoryError("mmf_oracle_alloc_local failed")
    buf = (C.c_float * (n * ld)).from_address(addr)
    arr = np.frombuffer(buf, dtype=np.float32).reshape(n, ld)
    import weakref
    weakref.finalize(buf, lib.mmf_oracle_free_local, addr)
    arr[...] = y
    return arr
