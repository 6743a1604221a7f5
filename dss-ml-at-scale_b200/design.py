"""Shared-calendar design matrix (host side, float64, vectorised).

The reference builds its regressors per row inside a pandas UDF
(``add_exo_variables``, group_apply/02_Fine_Grained_Demand_Forecasting.py:343-358)
and hands them to SARIMAX as ``exog=`` (02:441-449).  Every series in a bucket
shares one calendar, so here the design is built ONCE per calendar from the grid
dates and uploaded; the kernels never materialise it per row.

Columns (P = 16):
  0 intercept | 1 linear trend | 2 sqrt trend (the trend the reference's generator
  injects, _resources/01-data-generator.py:301) | 3-8 day-of-week dummies (Tue..Sun)
  | 9-12 yearly Fourier pairs k=1,2 | 13 covid | 14 christmas | 15 new_year
  (the reference's three dummies verbatim, 02:351-356).
"""
from __future__ import annotations

import numpy as np

P = 16
FREQ_DAYS = {"D": 1, "W-MON": 7}
COVID_BREAKPOINT = np.datetime64("2020-03-01", "D")     # 02:351
FOURIER_EPOCH = np.datetime64("2000-01-01", "D")
COLUMN_NAMES = (
    "intercept", "lin", "sqrt",
    "dow1", "dow2", "dow3", "dow4", "dow5", "dow6",
    "sin1", "cos1", "sin2", "cos2",
    "covid", "christmas", "new_year",
)
DESIGNS = ("trend_season_exog", "exog_only")


def as_days(dates) -> np.ndarray:
    """Anything date-like -> numpy datetime64[D]."""
    arr = np.asarray(dates)
    if arr.dtype.kind == "M":
        return arr.astype("datetime64[D]")
    import pandas as pd

    return pd.to_datetime(arr).values.astype("datetime64[D]")


def calendar_grid(start, n: int, freq: str) -> np.ndarray:
    """``n`` grid dates from ``start`` (what ``asfreq(freq)`` yields, 02:423)."""
    step = FREQ_DAYS[freq]
    start = np.datetime64(start, "D")
    if freq == "W-MON" and weekday(np.array([start]))[0] != 0:
        raise ValueError("W-MON grid must start on a Monday")
    return start + np.arange(n, dtype=np.int64) * np.timedelta64(step, "D")


def weekday(days: np.ndarray) -> np.ndarray:
    """Monday = 0 ... Sunday = 6 (1970-01-01 was a Thursday)."""
    return (days.astype("datetime64[D]").astype(np.int64) + 3) % 7


def iso_week(days: np.ndarray) -> np.ndarray:
    """ISO-8601 week number == ``timestamp.dt.isocalendar().week`` (02:347)."""
    days = days.astype("datetime64[D]")
    thursday = days - weekday(days).astype("timedelta64[D]") + np.timedelta64(3, "D")
    year_start = thursday.astype("datetime64[Y]").astype("datetime64[D]")
    return ((thursday - year_start).astype(np.int64) // 7 + 1).astype(np.int64)


def exo_variables(days: np.ndarray) -> np.ndarray:
    """[n,3] float64: covid, christmas, new_year (02:351-356)."""
    days = as_days(days)
    week = iso_week(days)
    out = np.empty((days.shape[0], 3), dtype=np.float64)
    out[:, 0] = days >= COVID_BREAKPOINT
    out[:, 1] = (week >= 51) & (week <= 52)
    out[:, 2] = (week >= 1) & (week <= 4)
    return out


def design_matrix(days, t_fit: int, design: str = "trend_season_exog") -> np.ndarray:
    """X [len(days), P] float64 for the grid ``days``; rows [0,t_fit) are the fit window."""
    days = as_days(days)
    n = days.shape[0]
    X = np.zeros((n, P), dtype=np.float64)
    exo = exo_variables(days)
    if design == "exog_only":
        X[:, 0:3] = exo            # SARIMAX(0,0,0) + exog, trend=None (02:441-449)
        return X
    if design != "trend_season_exog":
        raise ValueError(f"unknown design {design!r}; expected one of {DESIGNS}")
    t = np.arange(n, dtype=np.float64)
    tf = float(t_fit)
    X[:, 0] = 1.0
    X[:, 1] = (t - (tf - 1.0) / 2.0) / tf
    X[:, 2] = np.sqrt(t / tf)
    wd = weekday(days)
    for k in range(1, 7):
        X[:, 2 + k] = wd == k
    tau = (days - FOURIER_EPOCH).astype(np.float64) / 365.25
    X[:, 9] = np.sin(2.0 * np.pi * tau)
    X[:, 10] = np.cos(2.0 * np.pi * tau)
    X[:, 11] = np.sin(4.0 * np.pi * tau)
    X[:, 12] = np.cos(4.0 * np.pi * tau)
    X[:, 13:16] = exo
    return X


def design_has_constant(design: str) -> bool:
    """True iff column 0 of the design is identically 1 (enables per-series centring)."""
    return design == "trend_season_exog"
