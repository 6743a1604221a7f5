"""Synthetic demand, restated from the reference's simulator without Spark / statsmodels.

``reference_weekly_demand`` follows group_apply/_resources/01-data-generator.py
(parameters :57-62, products :70-75, SKU ids :96-113, calendar and factors :135-181,
ARMA parameters :207-214, ``generate_arma`` :242-254, post-processing :295-306).
``statsmodels.tsa.arma_generate_sample(ar, ma, n, scale, burnin)`` is
``scipy.signal.lfilter(ma, ar, scale * standard_normal(n + burnin))[burnin:]``.
The reference's quirks are kept on purpose: the seed is reset inside every
``generate_arma`` call and the ARMA parameters are per *Product*, so all 10 SKUs of a
product share one series; ``scale=var`` passes a "variance" as the noise std-dev.

``daily_store_item_demand`` is the daily (store,item) workload of BASELINE.json's
configs: the same ingredients -- level, sqrt trend, covid ramp,
christmas / new-year factors, AR(1) noise, rounding -- plus a weekly pattern, for N
independent series, in NumPy (tests, CPU box) or on the GPU with torch (bench).
"""
from __future__ import annotations

import datetime as dt
import random
import string

import numpy as np

from . import design as D

END_DATE = np.datetime64("2021-07-19", "D")     # 01-data-generator.py:135 (a Monday)
PRODUCTS = (("Long Range Lidar", "LRL"), ("Short Range Lidar", "SRL"), ("Camera", "CAM"),
            ("Long Range Radar", "LRR"), ("Short Range Radar", "SRR"))      # :70-75
XMAS_FACTORS = {51: 0.85, 52: 0.8, 53: 0.8, 1: 1.1, 2: 1.15, 3: 1.1, 4: 1.05}  # :163-181 (week >= 52 -> 0.8)


def _sku_postfixes(n: int):
    """``id_sequence_generator`` (:100-113): ``random.seed(123)`` then n distinct 6-char ids.
    The reference returns ``list(set)`` (hash order); sorted here for determinism."""
    random.seed(123)
    chars = string.ascii_uppercase + string.digits
    res = set()
    while len(res) < n:
        res.add("".join(random.choice(chars) for _ in range(6)))
    return sorted(res)


def reference_calendar(ts_length_in_years: int = 3):
    """Weekly Mondays and the corona / christmas factors (:135-181)."""
    n = ts_length_in_years * 52 + 1
    days = END_DATE - np.arange(n - 1, -1, -1, dtype=np.int64) * np.timedelta64(7, "D")
    corona_breakpoint = np.datetime64("2020-03-01", "D")            # :59
    bp = int(np.searchsorted(days, corona_breakpoint, side="left"))  # :149-150
    help_list = [0] * (bp - 1) + list(range(0, n - bp + 1))         # :151
    assert len(help_list) == n
    frm, to = 20.0, 7.0                                             # :60-61
    mx = max(help_list)
    pct = [frm - ((frm - to) / mx) * k if k > 0 else 0 for k in help_list]     # :156
    corona_factor = np.array([1.0 if k == 0 else (100 - k) / 100 for k in pct])  # :157
    week = D.iso_week(days)
    xmas = np.array([XMAS_FACTORS.get(int(w), 1.0) for w in week])  # :163-181
    return days, np.array(help_list), corona_factor, xmas


def reference_weekly_demand(n_skus: int = 10):
    """The reference's ``part_level_demand`` table as a long pandas frame
    (Product, SKU, Date, Demand float32): 5 products x ``n_skus`` SKUs x 157 weeks."""
    import pandas as pd
    from scipy.signal import lfilter

    days, helper, corona_factor, xmas = reference_calendar()
    n = days.shape[0]
    n_prod = len(PRODUCTS)
    np.random.seed(123)                                                        # :207
    variance = np.abs(np.random.normal(100, 50, n_prod))                       # :209
    offset = np.maximum(np.abs(np.random.normal(10000, 5000, n_prod)), 4000)   # :210
    ar_len = np.random.choice(list(range(1, 4)), n_prod)                       # :211
    ar_par = [np.random.uniform(low=0.1, high=0.9, size=x) for x in ar_len]    # :212
    ma_len = np.random.choice(list(range(1, 4)), n_prod)                       # :213
    ma_par = [np.random.uniform(low=0.1, high=0.9, size=x) for x in ma_len]    # :214
    row_number = np.arange(n, dtype=np.float64)                                # :290
    frames = []
    date_objs = [dt.date.fromisoformat(str(d)) for d in days]
    for p, (product, prefix) in enumerate(PRODUCTS):
        np.random.seed(123)                                                    # :243 (inside every call)
        ar = np.r_[1, ar_par[p]]
        ma = np.r_[1, ma_par[p]]
        eta = variance[p] * np.random.standard_normal(n + 3000)                # scale=var, burnin=3000 (:246)
        y = lfilter(ma, ar, eta)[3000:] + offset[p]
        y = y * corona_factor                                                  # :299
        y = np.where(helper == 0, y + 100.0 * np.sqrt(row_number), y)          # :300-302
        y = np.round(y * xmas)                                                 # :303-304
        for post in _sku_postfixes(n_skus):
            frames.append(pd.DataFrame({"Product": product, "SKU": f"{prefix}_{post}", "Date": date_objs,
                                        "Demand": y.astype(np.float32)}))
    return pd.concat(frames, ignore_index=True)


# ---- daily (store,item) workload -------------------------------------------------------------
def _daily_factors(t_len: int, end=END_DATE):
    days = end - np.arange(t_len - 1, -1, -1, dtype=np.int64) * np.timedelta64(1, "D")
    week = D.iso_week(days)
    xmas = np.array([XMAS_FACTORS.get(int(w), 1.0) for w in week])
    k = np.maximum((days - np.datetime64("2020-03-01", "D")).astype(np.int64) + 1, 0)   # days since covid start
    mx = max(int(k.max()), 1)
    covid = np.where(k > 0, (100.0 - (20.0 - 13.0 * k / mx)) / 100.0, 1.0)   # ramp 0.80 -> 0.93 (:60-61,156-157)
    return days, D.weekday(days), xmas, covid, k == 0


def daily_store_item_demand(n: int, t_len: int, seed: int = 1234, nan_frac: float = 0.0, end=END_DATE,
                            out: np.ndarray | None = None):
    """NumPy generator.  Returns ``(y [n,t_len] float32, start_date)``; last date = ``end``."""
    from scipy.signal import lfilter

    rng = np.random.default_rng(seed)
    days, wd, xmas, covid, pre = _daily_factors(t_len, end)
    if out is None:
        out = np.empty((n, t_len), dtype=np.float32)
    sq = np.sqrt(np.arange(t_len, dtype=np.float64))
    block = 4096
    for i0 in range(0, n, block):
        m = min(block, n - i0)
        level = np.maximum(np.abs(rng.normal(10000, 5000, m)), 4000)      # :210
        sigma = np.abs(rng.normal(100, 50, m))                            # :209
        slope = rng.uniform(0, 100, m)                                    # :62
        phi = rng.uniform(0.1, 0.9, m)                                    # :212
        weekly = rng.uniform(0.8, 1.2, (m, 7))
        e = rng.standard_normal((m, t_len)) * sigma[:, None]
        noise = np.empty_like(e)
        for r in range(m):
            noise[r] = lfilter([1.0], [1.0, -phi[r]], e[r])
        y = (level[:, None] + noise) * covid[None, :]
        y = y + np.where(pre[None, :], slope[:, None] * sq[None, :], 0.0)  # :300-302
        y = y * xmas[None, :] * weekly[:, wd]
        out[i0:i0 + m] = np.round(y)                                       # :304
    if nan_frac > 0:
        mask = rng.random((n, t_len)) < nan_frac
        out[mask] = np.nan
    return out, days[0]


def daily_store_item_demand_torch(n: int, t_len: int, seed: int = 1234, nan_frac: float = 0.0, device="cuda",
                                  ld: int | None = None, end=END_DATE):
    """Same recipe generated on the GPU with torch ops (bench set-up plumbing, not timed).
    Returns ``(y [n, t_len] float32 view with row pitch ld, start_date)``."""
    import torch

    days, wd, xmas, covid, pre = _daily_factors(t_len, end)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = ld or ((t_len + 3) & ~3)
    full = torch.empty((n, ld), dtype=torch.float32, device=device)
    y = full[:, :t_len]
    f32 = dict(dtype=torch.float32, device=device)
    xmas_t = torch.tensor(xmas, **f32)
    covid_t = torch.tensor(covid, **f32)
    pre_t = torch.tensor(pre, device=device)
    sq = torch.sqrt(torch.arange(t_len, **f32))
    wd_t = torch.tensor(wd, device=device, dtype=torch.long)
    block = 1 << 20                       # one pass for the bench size: ~2.2k launches for the AR recursion
    for i0 in range(0, n, block):
        m = min(block, n - i0)
        level = torch.clamp(torch.abs(torch.randn(m, generator=g, **f32) * 5000 + 10000), min=4000)
        sigma = torch.abs(torch.randn(m, generator=g, **f32) * 50 + 100)
        slope = torch.rand(m, generator=g, **f32) * 100
        phi = torch.rand(m, generator=g, **f32) * 0.8 + 0.1
        weekly = torch.rand((m, 7), generator=g, **f32) * 0.4 + 0.8
        e = torch.randn((m, t_len), generator=g, **f32) * sigma[:, None]
        prev = torch.zeros(m, **f32)
        for t in range(t_len):                      # AR(1) recursion along time
            prev = phi * prev + e[:, t]
            e[:, t] = prev
        yb = (level[:, None] + e) * covid_t[None, :]
        yb = yb + torch.where(pre_t[None, :], slope[:, None] * sq[None, :], torch.zeros((), **f32))
        yb = yb * xmas_t[None, :] * weekly[:, wd_t]
        y[i0:i0 + m] = torch.round(yb)
        if nan_frac > 0:
            mask = torch.rand((m, t_len), generator=g, **f32) < nan_frac
            y[i0:i0 + m][mask] = float("nan")
    return y, days[0]
