"""CPU: the oracle against the reference-executed fixtures, the frozen golden vectors,
an independent lstsq route and closed-form known-answer tests."""
import datetime as dt

import numpy as np
import pytest

import mmf
from oracle import mmf_oracle as O


# ---- pinned against the reference's own code (tests/golden/make_reference_fixtures.py) -----
@pytest.mark.parametrize("name", ["weekly", "daily"])
def test_exo_variables_match_reference_code(reference_fixtures, name):
    days = reference_fixtures[f"exo_{name}_days"].astype("datetime64[D]")
    want = reference_fixtures[f"exo_{name}"]
    got_oracle = O.exo_variables([dt.date.fromisoformat(str(d)) for d in days])
    got_pkg = mmf.design.exo_variables(days)
    assert np.array_equal(got_oracle, want)            # 0/1 floats: bit exact
    assert np.array_equal(got_pkg, want)


def test_exo_counts_on_reference_calendar(reference_fixtures):
    # covid 73 / christmas 6 / new_year 12 ones on the 157-week calendar
    assert reference_fixtures["exo_weekly"].sum(axis=0).tolist() == [73.0, 6.0, 12.0]


def test_split_matches_reference_code(reference_fixtures):
    for n, h, n_train, n_score, first_score, last_train in reference_fixtures["split_cases"]:
        data = np.arange(n)
        for split in (O.split_train_score_data, mmf.split_train_score_data):
            tr, sc = split(data, int(h))
            assert (len(tr), len(sc)) == (n_train, n_score)
            assert sc[0] == first_score and tr[-1] == last_train


def test_generator_calendar_matches_reference_code(reference_fixtures):
    days, helper, corona, xmas = mmf.synth.reference_calendar()
    assert np.array_equal(days.astype(np.int64), reference_fixtures["gen_days"])
    assert np.array_equal(helper, reference_fixtures["gen_helper"])
    assert np.allclose(corona, reference_fixtures["gen_corona_factor"], rtol=0, atol=1e-15)
    assert np.array_equal(xmas, reference_fixtures["gen_factor_xmas"])
    assert np.array_equal(mmf.design.iso_week(days), reference_fixtures["gen_week"])


# ---- frozen oracle outputs -------------------------------------------------------------------
def test_oracle_reproduces_golden(oracle_golden):
    g = oracle_golden
    y = g["ref_weekly_y"]
    T = y.shape[1]
    grid = O.calendar_grid(g["ref_weekly_start"][0].astype("datetime64[D]"), T, "W-MON")
    pred, status = O.fit_forecast_packed(y, O.design_matrix(grid, T - 40), T - 40, 0, T)
    assert np.allclose(pred, g["ref_weekly_fitted"], rtol=0, atol=1e-7)
    assert np.array_equal(status, g["ref_weekly_status"])

    y = g["daily365_y"]
    grid = O.calendar_grid(g["daily365_start"][0].astype("datetime64[D]"), 365 + 28, "D")
    pred, status = O.fit_forecast_packed(y, O.design_matrix(grid, 365), 365, 365, 28)
    assert np.allclose(pred, g["daily365_pred"], rtol=0, atol=1e-7)
    assert np.array_equal(status, g["daily365_status"])


def test_package_design_equals_oracle_design():
    for freq, n, t_fit, start in (("D", 1123, 1095, "2018-07-21"), ("W-MON", 157, 117, "2018-07-23"),
                                  ("D", 400, 365, "2019-12-20")):
        days = mmf.design.calendar_grid(start, n, freq)
        got = mmf.design.design_matrix(days, t_fit)
        want = O.design_matrix(O.calendar_grid(dt.date.fromisoformat(start), n, freq), t_fit)
        assert np.abs(got - want).max() < 1e-12
        got = mmf.design.design_matrix(days, t_fit, "exog_only")
        want = O.design_matrix(O.calendar_grid(dt.date.fromisoformat(start), n, freq), t_fit, "exog_only")
        assert np.array_equal(got, want)


# ---- the whitened-Cholesky route against an independent solver ----------------------------
def test_whitened_route_equals_lstsq(oracle_golden):
    y = oracle_golden["daily1095_y"].astype(np.float64)
    grid = O.calendar_grid(oracle_golden["daily1095_start"][0].astype("datetime64[D]"), 1095 + 28, "D")
    X = O.design_matrix(grid, 1095)
    pred, status = O.fit_forecast_packed(y, X, 1095, 1095, 28)
    assert (status == 0).all()
    for i in range(0, y.shape[0], 5):
        ref = O.lstsq_reference(y[i], X[:1095], X[1095:])
        assert np.abs(pred[i] - ref).max() < 1e-6
    # with gaps
    rng = np.random.default_rng(0)
    yi = y[3].copy()
    yi[rng.random(1095) < 0.1] = np.nan
    p1, st = O.fit_forecast_packed(yi[None], X, 1095, 0, 1123)
    assert st[0] == 0
    assert np.abs(p1[0] - O.lstsq_reference(yi, X[:1095], X)).max() < 1e-6


def test_whiten_is_orthonormal_and_drops_aliased():
    grid = O.calendar_grid(dt.date(2020, 7, 20), 365 + 28, "D")     # covid == 1 everywhere
    X = O.design_matrix(grid, 365)
    W, kept = O.whiten(X[:365])
    assert not kept[13] and kept.sum() == 15
    A = X[:365] @ W
    G = A.T @ A
    assert np.abs(G - np.diag(kept.astype(float))).max() < 1e-9
    # weekly grid: every date is a Monday -> the six day-of-week dummies vanish
    Xw = O.design_matrix(O.calendar_grid(dt.date(2018, 7, 23), 157, "W-MON"), 117)
    _, keptw = O.whiten(Xw[:117])
    assert not keptw[3:9].any() and keptw[[0, 1, 2, 9, 10, 11, 12, 13, 14, 15]].all()


# ---- closed-form known answers ------------------------------------------------------------------
def _daily_design(T=200, H=28, start=dt.date(2019, 1, 1)):
    grid = O.calendar_grid(start, T + H, "D")
    return O.design_matrix(grid, T)


def test_kat_pure_line_forecast_is_exact():
    X = _daily_design()
    t = np.arange(228, dtype=np.float64)
    y = 500.0 + 3.0 * t
    pred, st = O.fit_forecast_packed(y[None, :200], X, 200, 200, 28)
    assert st[0] == 0 and np.abs(pred[0] - y[200:]).max() < 1e-6


def test_kat_constant_and_weekday_pattern():
    X = _daily_design()
    pred, _ = O.fit_forecast_packed(np.full((1, 200), 42.0), X, 200, 200, 28)
    assert np.abs(pred - 42.0).max() < 1e-7
    grid = O.calendar_grid(dt.date(2019, 1, 1), 228, "D")
    pattern = np.array([10.0, 20, 30, 40, 50, 60, 70])
    y = np.array([pattern[d.weekday()] for d in grid])
    pred, _ = O.fit_forecast_packed(y[None, :200], X, 200, 200, 28)
    assert np.abs(pred[0] - y[200:]).max() < 1e-6


def test_kat_gaps_empty_and_rank_deficient():
    X = _daily_design()
    t = np.arange(228, dtype=np.float64)
    y = 100.0 + 2.0 * t
    yg = y[:200].copy()
    yg[[3, 50, 51, 52, 199]] = np.nan
    pred, st = O.fit_forecast_packed(yg[None], X, 200, 200, 28)
    assert st[0] == 0 and np.abs(pred[0] - y[200:]).max() < 1e-6
    # all missing
    pred, st = O.fit_forecast_packed(np.full((1, 200), np.nan), X, 200, 200, 28)
    assert st[0] == 1 and np.isnan(pred).all()
    # a single observation: everything but the intercept is aliased, forecast == that value
    y1 = np.full(200, np.nan)
    y1[17] = 7.5
    pred, st = O.fit_forecast_packed(y1[None], X, 200, 200, 28)
    assert st[0] == 2 and np.abs(pred - 7.5).max() < 1e-9
    # Inf counts as missing
    yi = y[:200].copy()
    yi[10] = np.inf
    pred, st = O.fit_forecast_packed(yi[None], X, 200, 200, 28)
    assert np.abs(pred[0] - y[200:]).max() < 1e-6


def test_exog_only_design_is_ols_on_the_three_dummies():
    # the literal p=d=q=0 corner of the reference's search space: y ~ covid + christmas + new_year, no constant
    grid = O.calendar_grid(dt.date(2018, 7, 23), 157, "W-MON")
    X = O.design_matrix(grid, 117, "exog_only")
    rng = np.random.default_rng(1)
    y = 3.0 * X[:, 0] - 2.0 * X[:, 1] + 5.0 * X[:, 2] + rng.normal(0, 0.1, 157)
    pred, st = O.fit_forecast_packed(y[None, :117], X, 117, 0, 157)
    beta, *_ = np.linalg.lstsq(X[:117, :3], y[:117], rcond=None)
    assert np.abs(pred[0] - X[:, :3] @ beta).max() < 1e-9


# ---- the per-group UDF skeleton (02:417-494) ---------------------------------------------------
def test_udf_contract_on_reference_data():
    df = mmf.synth.reference_weekly_demand(n_skus=2)
    assert len(df) == 5 * 2 * 157
    one = df[df["SKU"] == df["SKU"].iloc[0]].sample(frac=1.0, random_state=0)    # shuffled rows, 02:422 sorts
    out = O.build_tune_and_score_model(one)
    assert list(out.columns) == ["Product", "SKU", "Date", "Demand", "Demand_Fitted"]
    assert len(out) == 157 and out["Date"].is_monotonic_increasing
    assert out["Demand"].dtype == np.float32 and out["Demand_Fitted"].dtype == np.float32
    # a gap in the input becomes a NaN Demand row on the regular grid (asfreq, 02:423)
    holed = one[one["Date"] != sorted(one["Date"])[10]]
    out2 = O.build_tune_and_score_model(holed)
    assert len(out2) == 157 and np.isnan(out2["Demand"].iloc[10]) and np.isfinite(out2["Demand_Fitted"].iloc[10])
    allg = O.fanout_apply(df, O.build_tune_and_score_model, ("Product", "SKU"))
    assert len(allg) == len(df)
    # future mode
    fut = O.build_tune_and_score_model(one, mode="future", horizon=8)
    assert len(fut) == 8 and fut["Demand"].isna().all()
    assert fut["Date"].iloc[0] == dt.date(2021, 7, 26)


def test_select_truncation_equals_refitting_nested_models():
    """For a gap-free series the candidate 'first m whitened columns' equals the least-squares refit on the first m
    raw design columns -- the property the device-side selection relies on."""
    grid = O.calendar_grid(dt.date(2019, 1, 1), 300, "D")
    X = O.design_matrix(grid, 272)
    rng = np.random.default_rng(4)
    t = np.arange(300.0)
    y = 1000 + 2 * t + 50 * np.sin(2 * np.pi * t / 7) + rng.normal(0, 5, 300)
    pred, choice, mse, st = O.select_forecast_packed(y[None], X, 272, 28, (1, 3, 9, 13, 16), 0, 300)
    assert st[0] == 0 and choice[0] in (9, 13, 16)
    m = int(choice[0])
    ref = O.lstsq_reference(y[:272], X[:272, :m], X[:, :m])
    assert np.abs(pred[0] - ref).max() < 1e-6
    # a pure constant series: every candidate scores ~0, the first (smallest) one wins
    pred, choice, mse, _ = O.select_forecast_packed(np.full((1, 300), 7.0), X, 272, 28, (1, 3, 9, 13, 16), 0, 300)
    assert choice[0] == 1 and np.abs(pred - 7.0).max() < 1e-9


def test_bridge_to_reference_model_family_statsmodels():
    """The only statement that ties the oracle's arithmetic to the reference's model class: SARIMAX with
    order=(0,0,0), trend=None and exog=X (a corner of the reference's search space, 02:441-449, 461-465) is Gaussian
    regression on X, so its MLE forecast equals the oracle's exog_only design.  statsmodels is absent from this
    image (the test then skips); it runs wherever the reference's own dependencies exist."""
    sm = pytest.importorskip("statsmodels.tsa.statespace.sarimax")
    import pandas as pd
    df = mmf.synth.reference_weekly_demand(n_skus=1)
    one = df[df["SKU"] == df["SKU"].iloc[0]].sort_values("Date")
    ts = one.set_index(pd.DatetimeIndex(pd.to_datetime(one["Date"]), freq="W-MON"))
    exo = pd.DataFrame(O.exo_variables(list(one["Date"])), index=ts.index, columns=["covid", "christmas", "new_year"])
    train, score = ts.iloc[:117], ts.iloc[117:]
    fitted = sm.SARIMAX(train["Demand"], exog=exo.iloc[:117], order=(0, 0, 0), seasonal_order=(0, 0, 0, 0),
                        enforce_stationarity=False, enforce_invertibility=False).fit(disp=False)
    fc = fitted.predict(start=score.index.min(), end=score.index.max(), exog=exo.iloc[117:])
    ours = O.build_tune_and_score_model(one, design="exog_only")
    assert np.abs(fc.to_numpy() - ours["Demand_Fitted"].to_numpy()[117:]).max() <= 1e-3 * float(one["Demand"].max())


def test_reference_objective_optimum_is_the_oracle_fit():
    """The reference maximises the Gaussian likelihood of SARIMAX(p,d,q)+exog with Nelder-Mead
    (``model.fit(disp=False, method='nm')``, 02:441-450, 472-481).  At the (0,0,0) corner of its search space that
    likelihood is  -n/2 log(2 pi s2) - |y - X b|^2 / (2 s2)  in (b, s2): two independent third-party solvers
    (SciPy's Nelder-Mead on exactly that objective, started from the oracle's answer and from a perturbed point, and
    scikit-learn's least squares) must agree with the oracle's exog_only fit -- the closed form is the optimum the
    reference's optimiser is searching for."""
    from scipy.optimize import minimize
    from sklearn.linear_model import LinearRegression
    df = mmf.synth.reference_weekly_demand(n_skus=1)
    one = df[df["SKU"] == df["SKU"].iloc[0]].sort_values("Date")
    y = one["Demand"].to_numpy(dtype=np.float64)
    X = O.exo_variables(list(one["Date"])).astype(np.float64)
    ours = O.build_tune_and_score_model(one, design="exog_only")["Demand_Fitted"].to_numpy(dtype=np.float64)
    ytr, Xtr = y[:117], X[:117]

    def nll(theta):
        b, log_s2 = theta[:3], theta[3]
        r = ytr - Xtr @ b
        return 0.5 * (117 * (np.log(2 * np.pi) + log_s2) + r @ r / np.exp(log_s2))

    skl = LinearRegression(fit_intercept=False).fit(Xtr, ytr)
    assert np.abs(X @ skl.coef_ - ours).max() <= 1e-6 * np.abs(y).max()
    b0 = skl.coef_
    s2 = np.mean((ytr - Xtr @ b0) ** 2)
    f_star = nll(np.r_[b0, np.log(s2)])
    for start in (np.r_[b0, np.log(s2)], np.r_[b0 * 1.05 + 10.0, np.log(s2) + 0.3]):
        res = minimize(nll, start, method="Nelder-Mead", options={"xatol": 1e-9, "fatol": 1e-12, "maxiter": 20000,
                                                                  "maxfev": 40000})
        assert res.fun >= f_star - 1e-9                                   # nothing beats the closed form
        assert res.fun <= f_star + 1e-6 and np.abs(X @ res.x[:3] - ours).max() <= 1e-3 * np.abs(y).max()


def test_c_restatement_equals_numpy_oracle(oracle_golden):
    """oracle/mmf_oracle_c.c (the multi-core CPU baseline of bench.py) against the NumPy oracle, incl. gaps,
    an empty row and a rank-deficient mask."""
    import subprocess, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, capture_output=True)
    y = oracle_golden["daily365_y"].copy()
    y[0, :] = np.nan
    y[1, :] = np.nan; y[1, 40] = 3.0
    grid = O.calendar_grid(oracle_golden["daily365_start"][0].astype("datetime64[D]"), 365 + 28, "D")
    X = O.design_matrix(grid, 365)
    want, wst = O.fit_forecast_packed(y, X, 365, 365, 28)
    got, st = O.fit_forecast_packed_c(y, X, 365, 365, 28)
    assert np.array_equal(st, wst)
    ok = wst != 1
    assert np.isnan(got[~ok]).all() and np.abs(got[ok] - want[ok]).max() < 1e-8
    y2 = oracle_golden["daily1095_y"]
    grid = O.calendar_grid(oracle_golden["daily1095_start"][0].astype("datetime64[D]"), 1095, "D")
    Xh = O.design_matrix(grid, 1067)
    got, _ = O.fit_forecast_packed_c(y2, Xh, 1067, 0, 1095)
    assert np.abs(got - oracle_golden["daily1095_holdout"]).max() < 1e-7


def test_oracle_properties_on_random_series_and_masks():
    """Size-independent properties of the model, checked on the oracle itself (NumPy route and C restatement) for random
    lengths, horizons, scales and gap masks: shift equivariance f(y + c) = f(y) + c (the intercept is in the span),
    linearity for a common mask f(a y + b z) = a f(y) + b f(z), exact reproduction of series that lie in the design's span,
    and agreement of the two routes.  The GPU suite asserts the same properties at BASELINE sizes."""
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=40, deadline=None, derandomize=True, database=None,
                  suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(st.integers(60, 400), st.integers(1, 40), st.integers(0, 2**31 - 1), st.floats(0.0, 0.15),
               st.sampled_from(["D", "W-MON"]))
    def check(t, h, seed, gap_frac, freq):
        rng = np.random.default_rng(seed)
        grid = O.calendar_grid("2019-01-07", t + h, freq)
        X = O.design_matrix(grid, t)
        n = 6
        y = (rng.uniform(100, 20000, (n, 1)) + rng.normal(0, 50, (n, t)) + rng.uniform(-3, 3, (n, 1)) * np.arange(t))
        z = rng.uniform(0, 500, (n, t))
        mask = rng.random((n, t)) < gap_frac
        mask[:, 0] = False                                   # keep the centring value observed
        ym, zm = np.where(mask, np.nan, y), np.where(mask, np.nan, z)
        f = lambda a: O.fit_forecast_packed(a, X, t, t, h)   # noqa: E731
        base, st0 = f(ym)
        ok = st0 != 1                                        # status 1: no observed fit row (NaN out)
        scale = np.abs(y).max()
        shifted, _ = f(ym + 1234.5)
        assert np.allclose(shifted[ok], base[ok] + 1234.5, rtol=0, atol=1e-7 * scale + 1e-6)
        fz, st1 = f(zm)
        lin, st2 = f(2.0 * ym - 3.0 * zm)
        same = ok & (st0 == st1) & (st0 == st2)              # pivot dropping may differ only if a mask is degenerate
        assert np.allclose(lin[same], 2.0 * base[same] - 3.0 * fz[same], rtol=0, atol=1e-6 * scale + 1e-5)
        # a series in the span of the design (intercept + trend column) is reproduced exactly, gaps or not
        line = 50.0 + 7.0 * X[:t, 1]
        want = 50.0 + 7.0 * X[t:t + h, 1]
        got, _ = f(np.where(mask[:1], np.nan, line[None, :]))
        assert np.allclose(got[0], want, rtol=0, atol=1e-8 * (1 + np.abs(want).max()))
        # the C restatement (float32 input) agrees with the NumPy route on the same float32 values
        y32 = ym.astype(np.float32)
        a, sa = O.fit_forecast_packed(y32.astype(np.float64), X, t, t, h)
        b, sb = O.fit_forecast_packed_c(y32, X, t, t, h)
        assert np.array_equal(sa, sb)
        assert np.allclose(a, b, rtol=0, atol=1e-8 * scale + 1e-7, equal_nan=True)

    check()
