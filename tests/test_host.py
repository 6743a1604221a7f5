"""CPU: host-side logic -- packing, schema mirrors, sharding plan, and the world_size-2 all-gather (gloo)."""
import datetime as dt
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

import mmf
from mmf import sharding as SH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame():
    rows = []
    for k, (start, n) in enumerate(((dt.date(2021, 1, 4), 10), (dt.date(2021, 1, 4), 10), (dt.date(2021, 2, 1), 6))):
        for i in range(n):
            rows.append(("P%d" % (k % 2), "S%d" % k, start + dt.timedelta(weeks=i), float(100 * k + i)))
    return pd.DataFrame(rows, columns=["Product", "SKU", "Date", "Demand"])


def test_pack_groups_buckets_gaps_and_order():
    df = _frame()
    df = df[~((df["SKU"] == "S1") & (df["Date"] == dt.date(2021, 1, 18)))]          # a gap
    df = pd.concat([df, pd.DataFrame([("P1", "S1", dt.date(2021, 1, 20), 999.0)], columns=df.columns)])  # off-grid
    df = df.sample(frac=1.0, random_state=1)                                           # unsorted input
    buckets = mmf.pack_groups(df, freq="W-MON", pinned=False)
    assert [(str(b.start), b.t_len, len(b.key_frame)) for b in buckets] == [("2021-01-04", 10, 2), ("2021-02-01", 6, 1)]
    b0 = buckets[0]
    assert b0.key_frame["SKU"].tolist() == ["S0", "S1"]
    assert np.array_equal(b0.y[0], np.arange(10, dtype=np.float32))
    want = 100 + np.arange(10, dtype=np.float32)
    want[2] = np.nan
    assert np.array_equal(b0.y[1], want, equal_nan=True)
    assert b0.y.base is not None and b0.y.strides[0] % 16 == 0            # TMA-friendly pitch
    assert np.array_equal(buckets[1].y[0], 200 + np.arange(6, dtype=np.float32))


def test_pack_groups_daily_and_empty():
    assert mmf.pack_groups(_frame().iloc[:0], freq="W-MON", pinned=False) == []
    days = pd.date_range("2021-03-01", periods=40, freq="D")
    df = pd.DataFrame({"store": "a", "item": "b", "d": days, "q": np.arange(40.0)})
    (b,) = mmf.pack_groups(df, keys=("store", "item"), date_col="d", value_col="q", freq="D", pinned=False)
    assert b.t_len == 40 and np.array_equal(b.y[0], np.arange(40, dtype=np.float32))


class _OracleEngine:
    """Stands in for ForecastEngine where there is no GPU (test infrastructure only): same plan/fit interface."""

    def plan_calendar(self, start, t_len, freq="D", horizon=28, mode="future", design="trend_season_exog"):
        if mode == "holdout":
            if t_len - horizon < 1:
                raise ValueError("series shorter than the forecast horizon")      # as ForecastEngine.plan_calendar
            self.t_fit, n_rows, ps, npred = t_len - horizon, t_len, 0, t_len
        else:
            self.t_fit, n_rows, ps, npred = t_len, t_len + horizon, t_len, horizon
        days = mmf.design.calendar_grid(start, n_rows, freq)
        self.X = mmf.design.design_matrix(days, self.t_fit, design)
        return days[ps:ps + npred], ps, npred

    def fit_forecast(self, y, pred_start, n_pred, out=None):
        from oracle import mmf_oracle as O
        return O.fit_forecast_packed(np.asarray(y), self.X, self.t_fit, pred_start, n_pred)[0].astype(np.float32)


def test_arrow_packer_equals_pandas_packer():
    import pyarrow as pa
    df = _frame()
    df = df[~((df["SKU"] == "S1") & (df["Date"] == dt.date(2021, 1, 18)))].sample(frac=1.0, random_state=3)
    host = mmf.pack_groups(df, freq="W-MON", pinned=False)
    t = pa.Table.from_pandas(df, preserve_index=False)
    for table in (t, t.set_column(1, "SKU", pa.array(df["SKU"]).dictionary_encode()),
                  pa.Table.from_batches(t.to_batches(max_chunksize=7))):                 # chunked columns
        got = mmf.frames.pack_table_host(table, freq="W-MON", pinned=False)
        assert len(got) == len(host) == 2
        for a, b in zip(got, host):
            assert (a.start, a.t_len) == (b.start, b.t_len)
            assert a.key_frame.to_numpy().tolist() == b.key_frame.to_numpy().tolist()
            assert np.array_equal(a.y, b.y, equal_nan=True) and np.array_equal(a.rank, b.rank)


def test_group_codes_many_key_columns_in_key_order():
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"a": rng.integers(0, 4, 500), "b": rng.choice(list("xyz"), 500), "c": rng.integers(0, 3, 500),
                       "Date": dt.date(2021, 1, 4), "Demand": 1.0}).drop_duplicates(["a", "b", "c"])
    (b,) = mmf.pack_groups(df, keys=("a", "b", "c"), freq="W-MON", pinned=False)
    want = df[["a", "b", "c"]].sort_values(["a", "b", "c"]).to_numpy().tolist()
    assert b.key_frame.to_numpy().tolist() == want and b.key_frame["a"].dtype == df["a"].dtype


def test_forecast_groups_and_table_buckets_in_key_order():
    """Groups on different calendars (two buckets) come back interleaved in key order, dates ascending -- from the
    pandas route and the Arrow route alike; the Arrow route returns tuning_schema with NaN as null."""
    import pyarrow as pa
    rows = []
    for k, (start, n) in enumerate(((dt.date(2021, 1, 4), 30), (dt.date(2021, 2, 1), 26), (dt.date(2021, 1, 4), 30),
                                    (dt.date(2021, 2, 1), 26))):
        for i in range(n):
            rows.append(("P%d" % (k % 2), "S%d" % k, start + dt.timedelta(weeks=i), float(50 * k + 3 * i + (i % 5))))
    df = pd.DataFrame(rows, columns=["Product", "SKU", "Date", "Demand"])
    df = df[~((df["SKU"] == "S2") & (df["Date"] == dt.date(2021, 3, 1)))].sample(frac=1.0, random_state=2)
    kw = dict(horizon=4, mode="holdout", design="exog_only", engine=_OracleEngine())
    out = mmf.forecast_groups(df, **kw)
    assert list(out.columns) == ["Product", "SKU", "Date", "Demand", "Demand_Fitted"]
    assert out["SKU"].tolist() == ["S0"] * 30 + ["S2"] * 30 + ["S1"] * 26 + ["S3"] * 26          # (Product, SKU) order
    for _, g in out.groupby("SKU"):
        assert g["Date"].is_monotonic_increasing
    assert np.isnan(out[(out["SKU"] == "S2")]["Demand"].to_numpy()).sum() == 1
    one = mmf.forecast_groups(df[df["SKU"] == "S3"], **kw)                               # a group alone == inside a batch
    assert np.array_equal(one["Demand_Fitted"].to_numpy(), out[out["SKU"] == "S3"]["Demand_Fitted"].to_numpy())
    tab = mmf.forecast_table(pa.Table.from_pandas(df, preserve_index=False), **kw)
    assert tab.schema.equals(mmf.tuning_schema()) and tab.num_rows == len(out)
    assert tab.column("SKU").to_pylist() == out["SKU"].tolist()
    assert tab.column("Date").to_pylist() == [d.date() for d in out["Date"]]
    assert tab.column("Demand").null_count == 1
    assert np.array_equal(tab.column("Demand_Fitted").to_numpy(), out["Demand_Fitted"].to_numpy())
    fut = mmf.forecast_table(pa.Table.from_pandas(df, preserve_index=False), horizon=4, mode="future",
                             design="exog_only", engine=_OracleEngine())
    assert fut.num_rows == 4 * 4 and fut.column("Demand").null_count == 16
    assert mmf.forecast_table(pa.Table.from_pandas(df.iloc[:0], preserve_index=False), **kw).num_rows == 0


def test_add_exo_variables_mirror(reference_fixtures):
    days = reference_fixtures["exo_weekly_days"].astype("datetime64[D]")
    pdf = pd.DataFrame({"Date": [dt.date.fromisoformat(str(d)) for d in days], "Product": "P", "SKU": "S",
                        "Demand": np.float32(1)})
    out = mmf.add_exo_variables(pdf)
    assert list(out.columns) == ["Date", "Product", "SKU", "Demand", "covid", "christmas", "new_year"]   # 02:358
    assert np.array_equal(out[["covid", "christmas", "new_year"]].to_numpy(), reference_fixtures["exo_weekly"])


def test_schemas():
    ts = mmf.tuning_schema()
    assert ts.names == ["Product", "SKU", "Date", "Demand", "Demand_Fitted"]          # 02:498-506
    assert mmf.enriched_schema().names == ["Date", "Product", "SKU", "Demand", "covid", "christmas", "new_year"]
    assert mmf.FORECAST_HORIZON == 40 and mmf.DEFAULT_KEYS == ("Product", "SKU")


def test_shard_plan():
    keys = pd.DataFrame({"Product": ["a", "b", "c", "d", "e", "f", "g"], "SKU": list("1234567")})
    h1, h2 = SH.stable_hash64(keys), SH.stable_hash64(keys.copy())
    assert np.array_equal(h1, h2) and len(set(h1.tolist())) == 7
    owner = SH.owner_of_keys(keys, 3)
    plans = [SH.ShardPlan.build(owner, 3, r) for r in range(3)]
    assert sorted(np.concatenate([p.local_rows for p in plans]).tolist()) == list(range(7))
    gi = plans[0].gather_index()
    assert len(set(gi.tolist())) == 7 and gi.max() < 3 * plans[0].per
    for p in plans:
        assert np.array_equal(p.slot[p.local_rows], np.arange(p.local_rows.size))
    assert SH.owner_of_rows(10, 4).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3]


def test_all_gather_world2_gloo():
    """N>1 path on CPU: two gloo ranks each 'fit' their hash shard (the oracle stands in for the GPU
    engine, which cannot run here) and one all_gather reassembles the table in the original order."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK rank 0" in r.stdout and "GLOO_OK rank 1" in r.stdout


def test_sink_roundtrip(tmp_path):
    """the forecast table written with tuning_schema (02:498-506, sink 02:539-552)."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    df = pd.DataFrame({"Product": ["a"] * 3 + ["b"] * 3, "SKU": ["s1"] * 3 + ["s2"] * 3,
                       "Date": [dt.date(2021, 1, 4) + dt.timedelta(weeks=i) for i in range(3)] * 2,
                       "Demand": np.array([1, 2, np.nan, 4, 5, 6], dtype=np.float32),
                       "Demand_Fitted": np.arange(6, dtype=np.float32)})
    path = mmf.sink.write_forecasts(df, str(tmp_path / "out" / "forecasts.parquet"))
    meta = pq.read_schema(path)
    assert meta.names == ["Product", "SKU", "Date", "Demand", "Demand_Fitted"]
    assert pa.types.is_dictionary(meta.field("Product").type) and meta.field("Date").type == pa.date32()
    back = mmf.sink.read_forecasts(path)
    assert back["SKU"].tolist() == df["SKU"].tolist() and back["Date"].tolist() == df["Date"].tolist()
    assert np.array_equal(back["Demand_Fitted"].to_numpy(), df["Demand_Fitted"].to_numpy())
    assert np.isnan(back["Demand"].iloc[2])
    mmf.sink.write_forecasts(df.iloc[:3], path)                       # overwrite
    assert len(mmf.sink.read_forecasts(path)) == 3
    with pytest.raises(FileExistsError):
        mmf.sink.write_forecasts(df, path, mode="error")


def test_duplicate_dates_raise_like_asfreq():
    """The reference's set_index("Date").asfreq() raises on duplicate dates inside a group (02:423); so do the packers."""
    import pyarrow as pa
    df = _frame()
    df = pd.concat([df, df.iloc[[3]]], ignore_index=True)
    with pytest.raises(ValueError, match="duplicate"):
        mmf.pack_groups(df, freq="W-MON", pinned=False)
    with pytest.raises(ValueError, match="duplicate"):
        mmf.frames.pack_table_host(pa.Table.from_pandas(df, preserve_index=False), freq="W-MON", pinned=False)


def test_single_group_fast_path_equals_general_path():
    """The literal drop-in (one group per call, 02:527) takes a short cut around the many-groups machinery; it must
    return exactly what the general path returns: gaps, off-grid rows, unsorted input, both modes, null keys on gaps,
    non-string keys, datetime64 dates; more than one group or a null key falls through to the general path."""
    from mmf import frames as F
    df = mmf.synth.reference_weekly_demand(n_skus=2)
    sku = df["SKU"].iloc[0]
    one = df[df["SKU"] == sku].copy()
    one = one[one["Date"] != dt.date(2019, 5, 6)]                                           # a gap
    extra = pd.DataFrame([(one["Product"].iloc[0], sku, dt.date(2019, 5, 8), 123.0)], columns=one.columns)
    one = pd.concat([one, extra.astype({"Demand": np.float32})]).sample(frac=1.0, random_state=1)   # off-grid row, shuffled
    as_ts = one.assign(Date=pd.to_datetime(one["Date"]))                                    # datetime64 dates
    num = one.assign(Product=7, SKU=np.int64(42))                                           # numeric keys
    eng = _OracleEngine()

    def both(frame, **kw):
        fast = mmf.forecast_groups(frame, engine=eng, **kw)
        keep, F.SINGLE_GROUP_MAX_ROWS = F.SINGLE_GROUP_MAX_ROWS, 0
        try:
            general = mmf.forecast_groups(frame, engine=eng, **kw)
        finally:
            F.SINGLE_GROUP_MAX_ROWS = keep
        pd.testing.assert_frame_equal(fast, general)
        return fast

    for frame in (one, as_ts, num):
        for kw in ({}, {"mode": "future", "horizon": 8}, {"null_keys_on_gaps": True}, {"design": "exog_only"}):
            out = both(frame, **kw)
            assert len(out) == (157 if kw.get("mode") != "future" else 8)
    assert F._single_group_fast(df, ["Product", "SKU"], "Date", "Demand", "W-MON", 40, "holdout", "trend_season_exog",
                                eng, False) is None                                         # two groups: general path
    nul = one.assign(SKU=None)
    assert F._single_group_fast(nul, ["Product", "SKU"], "Date", "Demand", "W-MON", 40, "holdout", "trend_season_exog",
                                eng, False) is None
    with pytest.raises(ValueError, match="duplicate"):
        mmf.forecast_groups(pd.concat([one, one.iloc[[3]]]), engine=eng)
    daily = pd.DataFrame({"store": "a", "item": "b", "d": pd.date_range("2021-03-01", periods=90, freq="D"),
                          "q": np.arange(90, dtype=np.float32)})
    both(daily, keys=("store", "item"), date_col="d", value_col="q", freq="D", horizon=14, mode="future")


# ---- the packer against pandas itself: what the reference executes per group is sort_values + set_index + asfreq ----
def _asfreq_reference(df, freq):
    """02:422-423 verbatim in spirit, per group: the regular grid pandas builds and the values it leaves on it."""
    out = {}
    for key, g in df.groupby(["Product", "SKU"], sort=True):
        s = g.sort_values("Date").set_index("Date")["Demand"].asfreq(freq)
        out[key] = (s.index[0].date(), len(s), s.to_numpy(dtype=np.float32))
    return out


def test_packer_equals_pandas_asfreq_on_random_frames():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @st.composite
    def frames(draw):
        freq = draw(st.sampled_from(["D", "W-MON"]))
        step = 1 if freq == "D" else 7
        rows = []
        for gi in range(draw(st.integers(1, 5))):
            start = dt.date(2021, 1, 4) + dt.timedelta(days=step * draw(st.integers(0, 30)))     # a Monday
            n = draw(st.integers(1, 40))
            keep = draw(st.lists(st.booleans(), min_size=n, max_size=n))
            if not any(keep):
                keep[draw(st.integers(0, n - 1))] = True
            first, last = keep.index(True), n - 1 - keep[::-1].index(True)
            for i in range(n):
                if keep[i]:
                    v = draw(st.one_of(st.integers(0, 60000).map(float), st.just(float("nan"))))
                    rows.append(("P%d" % (gi % 2), "S%d" % gi, start + dt.timedelta(days=step * i), v))
                    if step == 7 and first < i < last and draw(st.integers(0, 9)) == 0:     # an off-grid row: asfreq drops it
                        rows.append(("P%d" % (gi % 2), "S%d" % gi, start + dt.timedelta(days=step * i + 2), 777.0))
        order = draw(st.permutations(list(range(len(rows)))))
        df = pd.DataFrame([rows[i] for i in order], columns=["Product", "SKU", "Date", "Demand"])
        df["Date"] = pd.to_datetime(df["Date"])
        return freq, df

    @hyp.settings(max_examples=60, deadline=None, derandomize=True, database=None,
                  suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(frames())
    def check(case):
        freq, df = case
        want = _asfreq_reference(df, freq)
        got = {}
        for b in mmf.pack_groups(df, freq=freq, pinned=False):
            for r, key in enumerate(b.key_frame.itertuples(index=False)):
                got[(key.Product, key.SKU)] = (b.start.astype("datetime64[D]").astype(object), b.t_len, np.array(b.y[r, :b.t_len]))
        assert got.keys() == want.keys()
        for key, (start, t_len, vals) in want.items():
            gs, gt, gv = got[key]
            assert (gs, gt) == (start, t_len), key
            assert np.array_equal(gv, vals, equal_nan=True), key
        import pyarrow as pa
        table = pa.Table.from_pandas(df, preserve_index=False)
        got2 = {}
        for b in mmf.frames.pack_table_host(table, freq=freq, pinned=False):
            for r, key in enumerate(b.key_frame.itertuples(index=False)):
                got2[(key.Product, key.SKU)] = (b.t_len, np.array(b.y[r, :b.t_len]))
        for key, (_, t_len, vals) in want.items():
            assert got2[key][0] == t_len and np.array_equal(got2[key][1], vals, equal_nan=True), key

    check()


def test_groups_not_longer_than_the_horizon_fail_like_the_reference():
    """A group with no more rows than ``forecast_horizon`` has nothing to train on: the reference's
    ``split_train_score_data`` builds a mask of the wrong length / an empty train frame and the Spark task fails
    (02:372-380, 441-450).  The engine's planning raises before any GPU work -- exercised here on the real
    ``ForecastEngine`` methods without a context -- and ``forecast_groups`` lets the error through."""
    eng = object.__new__(mmf.ForecastEngine)                 # planning checks its arguments before it touches the library
    for t_len in (1, 5, 40):
        with pytest.raises(ValueError, match="shorter than the forecast horizon"):
            eng.plan_calendar("2021-01-04", t_len, "W-MON", 40, "holdout")
    with pytest.raises(ValueError, match="shorter than the forecast horizon"):
        eng.plan_calendars(["2021-01-04", "2021-01-04"], [200, 40], "W-MON", 40, "trend_season_exog", "holdout")
    with pytest.raises(ValueError, match="mode must be"):
        eng.plan_calendar("2021-01-04", 100, "W-MON", 40, "backtest")
    short = _frame()                                         # 10 / 10 / 6 weekly rows per group
    with pytest.raises(ValueError, match="shorter than the forecast horizon"):
        mmf.forecast_groups(short, engine=_OracleEngine(), horizon=40, mode="holdout")
    out = mmf.forecast_groups(short, engine=_OracleEngine(), horizon=40, mode="future")     # forecasting past the end is fine
    assert len(out) == 3 * 40
