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
