"""The literal drop-in (reference 02:523-528) exercised without a JVM: ``forecast_groups`` /
``forecast_arrow_batches`` are fed exactly what PySpark's grouped-map / mapInArrow serializers produce and must hand
back something those serializers accept (tests/spark_harness.py restates them).  On the CPU box the float64 oracle
stands in for the engine (host logic under test); the ``gpu`` variant runs the real engine."""
import datetime as dt

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import mmf
import spark_harness as SP
from oracle import mmf_oracle as O
from test_host import _OracleEngine


def _enriched_table(n_skus=2, drop=()):
    """The reference's UDF input: enriched_schema rows (02:360-370) as Spark would hold them."""
    df = mmf.synth.reference_weekly_demand(n_skus=n_skus)
    df = mmf.add_exo_variables(df)
    for sku_idx, date in drop:
        sku = sorted(df["SKU"].unique())[sku_idx]
        df = df[~((df["SKU"] == sku) & (df["Date"] == date))]
    df = df.sample(frac=1.0, random_state=4)                                      # Spark gives no row order
    return pa.Table.from_pandas(df, schema=mmf.enriched_schema(), preserve_index=False), df


def _udf(engine):
    def forecast(pdf):
        # what arrives: datetime.date objects, str keys, float32 demand, the three exog columns
        assert isinstance(pdf["Date"].iloc[0], dt.date) and not isinstance(pdf["Date"].iloc[0], dt.datetime)
        assert pdf["Demand"].dtype == np.float32 and isinstance(pdf["Product"].iloc[0], str)
        assert list(pdf.columns) == ["Date", "Product", "SKU", "Demand", "covid", "christmas", "new_year"]
        return mmf.forecast_groups(pdf, engine=engine)                            # defaults = the reference's (02:341,526)
    return forecast


def _check_against_reference_shaped_udf(got: pa.Table, df: pd.DataFrame, tol):
    want = SP.apply_in_pandas(pa.Table.from_pandas(df, schema=mmf.enriched_schema(), preserve_index=False),
                              ["Product", "SKU"], lambda p: O.build_tune_and_score_model(p), mmf.tuning_schema())
    assert got.schema.equals(mmf.tuning_schema())
    assert got.num_rows == want.num_rows
    for c in ("Product", "SKU", "Date"):
        assert got.column(c).to_pylist() == want.column(c).to_pylist(), c
    assert got.column("Demand").null_count == want.column("Demand").null_count
    a = got.column("Demand").to_pandas().to_numpy(dtype=np.float64, na_value=np.nan)
    b = want.column("Demand").to_pandas().to_numpy(dtype=np.float64, na_value=np.nan)
    assert np.array_equal(a, b, equal_nan=True)
    err = np.abs(got.column("Demand_Fitted").to_numpy() - want.column("Demand_Fitted").to_numpy()).max()
    assert err <= tol, (err, tol)


def test_apply_in_pandas_one_group_per_call_cpu():
    table, df = _enriched_table(2, drop=[(0, dt.date(2019, 5, 6))])
    got = SP.apply_in_pandas(table, ["Product", "SKU"], _udf(_OracleEngine()), mmf.tuning_schema())
    assert got.num_rows == 10 * 157
    _check_against_reference_shaped_udf(got, df, 1e-2)                            # oracle vs oracle through float32


def test_apply_in_pandas_many_groups_per_call_and_map_in_arrow_cpu():
    """The fast uses: groupBy(shard).applyInPandas (many groups per frame) and mapInArrow over partitions that hold
    whole groups (repartition by key first, as the reference does at 02:525)."""
    table, df = _enriched_table(2)
    shard = pa.array((pd.util.hash_pandas_object(df[["Product", "SKU"]], index=False).to_numpy() % 3).astype(np.int32))
    t2 = table.append_column("shard", shard)

    def per_shard(pdf):
        return mmf.forecast_groups(pdf.drop(columns=["shard"]), engine=_OracleEngine())

    a = SP.apply_in_pandas(t2, ["shard"], per_shard, mmf.tuning_schema())
    b = SP.apply_in_pandas(table, ["Product", "SKU"], _udf(_OracleEngine()), mmf.tuning_schema())
    key = [("Product", "ascending"), ("SKU", "ascending"), ("Date", "ascending")]
    assert a.sort_by(key).equals(b.sort_by(key))
    by_key = table.sort_by([("Product", "ascending"), ("SKU", "ascending")])      # partitions hold whole groups
    n_groups = 10
    c = SP.map_in_arrow(by_key, lambda it: mmf.forecast_arrow_batches(it, engine=_OracleEngine()), mmf.tuning_schema(),
                        n_partitions=2 if (n_groups % 2 == 0) else 1, max_records_per_batch=500)
    assert c.sort_by(key).equals(b.sort_by(key))


def test_serializer_rejects_wrong_frames():
    table, _ = _enriched_table(1)
    with pytest.raises(RuntimeError):
        SP.apply_in_pandas(table, ["Product", "SKU"], lambda p: p[["Product", "SKU"]], mmf.tuning_schema())
    with pytest.raises(Exception):                                                # a string where a float is declared
        SP.apply_in_pandas(table, ["Product", "SKU"],
                           lambda p: p.assign(Demand_Fitted="x")[["Product", "SKU", "Date", "Demand", "Demand_Fitted"]],
                           mmf.tuning_schema())


def test_null_keys_on_gap_rows_option():
    """Reference detail (02:490): rows that asfreq() inserts for missing dates carry NaN in Product / SKU, because the
    key columns are taken from the re-indexed frame.  The engine fills the keys by default (documented deviation,
    INTEGRATION.md); ``null_keys_on_gaps=True`` reproduces the reference's nulls."""
    table, df = _enriched_table(1, drop=[(0, dt.date(2019, 5, 6))])
    eng = _OracleEngine()
    ref = SP.apply_in_pandas(table, ["Product", "SKU"], lambda p: O.build_tune_and_score_model(p, null_keys_on_gaps=True),
                             mmf.tuning_schema())
    got = SP.apply_in_pandas(table, ["Product", "SKU"], lambda p: mmf.forecast_groups(p, engine=eng, null_keys_on_gaps=True),
                             mmf.tuning_schema())
    assert ref.column("SKU").null_count == 1 and got.column("SKU").null_count == 1
    assert got.column("Product").to_pylist() == ref.column("Product").to_pylist()
    filled = SP.apply_in_pandas(table, ["Product", "SKU"], _udf(eng), mmf.tuning_schema())
    assert filled.column("SKU").null_count == 0


@pytest.mark.gpu
def test_apply_in_pandas_and_map_in_arrow_on_gpu():
    from conftest import tolerance
    table, df = _enriched_table(3, drop=[(1, dt.date(2019, 7, 1))])
    eng = mmf.ForecastEngine()
    got = SP.apply_in_pandas(table, ["Product", "SKU"], _udf(eng), mmf.tuning_schema())
    _check_against_reference_shaped_udf(got, df, tolerance(df["Demand"].to_numpy()))
    by_key = table.sort_by([("Product", "ascending"), ("SKU", "ascending")])
    c = SP.map_in_arrow(by_key, lambda it: mmf.forecast_arrow_batches(it, engine=eng), mmf.tuning_schema(), n_partitions=1)
    key = [("Product", "ascending"), ("SKU", "ascending"), ("Date", "ascending")]
    assert c.sort_by(key).column("Demand_Fitted").equals(got.sort_by(key).column("Demand_Fitted"))
    d = SP.map_in_arrow(by_key, lambda it: mmf.forecast_arrow_batches(it, engine=eng, pack="device"), mmf.tuning_schema(),
                        n_partitions=1)
    assert d.sort_by(key).column("Demand_Fitted").equals(got.sort_by(key).column("Demand_Fitted"))
    eng.close()
