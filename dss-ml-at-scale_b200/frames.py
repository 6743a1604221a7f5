"""DataFrame-in / DataFrame-out boundary: the drop-in for the reference's grouped-map UDF.

``forecast_groups`` has the contract of ``build_tune_and_score_model``
(group_apply/02_Fine_Grained_Demand_Forecasting.py:417-494) -- rows of
``enriched_schema`` (02:360-370) in, rows of ``tuning_schema`` (02:498-506) out,
one output row per date of each group's regular grid, sorted by date -- but it
accepts ANY number of groups in one frame and fits them in one GPU pass.  It
can therefore be passed to ``applyInPandas`` unchanged (one group per call, the
literal drop-in at 02:527) or, the intended fast use, once per shard:

    df.groupBy(shard_id).applyInPandas(forecast_groups, schema=tuning_schema)

The packer replaces the per-group ``sort_values("Date")`` +
``set_index("Date").asfreq(freq)`` (02:422-423) with one vectorised scatter into
padded ``y[N, T]`` float32 rows (NaN = missing).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import design as D
from .engine import ForecastEngine, alloc_packed, default_engine

FORECAST_HORIZON = 40                      # 02:341
DEFAULT_KEYS = ("Product", "SKU")          # 02:526
EXO_FIELDS = ("covid", "christmas", "new_year")   # 02:431


# ---- schemas (02:360-370, 02:498-506) as pyarrow; pyspark StructTypes on demand ---------
def tuning_schema(keys=DEFAULT_KEYS, date_col="Date", value_col="Demand"):
    import pyarrow as pa

    return pa.schema([(k, pa.string()) for k in keys]
                     + [(date_col, pa.date32()), (value_col, pa.float32()), (value_col + "_Fitted", pa.float32())])


def enriched_schema(keys=DEFAULT_KEYS, date_col="Date", value_col="Demand"):
    import pyarrow as pa

    return pa.schema([(date_col, pa.date32())] + [(k, pa.string()) for k in keys]
                     + [(value_col, pa.float32())] + [(c, pa.float32()) for c in EXO_FIELDS])


def spark_schemas(keys=DEFAULT_KEYS, date_col="Date", value_col="Demand"):
    """(enriched_schema, tuning_schema) as pyspark StructTypes -- needs pyspark."""
    from pyspark.sql.types import DateType, FloatType, StringType, StructField, StructType

    enriched = StructType([StructField(date_col, DateType())] + [StructField(k, StringType()) for k in keys]
                          + [StructField(value_col, FloatType())] + [StructField(c, FloatType()) for c in EXO_FIELDS])
    tuning = StructType([StructField(k, StringType()) for k in keys]
                        + [StructField(date_col, DateType()), StructField(value_col, FloatType()),
                           StructField(value_col + "_Fitted", FloatType())])
    return enriched, tuning


# ---- mirrors of the small reference helpers ------------------------------------------------
def add_exo_variables(pdf: pd.DataFrame, keys=DEFAULT_KEYS, date_col="Date", value_col="Demand") -> pd.DataFrame:
    """Vectorised ``add_exo_variables`` (02:343-358): same columns, same order, same 0/1 floats."""
    exo = D.exo_variables(D.as_days(pdf[date_col].to_numpy()))
    out = pdf.assign(covid=exo[:, 0], christmas=exo[:, 1], new_year=exo[:, 2])
    return out[[date_col, *keys, value_col, *EXO_FIELDS]]


def split_train_score_data(data, forecast_horizon: int = FORECAST_HORIZON):
    """02:372-380: first ``len - horizon`` rows train, last ``horizon`` rows score."""
    n = len(data)
    is_history = np.arange(n) < (n - forecast_horizon)
    if hasattr(data, "iloc"):
        return data.iloc[is_history], data.iloc[~is_history]
    return data[is_history], data[~is_history]


# ---- packing ---------------------------------------------------------------------------------
@dataclass
class Bucket:
    """Groups that share one calendar (same first date, same grid length)."""
    start: np.datetime64
    t_len: int
    key_frame: pd.DataFrame      # one row per series, key columns only
    y: np.ndarray                # [n, t_len] float32 view of a pitched (pinned) buffer, NaN = missing


def pack_groups(pdf: pd.DataFrame, keys=DEFAULT_KEYS, date_col="Date", value_col="Demand", freq="W-MON",
                pinned: bool | None = None) -> list:
    """Long frame -> calendar buckets of packed series (02:422-423 for all groups at once).
    ``pinned=None`` page-locks buckets of >= 1 MiB (worth the cudaHostAlloc), False never."""
    keys = list(keys)
    step = D.FREQ_DAYS[freq]
    if len(pdf) == 0:
        return []
    days = D.as_days(pdf[date_col].to_numpy()).astype(np.int64)
    gid, uniq = pd.MultiIndex.from_frame(pdf[keys]).factorize(sort=True)
    n_groups = len(uniq)
    vals = pdf[value_col].to_numpy(dtype=np.float32, na_value=np.nan)
    gmin = np.full(n_groups, np.iinfo(np.int64).max)
    gmax = np.full(n_groups, np.iinfo(np.int64).min)
    np.minimum.at(gmin, gid, days)
    np.maximum.at(gmax, gid, days)
    if freq == "W-MON" and np.any((gmin + 3) % 7 != 0):
        raise ValueError("W-MON series must start on a Monday")
    t_len = (gmax - gmin) // step + 1
    off = days - gmin[gid]
    on_grid = off % step == 0                       # off-grid rows vanish under asfreq
    pos = off // step
    key_frame = pd.DataFrame(list(uniq), columns=keys)
    buckets = []
    bucket_id, bucket_keys = pd.MultiIndex.from_arrays([gmin, t_len]).factorize(sort=True)
    for b, (start_day, tl) in enumerate(bucket_keys):
        members = np.flatnonzero(bucket_id == b)
        local = np.full(n_groups, -1, dtype=np.int64)
        local[members] = np.arange(members.size)
        pin = (members.size * int(tl) * 4 >= (1 << 20)) if pinned is None else pinned
        y = alloc_packed(members.size, int(tl), pinned=pin)
        y[...] = np.nan
        sel = on_grid & (local[gid] >= 0)
        y[local[gid[sel]], pos[sel]] = vals[sel]
        buckets.append(Bucket(np.datetime64(int(start_day), "D"), int(tl),
                              key_frame.iloc[members].reset_index(drop=True), y))
    return buckets


# ---- the drop-in UDF ---------------------------------------------------------------------------
def forecast_groups(pdf, *, keys=DEFAULT_KEYS, date_col="Date", value_col="Demand",
                    freq="W-MON", horizon=FORECAST_HORIZON, mode="holdout", design="trend_season_exog",
                    engine: ForecastEngine | None = None, pack: str = "host", select=None) -> pd.DataFrame:
    """Fit + forecast every group in ``pdf``; returns ``tuning_schema`` rows
    (keys..., Date, Demand, Demand_Fitted), groups in key order, dates ascending.

    ``mode="holdout"`` reproduces the reference's contract (fit on all but the last
    ``horizon`` grid rows, emit fitted + forecast values for every grid date, 02:484-494);
    ``mode="future"`` fits on everything and emits the ``horizon`` dates after the end
    (``Demand`` is NaN there).

    ``select=(1, 3, 9, 13, 16)`` (holdout mode only) turns on the per-series model selection that stands in for the
    reference's hyperopt loop (02:435-481): nested designs on the first ``m`` whitened columns are scored by their
    MSE over the held-out rows and the winner produces ``Demand_Fitted`` (``ForecastEngine.fit_select_forecast``).

    ``pack="device"`` groups, sorts and re-grids the rows on the GPU (``packer.pack_table_device``: Arrow
    buffers in, padded series out) instead of with pandas on the host; ``pdf`` may then be an Arrow table.
    """
    eng = engine or default_engine()
    keys = list(keys)
    fitted_col = value_col + "_Fitted"
    parts = []
    if pack == "device":
        from .packer import pack_table_device
        buckets = pack_table_device(pdf, keys, date_col, value_col, freq, engine=eng)
    elif pack == "host":
        buckets = pack_groups(pdf, keys, date_col, value_col, freq)
    else:
        raise ValueError("pack must be 'host' or 'device'")
    for b in buckets:
        out_days, pred_start, n_pred = eng.plan_calendar(b.start, b.t_len, freq, horizon, mode, design)
        if select is not None:
            if mode != "holdout":
                raise ValueError("select= needs mode='holdout' (the held-out rows score the candidates)")
            from .engine import device_packed
            yd = b.y if pack == "device" else device_packed(b.y)
            pred = eng.fit_select_forecast(yd, horizon, tuple(select), pred_start, n_pred)["pred"]
            pred = pred.cpu().numpy()
            if pack == "device":
                b.y = b.y.cpu().numpy()
        else:
            pred = eng.fit_forecast(b.y, pred_start, n_pred)
            if pack == "device":
                pred = pred.cpu().numpy()
                b.y = b.y.cpu().numpy()
        n = b.y.shape[0]
        frame = {k: np.repeat(b.key_frame[k].to_numpy(), n_pred) for k in keys}
        frame[date_col] = np.tile(out_days.astype("datetime64[ns]"), n)
        if mode == "holdout":
            frame[value_col] = np.ascontiguousarray(b.y).reshape(-1)
        else:
            frame[value_col] = np.full(n * n_pred, np.nan, dtype=np.float32)
        frame[fitted_col] = pred.reshape(-1)
        parts.append(pd.DataFrame(frame))
    if not parts:
        return pd.DataFrame({**{k: pd.Series(dtype=object) for k in keys},
                             date_col: pd.Series(dtype="datetime64[ns]"),
                             value_col: pd.Series(dtype=np.float32), fitted_col: pd.Series(dtype=np.float32)})
    out = parts[0] if len(parts) == 1 else pd.concat(parts, ignore_index=True)
    if len(parts) > 1:
        out = out.sort_values(keys + [date_col], kind="stable", ignore_index=True)
    return out


def forecast_table(table, **kw):
    """Arrow ``Table``/``RecordBatch`` in -> Arrow ``Table`` with ``tuning_schema`` out
    (the ``mapInArrow`` flavour of the boundary)."""
    import pyarrow as pa

    if isinstance(table, pa.RecordBatch):
        table = pa.Table.from_batches([table])
    keys = kw.get("keys", DEFAULT_KEYS)
    date_col, value_col = kw.get("date_col", "Date"), kw.get("value_col", "Demand")
    out = forecast_groups(table if kw.get("pack") == "device" else table.to_pandas(), **kw)
    return pa.Table.from_pandas(out, schema=tuning_schema(keys, date_col, value_col), preserve_index=False)


def forecast_arrow_batches(batches, **kw):
    """``mapInArrow`` adapter: an iterator of RecordBatches (one Spark partition) -> batches."""
    import pyarrow as pa

    batches = list(batches)
    if not batches:
        return
    yield from forecast_table(pa.Table.from_batches(batches), **kw).to_batches()
