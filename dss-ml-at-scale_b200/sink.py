"""Sink for the forecast table.

The reference persists ``forecast_df`` as a Delta table and registers it
(group_apply/02_Fine_Grained_Demand_Forecasting.py:539-552).  Delta Lake is a Parquet directory plus a
transaction log; without Spark the closest faithful artefact is the Parquet data itself: ``tuning_schema`` columns
(02:498-506), key columns dictionary-encoded (one dictionary entry per group instead of T copies of each string),
``overwrite`` semantics like ``.mode("overwrite")`` (02:545).
"""
from __future__ import annotations

import os

from .frames import DEFAULT_KEYS, tuning_schema


def to_arrow(forecasts, keys=DEFAULT_KEYS, date_col="Date", value_col="Demand"):
    """pandas frame / Arrow table of forecasts -> Arrow table with ``tuning_schema`` and dictionary-encoded keys."""
    import pyarrow as pa

    if not isinstance(forecasts, pa.Table):
        forecasts = pa.Table.from_pandas(forecasts, schema=tuning_schema(keys, date_col, value_col), preserve_index=False)
    for k in keys:
        i = forecasts.schema.get_field_index(k)
        col = forecasts.column(k)
        if not pa.types.is_dictionary(col.type):
            forecasts = forecasts.set_column(i, k, col.dictionary_encode())
    return forecasts


def write_forecasts(forecasts, path: str, keys=DEFAULT_KEYS, date_col="Date", value_col="Demand",
                    mode: str = "overwrite") -> str:
    """Write the forecast table as Parquet (zstd, dictionary-encoded keys).  Returns the file path."""
    import pyarrow.parquet as pq

    if mode not in ("overwrite", "error"):
        raise ValueError("mode must be 'overwrite' or 'error'")
    if os.path.exists(path) and mode == "error":
        raise FileExistsError(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    table = to_arrow(forecasts, keys, date_col, value_col)
    tmp = path + ".tmp"
    pq.write_table(table, tmp, compression="zstd", use_dictionary=list(keys))
    os.replace(tmp, path)
    return path


def read_forecasts(path: str):
    """Read a table written by :func:`write_forecasts` back as pandas (keys as plain strings)."""
    import pyarrow.parquet as pq

    return pq.read_table(path, read_dictionary=[]).to_pandas()
