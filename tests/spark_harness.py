"""A pyarrow/pandas-only imitation of what PySpark does around a grouped-map pandas UDF -- TEST INFRASTRUCTURE.

There is no JVM or pyspark in this image, so the literal call site of the reference
(``enriched_df.groupBy("Product","SKU").applyInPandas(build_tune_and_score_model, schema=tuning_schema)``,
group_apply/02_Fine_Grained_Demand_Forecasting.py:523-528) cannot run.  What the UDF sees and what Spark accepts back
is fully determined by PySpark's Arrow serializers, restated here from PySpark >= 3.0's documented behaviour
(pyspark/sql/pandas/serializers.py, ArrowStreamPandasUDFSerializer / _create_batch, and
pyspark/sql/pandas/group_ops.py):

 in   one Arrow RecordBatch stream per group, typed by the DataFrame's schema (DateType -> date32, StringType ->
      string, FloatType -> float32); every column becomes a pandas Series with ``to_pandas(date_as_object=True)``
      -- so ``Date`` arrives as ``datetime.date`` OBJECTS -- and the Series are concatenated into one frame whose
      column labels are the field names.
 out  the returned frame's columns are matched to the declared return schema BY NAME when all labels are strings
      (else by position); a frame with a different number of columns is an error; every column is converted with
      ``pa.Array.from_pandas(series, mask=series.isnull(), type=<declared type>)`` -- datetime64 columns declared as
      DateType go through ``series.dt.date`` first -- so NaN/None become SQL NULL and a value that cannot be cast
      to the declared type fails the task.
 mapInArrow  the function receives an iterator of RecordBatches of a partition and yields RecordBatches whose
      schema must equal the declared one.
"""
from __future__ import annotations

import pandas as pd
import pyarrow as pa
import pyarrow.compute as pc


def _arrow_to_pandas_frame(batch: pa.RecordBatch) -> pd.DataFrame:
    table = pa.Table.from_batches([batch])
    series = []
    for name, col in zip(table.column_names, table.itercolumns()):
        s = col.to_pandas(date_as_object=True).rename(name)
        if pa.types.is_string(col.type) or pa.types.is_large_string(col.type):
            s = s.astype(object)          # what the PySpark releases the reference ran on (pandas < 3) hand to the UDF
        series.append(s)
    return pd.concat(series, axis=1) if series else pd.DataFrame()


def _pandas_to_arrow_batch(pdf: pd.DataFrame, schema: pa.Schema) -> pa.RecordBatch:
    if not isinstance(pdf, pd.DataFrame):
        raise TypeError(f"Return type of the user-defined function should be pandas.DataFrame, but is {type(pdf)}")
    if len(pdf.columns) != len(schema) and not (len(pdf.columns) == 0 and len(pdf) == 0):
        raise RuntimeError(f"Number of columns of the returned pandas.DataFrame doesn't match specified schema. "
                           f"Expected: {len(schema)} Actual: {len(pdf.columns)}")
    by_name = all(isinstance(c, str) for c in pdf.columns)
    arrays = []
    for i, field in enumerate(schema):
        s = pdf[field.name] if by_name else pdf.iloc[:, i]
        if pa.types.is_date32(field.type) and pd.api.types.is_datetime64_any_dtype(s.dtype):
            s = s.dt.date
        mask = s.isnull()
        if isinstance(s.dtype, pd.api.extensions.ExtensionDtype):      # e.g. pandas >= 3 "str": plain objects for Arrow
            s = s.astype(object)
        arrays.append(pa.Array.from_pandas(s, mask=mask, type=field.type, safe=True))
    return pa.RecordBatch.from_arrays(arrays, schema=schema)


def group_batches(table: pa.Table, keys, max_records_per_batch: int = 10_000):
    """One list of RecordBatches per group (what the JVM ships to a Python worker), groups in key order."""
    keys = list(keys)
    order = pc.sort_indices(table, sort_keys=[(k, "ascending") for k in keys])
    t = table.take(order)
    combo = [tuple(x) for x in zip(*[t.column(k).to_pylist() for k in keys])]
    start = 0
    for i in range(1, len(combo) + 1):
        if i == len(combo) or combo[i] != combo[start]:
            yield combo[start], t.slice(start, i - start).combine_chunks().to_batches(max_chunksize=max_records_per_batch)
            start = i


def apply_in_pandas(table: pa.Table, keys, func, return_schema: pa.Schema) -> pa.Table:
    """``table.groupBy(*keys).applyInPandas(func, schema=return_schema)`` without Spark."""
    out = []
    for _, batches in group_batches(table, keys):
        pdf = pd.concat([_arrow_to_pandas_frame(b) for b in batches], ignore_index=True)
        out.append(_pandas_to_arrow_batch(func(pdf), return_schema))
    return pa.Table.from_batches(out, schema=return_schema)


def map_in_arrow(table: pa.Table, func, return_schema: pa.Schema, n_partitions: int = 2,
                 max_records_per_batch: int = 10_000) -> pa.Table:
    """``df.mapInArrow(func, schema=return_schema)``: the rows are split into partitions; ``func`` gets an iterator
    of the partition's RecordBatches and yields batches that must carry the declared schema."""
    out = []
    rows = table.num_rows
    cuts = [rows * p // n_partitions for p in range(n_partitions + 1)]
    for p in range(n_partitions):
        part = table.slice(cuts[p], cuts[p + 1] - cuts[p]).combine_chunks()
        for b in func(iter(part.to_batches(max_chunksize=max_records_per_batch))):
            if not isinstance(b, pa.RecordBatch):
                raise TypeError(f"mapInArrow functions must yield pyarrow.RecordBatch, got {type(b)}")
            if not b.schema.equals(return_schema, check_metadata=False):
                raise RuntimeError(f"schema mismatch: {b.schema} vs declared {return_schema}")
            out.append(b)
    return pa.Table.from_batches(out, schema=return_schema)
