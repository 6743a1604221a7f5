"""Device-side packer: long-format Arrow rows -> padded series ``y[N, T]`` on the GPU.

The reference gets each group's rows together with a Spark hash shuffle
(``repartition(n_tasks, "Product", "SKU")`` + ``groupBy``, group_apply/
02_Fine_Grained_Demand_Forecasting.py:525-526) and then, per group, sorts by date and
re-indexes on the regular grid (``sort_values("Date")``, ``set_index("Date").asfreq(freq)``,
02:422-423).  Here the Arrow column buffers of the whole table go to the GPU unchanged and
four small kernels of ``libmmf.so`` (csrc/pack.cu) do the same for all groups at once:
64-bit hash of the key columns -> dense group codes (radix sort), checked against the key bytes -> per-group first/last day
-> scatter into NaN-filled rows.  Only G-sized metadata (first/last day and one key row per
group) ever comes back to the host.

torch is used for device buffers and copies only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from . import design as D
from .engine import ForecastEngine, default_engine


def _dev_from_numpy(a: np.ndarray, device):
    """Host array (often a read-only view of an Arrow buffer) -> device tensor; the host side is never written."""
    import warnings

    import torch

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False)


def _stage_keys(table, keys, n, device):
    """Key columns -> device buffers, as Arrow holds them: ("i32", values) or ("utf8", offsets, bytes)."""
    import pyarrow as pa
    import pyarrow.compute as pc

    staged = []
    for k in keys:
        col = table.column(k)
        col = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
        if col.null_count and not pa.types.is_dictionary(col.type):
            # the kernels read offsets / values, not the validity bitmap: give a null key its own dictionary code, so
            # that it stays distinct from "" or 0 exactly like the host packer (null_encoding="encode")
            col = pc.dictionary_encode(col, null_encoding="encode")
        elif col.null_count:
            col = pc.dictionary_encode(col.cast(col.type.value_type), null_encoding="encode")
        if pa.types.is_dictionary(col.type):
            idx = col.indices.cast(pa.int32()).to_numpy(zero_copy_only=False)
            staged.append(("i32", _dev_from_numpy(idx.astype(np.int32, copy=False), device)))
        elif pa.types.is_integer(col.type):
            staged.append(("i32", _dev_from_numpy(pc.cast(col, pa.int32()).to_numpy(zero_copy_only=False), device)))
        else:
            if pa.types.is_large_string(col.type):
                col = col.cast(pa.string())
            if not pa.types.is_string(col.type):
                col = pc.cast(col, pa.string())
            bufs = col.buffers()
            offsets = np.frombuffer(bufs[1], dtype=np.int32)[col.offset:col.offset + n + 1]
            data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
            staged.append(("utf8", _dev_from_numpy(offsets, device), _dev_from_numpy(data, device)))
    return staged


def _hash_keys(lib, h, staged, n, device, seed: int = 1):
    """Chain every key column into the per-row 64-bit hash (device).  ``seed`` >= 1 picks the FNV basis."""
    import torch

    hash_dev = torch.empty(n, dtype=torch.int64, device=device)      # bit container for uint64
    first = seed
    for col in staged:
        if col[0] == "i32":
            N.check(lib.mmf_pack_hash_i32(h, col[1].data_ptr(), n, hash_dev.data_ptr(), first))
        else:
            N.check(lib.mmf_pack_hash_utf8(h, col[1].data_ptr(), col[2].data_ptr(), n, hash_dev.data_ptr(), first))
        first = 0
    return hash_dev


def _count_collisions(lib, h, staged, n, gid, first_row, device) -> int:
    """Rows whose key differs from the key of their group's first row (two keys sharing one 64-bit hash)."""
    import torch

    bad = torch.zeros(1, dtype=torch.int64, device=device)
    for col in staged:
        if col[0] == "i32":
            N.check(lib.mmf_pack_verify_i32(h, col[1].data_ptr(), n, gid.data_ptr(), first_row.data_ptr(), bad.data_ptr()))
        else:
            N.check(lib.mmf_pack_verify_utf8(h, col[1].data_ptr(), col[2].data_ptr(), n, gid.data_ptr(),
                                             first_row.data_ptr(), bad.data_ptr()))
    return int(bad.item())


def group_rows_device(lib, h, staged, n, device, max_rehash: int = 3):
    """(gid[n], first_row[>=G], G): dense group code per row, verified against the key bytes; a detected hash
    collision re-hashes with another basis (and raises after ``max_rehash`` tries, never merges groups silently)."""
    import torch

    for seed in range(1, max_rehash + 1):
        hash_dev = _hash_keys(lib, h, staged, n, device, seed)
        gid = torch.empty(n, dtype=torch.int32, device=device)
        first_row = torch.empty(n, dtype=torch.int32, device=device)
        g_host = C.c_int32(0)
        N.check(lib.mmf_pack_group_codes(h, hash_dev.data_ptr(), n, gid.data_ptr(), first_row.data_ptr(), C.byref(g_host)))
        del hash_dev
        if _count_collisions(lib, h, staged, n, gid, first_row, device) == 0:
            return gid, first_row, int(g_host.value)
    raise RuntimeError(f"key hash collisions persisted over {max_rehash} hash bases")


def pack_table_device(table, keys=("Product", "SKU"), date_col="Date", value_col="Demand", freq="W-MON",
                      engine: ForecastEngine | None = None, sort_keys: bool = True):
    """Arrow ``Table`` (or pandas frame) of long-format rows -> list of ``frames.Bucket`` whose ``y`` are CUDA
    tensors (row pitch multiple of 4 floats, NaN = missing).  Same buckets, rows and values as the host packer
    ``frames.pack_groups`` (tests compare them bit for bit)."""
    import pandas as pd
    import pyarrow as pa
    import pyarrow.compute as pc
    import torch

    from .frames import Bucket

    if not isinstance(table, pa.Table):
        table = pa.Table.from_pandas(table, preserve_index=False)
    eng = engine or default_engine()
    lib, h = eng._lib, eng._h
    keys = list(keys)
    n = table.num_rows
    if n == 0:
        return []
    device = torch.device("cuda", torch.cuda.current_device())
    eng.set_stream(torch.cuda.current_stream(device).cuda_stream)
    step = D.FREQ_DAYS[freq]

    dcol = table.column(date_col).combine_chunks()
    if dcol.null_count:
        raise ValueError(f"{dcol.null_count} rows have a null {date_col}: a row without a date has no place on the grid")
    if not pa.types.is_date32(dcol.type):
        dcol = pc.cast(dcol, pa.date32())
    day = _dev_from_numpy(dcol.cast(pa.int32()).to_numpy(zero_copy_only=False), device)
    val = _dev_from_numpy(pc.cast(table.column(value_col).combine_chunks(), pa.float32())
                          .to_numpy(zero_copy_only=False).astype(np.float32, copy=False), device)

    staged = _stage_keys(table, keys, n, device)
    gid, first_row, G = group_rows_device(lib, h, staged, n, device)
    del staged
    gmin = torch.empty(G, dtype=torch.int32, device=device)
    gmax = torch.empty(G, dtype=torch.int32, device=device)
    N.check(lib.mmf_pack_minmax(h, gid.data_ptr(), day.data_ptr(), n, G, gmin.data_ptr(), gmax.data_ptr()))
    gmin_h = gmin.cpu().numpy().astype(np.int64)
    gmax_h = gmax.cpu().numpy().astype(np.int64)
    if freq == "W-MON" and np.any((gmin_h + 3) % 7 != 0):
        raise ValueError("W-MON series must start on a Monday")
    t_len = (gmax_h - gmin_h) // step + 1

    # one key row per group (G rows, host), optionally in key order like the host packer
    key_frame = table.select(keys).take(pa.array(first_row[:G].cpu().numpy())).to_pandas()
    key_frame = key_frame.astype({k: object for k in keys}) if len(key_frame) else key_frame
    order = (key_frame.sort_values(keys, kind="stable").index.to_numpy() if sort_keys else np.arange(G))

    buckets = []
    dups = torch.zeros(1, dtype=torch.int64, device=device)
    bucket_id, bucket_keys = pd.MultiIndex.from_arrays([gmin_h, t_len]).factorize(sort=True)
    for b, (start_day, tl) in enumerate(bucket_keys):
        members = order[bucket_id[order] == b]                        # group codes of this bucket, in key order
        row_of_group = np.full(G, -1, dtype=np.int64)
        row_of_group[members] = np.arange(members.size)
        rog = _dev_from_numpy(row_of_group, device)
        ld = (int(tl) + 3) & ~3
        full = torch.empty((members.size, ld), dtype=torch.float32, device=device)
        N.check(lib.mmf_pack_scatter_f32(h, gid.data_ptr(), day.data_ptr(), val.data_ptr(), n, rog.data_ptr(),
                                         gmin.data_ptr(), step, full.data_ptr(), members.size, ld, int(tl),
                                         dups.data_ptr()))
        buckets.append(Bucket(np.datetime64(int(start_day), "D"), int(tl),
                              key_frame.iloc[members].reset_index(drop=True), full[:, :int(tl)]))
    torch.cuda.current_stream(device).synchronize()                   # the staged input tensors may be freed now
    if int(dups.item()):
        # the reference's set_index("Date").asfreq() raises on duplicate dates within a group (02:423)
        raise ValueError(f"cannot reindex on an axis with duplicate labels: {int(dups.item())} rows repeat a "
                         f"({', '.join(keys)}, {date_col}) combination")
    return buckets
