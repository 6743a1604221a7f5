"""DataFrame / Arrow boundary throughput: long-format rows in -> tuning_schema rows out (holdout mode, the
reference's contract, group_apply/02_Fine_Grained_Demand_Forecasting.py:417-528) through
forecast_groups (pandas), forecast_table (Arrow, host packer) and forecast_table(pack="device").

    python scripts/bench_frames.py [n_groups=20000] [weeks=157]
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import pyarrow as pa

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmf  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 157
    rng = np.random.default_rng(0)
    days = mmf.design.calendar_grid(np.datetime64("2018-07-23"), T, "W-MON").astype("datetime64[D]")
    prod = np.array([f"Product_{i % 5}" for i in range(G)], dtype=object)
    sku = np.array([f"SKU_{i:07d}" for i in range(G)], dtype=object)
    level = rng.uniform(3000, 18000, G).astype(np.float32)
    y = np.round(level[:, None] + rng.normal(0, 100, (G, T)).astype(np.float32) + 4 * np.arange(T, dtype=np.float32))
    df = pd.DataFrame({"Date": np.tile(days, G), "Product": np.repeat(prod, T), "SKU": np.repeat(sku, T),
                       "Demand": y.reshape(-1).astype(np.float32)})
    df = df.sample(frac=1.0, random_state=1).reset_index(drop=True)            # arbitrary row order, like a shuffle
    table = pa.Table.from_pandas(df, preserve_index=False)
    eng = mmf.default_engine()
    res = {"groups": G, "weeks": T, "rows": len(df), "host_cores": os.cpu_count()}
    ref = None
    for name, fn in (("forecast_groups_pandas", lambda: mmf.forecast_groups(df, engine=eng)),
                     ("forecast_table_arrow_host_pack", lambda: mmf.forecast_table(table, engine=eng)),
                     ("forecast_table_arrow_device_pack", lambda: mmf.forecast_table(table, engine=eng, pack="device"))):
        fn()                                                                    # warm-up (plan, allocator, page-in)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            out = fn()
            best = min(best, time.perf_counter() - t0)
        fitted = (out["Demand_Fitted"].to_numpy() if isinstance(out, pd.DataFrame)
                  else out.column("Demand_Fitted").to_numpy())
        if ref is None:
            ref = fitted
        res[name] = {"seconds": best, "rows_per_s": len(df) / best, "groups_per_s": G / best,
                     "max_abs_diff_vs_pandas_route": float(np.abs(fitted - ref).max())}
    # the literal drop-in: one call per group, as applyInPandas(forecast_groups) makes them (02:523-528)
    parts = [g for _, g in df[df["SKU"].isin(sku[:300])].groupby(["Product", "SKU"], sort=False)]
    mmf.forecast_groups(parts[0], engine=eng)
    t0 = time.perf_counter()
    for g in parts:
        mmf.forecast_groups(g, engine=eng)
    dt = time.perf_counter() - t0
    res["one_group_per_call"] = {"calls": len(parts), "seconds": dt, "groups_per_s": len(parts) / dt,
                                 "ms_per_call": 1e3 * dt / len(parts)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
