"""cProfile of the Arrow boundary with the device packer (forecast_table(pack="device")): where the host time goes."""
import cProfile
import os
import pstats
import sys

import numpy as np
import pandas as pd
import pyarrow as pa

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmf  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T = 157
rng = np.random.default_rng(0)
days = mmf.design.calendar_grid(np.datetime64("2018-07-23"), T, "W-MON").astype("datetime64[D]")
prod = np.array([f"Product_{i % 5}" for i in range(G)], dtype=object)
sku = np.array([f"SKU_{i:07d}" for i in range(G)], dtype=object)
y = np.round(rng.uniform(3000, 18000, G).astype(np.float32)[:, None] + rng.normal(0, 100, (G, T)).astype(np.float32))
df = pd.DataFrame({"Date": np.tile(days, G), "Product": np.repeat(prod, T), "SKU": np.repeat(sku, T),
                   "Demand": y.reshape(-1).astype(np.float32)}).sample(frac=1.0, random_state=1).reset_index(drop=True)
table = pa.Table.from_pandas(df, preserve_index=False)
eng = mmf.default_engine()
for mode in ("device", "host"):
    fn = (lambda: mmf.forecast_table(table, engine=eng, pack=mode))
    fn(); fn()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        fn()
    pr.disable()
    print("=====", mode)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
