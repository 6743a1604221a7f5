"""Mint golden vectors from the float64 oracle (the reference's SARIMAX arithmetic cannot run
here -- see oracle/mmf_oracle.py "PARITY STATUS").  Inputs come from the restated synthetic
recipes; outputs are what ``oracle.mmf_oracle`` computes today, frozen so that neither the
oracle nor the CUDA path can drift silently.
    python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import mmf  # noqa: E402
from oracle import mmf_oracle as O  # noqa: E402


def main():
    out = {}
    # 1. the reference's five distinct weekly series, holdout 40 (02:341, 372-380, 484-494)
    df = mmf.synth.reference_weekly_demand(n_skus=1)
    skus = sorted(df["SKU"].unique())
    y = np.stack([df[df["SKU"] == s].sort_values("Date")["Demand"].to_numpy(np.float32) for s in skus])
    start = df["Date"].min()
    grid = O.calendar_grid(start, y.shape[1], "W-MON")
    X = O.design_matrix(grid, y.shape[1] - 40)
    pred, status = O.fit_forecast_packed(y, X, y.shape[1] - 40, 0, y.shape[1])
    out["ref_weekly_y"], out["ref_weekly_fitted"], out["ref_weekly_status"] = y, pred, status
    out["ref_weekly_start"] = np.array([np.datetime64(start, "D").astype(np.int64)])

    # 2. daily, one year (covid == 1 on the whole window -> aliased column), future 28, half the rows with gaps
    y, start = mmf.synth.daily_store_item_demand(48, 365, seed=11)
    rng = np.random.default_rng(5)
    holes = rng.random(y.shape) < 0.03
    holes[:24] = False
    y[holes] = np.nan
    grid = O.calendar_grid(start, 365 + 28, "D")
    X = O.design_matrix(grid, 365)
    pred, status = O.fit_forecast_packed(y, X, 365, 365, 28)
    out["daily365_y"], out["daily365_pred"], out["daily365_status"] = y, pred, status
    out["daily365_start"] = np.array([np.datetime64(start, "D").astype(np.int64)])

    # 3. daily, three years, future 28 and holdout 28
    y, start = mmf.synth.daily_store_item_demand(32, 1095, seed=12)
    grid = O.calendar_grid(start, 1095 + 28, "D")
    X = O.design_matrix(grid, 1095)
    out["daily1095_y"] = y
    out["daily1095_future"], _ = O.fit_forecast_packed(y, X, 1095, 1095, 28)
    Xh = O.design_matrix(grid[:1095], 1095 - 28)
    out["daily1095_holdout"], _ = O.fit_forecast_packed(y, Xh, 1095 - 28, 0, 1095)
    out["daily1095_start"] = np.array([np.datetime64(start, "D").astype(np.int64)])

    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
