"""world_size-2 gloo worker for tests/test_host.py::test_all_gather_world2_gloo."""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mmf  # noqa: E402
from mmf import sharding as SH  # noqa: E402
from oracle import mmf_oracle as O  # noqa: E402


class OracleEngine:
    """Stands in for ForecastEngine on the CPU box (test infrastructure only)."""

    def __init__(self, X=None, t_fit=None):
        self.X, self.t_fit = X, t_fit

    def plan_calendar(self, start, t_len, freq="D", horizon=28, mode="future", design="trend_season_exog"):
        if mode == "holdout":
            self.t_fit, n_rows, ps, npred = t_len - horizon, t_len, 0, t_len
        else:
            self.t_fit, n_rows, ps, npred = t_len, t_len + horizon, t_len, horizon
        days = mmf.design.calendar_grid(start, n_rows, freq)
        self.X = mmf.design.design_matrix(days, self.t_fit, design)
        return days[ps:ps + npred], ps, npred

    def fit_forecast(self, y, pred_start, n_pred, out=None):
        pred, _ = O.fit_forecast_packed(y, self.X, self.t_fit, pred_start, n_pred)
        if out is None:
            return pred.astype(np.float32)
        out[...] = pred.astype(np.float32)
        return out


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, t, h = 37, 120, 14
    y, start = mmf.synth.daily_store_item_demand(n, t, seed=3)
    keys = pd.DataFrame({"store": ["s%d" % (i % 5) for i in range(n)], "item": ["i%d" % i for i in range(n)]})
    plan = SH.ShardPlan.build(SH.owner_of_keys(keys, world), world, rank)
    X = mmf.design.design_matrix(mmf.design.calendar_grid(start, t + h, "D"), t)
    eng = OracleEngine(X, t)
    table = SH.forecast_packed_sharded(y[plan.local_rows], plan, eng, t, h)
    want, _ = O.fit_forecast_packed(y, X, t, t, h)
    assert table.shape == (n, h)
    assert np.allclose(table.numpy(), want.astype(np.float32))
    # in-place flavour (packed benchmarks: contiguous blocks)
    plan2 = SH.ShardPlan.build(SH.owner_of_rows(n, world), world, rank)
    full = torch.zeros((world * plan2.per, h))
    lo = plan2.local_rows
    full[rank * plan2.per: rank * plan2.per + lo.size] = torch.from_numpy(want[lo].astype(np.float32))
    SH.all_gather_inplace(full, plan2)
    assert np.allclose(full[torch.as_tensor(plan2.gather_index())].numpy(), want.astype(np.float32))
    # DataFrame level: hash-shard the groups of a long frame, fit locally, gather the whole tuning_schema frame
    df = mmf.synth.reference_weekly_demand(n_skus=2)
    got = SH.forecast_groups_sharded(df, engine=OracleEngine())
    ref = mmf.forecast_groups(df, engine=OracleEngine())
    assert list(got.columns) == ["Product", "SKU", "Date", "Demand", "Demand_Fitted"] and len(got) == len(ref) == 10 * 157
    assert (got["SKU"].to_numpy() == ref["SKU"].to_numpy()).all()
    assert np.array_equal(got["Demand_Fitted"].to_numpy(), ref["Demand_Fitted"].to_numpy())
    mine = SH.forecast_groups_sharded(df, engine=OracleEngine(), gather=False)
    assert 0 < len(mine) < len(ref) and len(mine) % 157 == 0
    print(f"GLOO_OK rank {rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
