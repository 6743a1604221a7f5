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

    def __init__(self, X, t_fit):
        self.X, self.t_fit = X, t_fit

    def fit_forecast(self, y, pred_start, n_pred, out=None):
        pred, _ = O.fit_forecast_packed(y, self.X, self.t_fit, pred_start, n_pred)
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
    print(f"GLOO_OK rank {rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
