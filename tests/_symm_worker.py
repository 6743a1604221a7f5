"""torchrun worker of tests/test_gpu_configs.py::test_symmetric_table_fused_gather_matches_oracle (one process per
GPU, NCCL).  Every rank fits its shard with ``SymmetricTable.fit_into`` -- the fit kernel's epilogue stores each
forecast tile into every rank's copy of the table -- and compares ITS WHOLE TABLE with the float64 oracle's forecasts
of all shards, and with a plain NCCL all_gather of the same rows."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mmf  # noqa: E402
from mmf.sharding import SymmetricTable  # noqa: E402
from oracle import mmf_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="p2p")
    ap.add_argument("--out", required=True)
    ap.add_argument("--per", type=int, default=20_000)       # rows per rank: 157 tiles, ragged last tile
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per, t, h = a.per + 37, 400, 28
    # every rank generates ALL shards (same seeds) so that it can check the whole table against the oracle
    shards = []
    for r in range(world):
        y, start = mmf.synth.daily_store_item_demand(per, t, seed=900 + r)
        y[5 + r, 10:40] = np.nan                              # in-stream gap path (solve_rows writes the peers too)
        y[6 + r, :9] = np.nan                                 # first values missing: general pass
        y[7 + r, :] = np.nan                                  # empty row: NaN everywhere, status 1
        shards.append(y)
    grid = O.calendar_grid(start, t + h, "D")
    X = O.design_matrix(grid, t)
    want = np.concatenate([O.fit_forecast_packed_c(y, X, t, t, h)[0] for y in shards])
    wst = O.fit_forecast_packed_c(shards[rank], X, t, t, h)[1]
    tol = 5e-6 * float(np.nanmax(np.abs(np.concatenate(shards)))) + 1e-3
    skipped = None
    try:
        sym = SymmetricTable(per, h, dev, mode=a.mode)
        ok = torch.ones(1, device=dev)
    except Exception as exc:                                  # no NVLink symmetric memory / NVLS on this box
        sym, ok, skipped = None, torch.zeros(1, device=dev), repr(exc)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res = None
    if float(ok[0]) >= 1.0:
        eng = mmf.ForecastEngine(device=local)
        _, ps, npred = eng.plan_calendar(start, t, "D", h, "future")
        yd = mmf.device_packed(shards[rank], device=dev)
        status = torch.empty(per, dtype=torch.int32, device=dev)
        for _ in range(2):                                    # twice: the second call reuses zeroed work counters
            sym.table.fill_(-1.0)
            sym.barrier()
            sym.fit_into(eng, yd, ps, npred, status=status)
            sym.barrier()
        torch.cuda.synchronize()
        table = sym.table.cpu().numpy()
        loc = torch.empty((per, h), device=dev)
        eng.fit_forecast(yd, ps, npred, out=loc)
        ref = torch.empty((world * per, h), device=dev)
        dist.all_gather_into_tensor(ref, loc)
        torch.cuda.synchronize()
        both_nan = np.isnan(table) & np.isnan(want)
        err = float(np.abs(np.where(both_nan, 0.0, table - want)).max())
        res = {"rank": rank, "max_err_vs_oracle": err if np.isfinite(err) else 1e30,
               "status_equal": bool(np.array_equal(status.cpu().numpy(), wst)),
               "equals_nccl_gather": bool(np.array_equal(table, ref.cpu().numpy(), equal_nan=True))}
        eng.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump({"world": world, "mode": a.mode, "tol": tol, "ranks": gathered,
                       **({"skipped": f"symmetric memory unavailable: {skipped}"} if gathered[0] is None else {})}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
