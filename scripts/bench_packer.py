#!/usr/bin/env python
"""Device-side packer micro-benchmark: G groups x T days of long-format rows (shuffled) -> y[G,T].
Prints one JSON line: rows/s for the kernels alone (inputs resident on the GPU, CUDA events) and end to end from
host Arrow buffers, next to the pandas packer on the host cores."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pyarrow as pa
    import torch

    import mmf
    from mmf import _native as N
    from mmf.packer import pack_table_device

    G, T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, int(sys.argv[2]) if len(sys.argv) > 2 else 1095
    n = G * T
    y, start = mmf.synth.daily_store_item_demand_torch(G, T, seed=5)
    days = mmf.design.calendar_grid(start, T, "D").astype("datetime64[D]").astype(np.int32)
    perm = torch.randperm(n, device="cuda")
    item = torch.arange(G, device="cuda", dtype=torch.int32).repeat_interleave(T)[perm]
    day = torch.as_tensor(days, device="cuda").repeat(G)[perm]
    val = y.contiguous().reshape(-1)[perm]
    eng = mmf.ForecastEngine()
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    lib, h = eng._lib, eng._h
    hsh = torch.empty(n, dtype=torch.int64, device="cuda")
    gid = torch.empty(n, dtype=torch.int32, device="cuda")
    first = torch.empty(n, dtype=torch.int32, device="cuda")
    out = torch.empty((G, (T + 3) & ~3), device="cuda")
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")

    def device_pass():
        g = C.c_int32(0)
        N.check(lib.mmf_pack_hash_i32(h, item.data_ptr(), n, hsh.data_ptr(), 1))
        N.check(lib.mmf_pack_group_codes(h, hsh.data_ptr(), n, gid.data_ptr(), first.data_ptr(), C.byref(g)))
        bad.zero_()
        N.check(lib.mmf_pack_verify_i32(h, item.data_ptr(), n, gid.data_ptr(), first.data_ptr(), bad.data_ptr()))
        gmin = torch.empty(g.value, dtype=torch.int32, device="cuda")
        gmax = torch.empty(g.value, dtype=torch.int32, device="cuda")
        N.check(lib.mmf_pack_minmax(h, gid.data_ptr(), day.data_ptr(), n, g.value, gmin.data_ptr(), gmax.data_ptr()))
        rog = torch.arange(g.value, device="cuda", dtype=torch.int64)
        N.check(lib.mmf_pack_scatter_f32(h, gid.data_ptr(), day.data_ptr(), val.data_ptr(), n, rog.data_ptr(),
                                         gmin.data_ptr(), 1, out.data_ptr(), g.value, out.stride(0), T, None))
        return g.value

    for _ in range(2):
        device_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        g = device_pass()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / K
    assert g == G

    table = pa.table({"item": pa.array(item.cpu().numpy()), "date": pa.array(day.cpu().numpy(), type=pa.int32()).cast(pa.date32()),
                      "sales": pa.array(val.cpu().numpy())})
    t0 = time.perf_counter()
    (b,) = pack_table_device(table, keys=("item",), date_col="date", value_col="sales", freq="D", engine=eng)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    chk = torch.empty_like(b.y)
    chk[torch.as_tensor(b.key_frame["item"].to_numpy().astype(np.int64), device="cuda")] = b.y
    ok = bool(torch.equal(chk, y))

    host_s = None
    if n <= 40_000_000:
        df = table.to_pandas()
        t0 = time.perf_counter()
        mmf.pack_groups(df, keys=("item",), date_col="date", value_col="sales", freq="D", pinned=False)
        host_s = time.perf_counter() - t0
    print(json.dumps({"groups": G, "t": T, "rows": n, "device_kernels_ms": dev_ms, "device_rows_per_s": n / dev_ms * 1e3,
                      "device_GBps_algorithmic(12B_in+4B_out_per_row)": n * 16 / dev_ms / 1e6,
                      "e2e_from_host_arrow_s": e2e_s, "e2e_rows_per_s": n / e2e_s, "matches_direct_array": ok,
                      "pandas_packer_s": host_s, "pandas_rows_per_s": (n / host_s) if host_s else None,
                      "host_cores": len(os.sched_getaffinity(0))}))


if __name__ == "__main__":
    main()
