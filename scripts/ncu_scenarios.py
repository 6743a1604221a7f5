#!/usr/bin/env python
"""One profiled invocation of a kernel family that bench.py's timed region does not reach, bracketed by
cudaProfilerStart/Stop (run under `ncu --profile-from-start off`):
    select   : mmf_fit_select_forecast_f32 (select_kernel + fit + predict_tc_kernel), N x 1095, hold-out 28
    widen    : mmf_fit_forecast_int on a device-resident uint16/int16 buffer (widen_kernel)
    packer   : the device-side packer's kernels on G x T shuffled long-format rows
    ragged   : one ragged launch over C calendars
"""
import ctypes as C
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import mmf
    from mmf import _native as N
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    t, h = 1095, 28
    torch.cuda.set_device(0)
    prof = torch.cuda.profiler
    if what == "select":
        y, start = mmf.synth.daily_store_item_demand_torch(n, t, seed=3)
        eng = mmf.ForecastEngine()
        eng.plan_calendar(start, t, "D", h, "holdout")
        eng.fit_select_forecast(y, h, (1, 3, 9, 13, 16), 0, t)
        torch.cuda.synchronize()
        prof.start()
        eng.fit_select_forecast(y, h, (1, 3, 9, 13, 16), 0, t)
        torch.cuda.synchronize()
        prof.stop()
    elif what == "widen":
        y, start = mmf.synth.daily_store_item_demand_torch(n, t, seed=3)
        full = torch.zeros((n, 1096), dtype=torch.int16, device="cuda")     # 16-B row pitch: the vector path of the kernel
        yi = full[:, :t]
        yi.copy_(y.clamp(0, 32000).to(torch.int16))
        eng = mmf.ForecastEngine()
        _, ps, npred = eng.plan_calendar(start, t, "D", h, "future")
        eng.fit_forecast(yi, ps, npred)
        torch.cuda.synchronize()
        prof.start()
        eng.fit_forecast(yi, ps, npred)
        torch.cuda.synchronize()
        prof.stop()
    elif what == "ragged":
        cals = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        y, start = mmf.synth.daily_store_item_demand_torch(n, t, seed=3)
        eng = mmf.ForecastEngine()
        rows = np.linspace(0, n, cals + 1).astype(np.int64)
        eng.plan_calendars([np.datetime64(start, "D") - np.timedelta64(c, "D") for c in range(cals)], [t] * cals, "D", h)
        eng.fit_forecast_ragged(y, rows)
        torch.cuda.synchronize()
        prof.start()
        eng.fit_forecast_ragged(y, rows)
        torch.cuda.synchronize()
        prof.stop()
    elif what == "packer":
        G, T = n, 365
        nr = G * T
        y, start = mmf.synth.daily_store_item_demand_torch(G, T, seed=5)
        days = mmf.design.calendar_grid(start, T, "D").astype("datetime64[D]").astype(np.int32)
        perm = torch.randperm(nr, device="cuda")
        item = torch.arange(G, device="cuda", dtype=torch.int32).repeat_interleave(T)[perm]
        day = torch.as_tensor(days, device="cuda").repeat(G)[perm]
        val = y.contiguous().reshape(-1)[perm]
        eng = mmf.ForecastEngine()
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        lib, hd = eng._lib, eng._h
        hsh = torch.empty(nr, dtype=torch.int64, device="cuda")
        gid = torch.empty(nr, dtype=torch.int32, device="cuda")
        first = torch.empty(nr, dtype=torch.int32, device="cuda")
        out = torch.empty((G, (T + 3) & ~3), device="cuda")
        bad = torch.zeros(2, dtype=torch.int64, device="cuda")

        def device_pass():
            g = C.c_int32(0)
            N.check(lib.mmf_pack_hash_i32(hd, item.data_ptr(), nr, hsh.data_ptr(), 1))
            N.check(lib.mmf_pack_group_codes(hd, hsh.data_ptr(), nr, gid.data_ptr(), first.data_ptr(), C.byref(g)))
            bad.zero_()
            N.check(lib.mmf_pack_verify_i32(hd, item.data_ptr(), nr, gid.data_ptr(), first.data_ptr(), bad.data_ptr()))
            gmin = torch.empty(g.value, dtype=torch.int32, device="cuda")
            gmax = torch.empty(g.value, dtype=torch.int32, device="cuda")
            N.check(lib.mmf_pack_minmax(hd, gid.data_ptr(), day.data_ptr(), nr, g.value, gmin.data_ptr(), gmax.data_ptr()))
            rog = torch.arange(g.value, device="cuda", dtype=torch.int64)
            N.check(lib.mmf_pack_scatter_f32(hd, gid.data_ptr(), day.data_ptr(), val.data_ptr(), nr, rog.data_ptr(),
                                             gmin.data_ptr(), 1, out.data_ptr(), g.value, out.stride(0), T,
                                             bad[1:].data_ptr()))
        device_pass()
        torch.cuda.synchronize()
        prof.start()
        device_pass()
        torch.cuda.synchronize()
        prof.stop()
    else:
        raise SystemExit(f"unknown scenario {what}")


if __name__ == "__main__":
    main()
