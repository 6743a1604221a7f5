import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mmf
for (n, t, mode) in [(20011, 365, "holdout"), (20011, 365 + 28 - 1, "future"), (20011, 365, "future"), (3000, 400, "future")]:
    h = 28
    for nanmode in ("gaps", "clean"):
        y2, start = mmf.synth.daily_store_item_demand_torch(n, t, seed=77 + n, nan_frac=0.02 if nanmode == "gaps" else 0.0)
        got = {}
        for variant in (1, 3):
            eng = mmf.ForecastEngine(kernel="tc", tc_variant=variant)
            res = mmf.forecast_packed(y2, start, "D", h, mode, engine=eng, want_status=True)
            torch.cuda.synchronize()
            got[variant] = (res["pred"].cpu().numpy(), res["status"].cpu().numpy())
            eng.close()
        a, b = got[1][0], got[3][0]
        neq = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        rows = np.where(neq.any(axis=1))[0]
        print(n, t, mode, nanmode, "rows differing", len(rows), rows[:20], "status", got[1][1][rows[:10]])
        if len(rows):
            r = rows[0]
            print("  maxabs", np.nanmax(np.abs(a[rows] - b[rows])), "row", r, a[r][:4], b[r][:4])
            R = ((n + 147) // 148 + 7) // 8 * 8
            print("  R", R, "row%R", rows[:20] % R)
