"""Generate golden fixtures by EXECUTING the reference's own pure-pandas code.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_reference_fixtures.py
It extracts, by line range, the functions the reference can run without Spark /
statsmodels / hyperopt, executes them unmodified, and stores their outputs in
``tests/golden/reference_fixtures.npz``:
  * ``add_exo_variables``         group_apply/02_Fine_Grained_Demand_Forecasting.py:343-358
  * ``split_train_score_data``    group_apply/02_Fine_Grained_Demand_Forecasting.py:372-380
  * the demand calendar + factors group_apply/_resources/01-data-generator.py:57-62,135-181
No reference source is copied into the repo: only the numeric outputs are stored.
"""
import datetime
import datetime as dt
import os

import numpy as np
import pandas as pd

REF = "/root/reference/group_apply"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(path, lo, hi):
    with open(path) as f:
        src = f.read().split("\n")
    return "\n".join(src[lo - 1:hi])


def main():
    nb = os.path.join(REF, "02_Fine_Grained_Demand_Forecasting.py")
    gen = os.path.join(REF, "_resources", "01-data-generator.py")

    # ---- 02:341-358 + 372-380 ------------------------------------------------------------
    ns = {"pd": pd, "np": np, "dt": dt}
    exec(lines(nb, 341, 358), ns)
    exec(lines(nb, 372, 380), ns)
    add_exo_variables = ns["add_exo_variables"]
    split = ns["split_train_score_data"]
    assert ns["FORECAST_HORIZON"] == 40

    out = {}
    # (a) the reference's weekly calendar, (b) a daily range covering 3 years incl. a week-53 year
    weekly = [datetime.date(2021, 7, 19) - datetime.timedelta(weeks=k) for k in range(156, -1, -1)]
    daily = [datetime.date(2018, 7, 1) + datetime.timedelta(days=k) for k in range(1115)]
    for name, dates in (("weekly", weekly), ("daily", daily)):
        pdf = pd.DataFrame({"Date": dates, "Product": "P", "SKU": "S", "Demand": np.float32(1.0)})
        enriched = add_exo_variables(pdf)
        assert list(enriched.columns) == ["Date", "Product", "SKU", "Demand", "covid", "christmas", "new_year"]
        out[f"exo_{name}_days"] = np.array([np.datetime64(d, "D") for d in dates]).astype(np.int64)
        out[f"exo_{name}"] = enriched[["covid", "christmas", "new_year"]].to_numpy(dtype=np.float64)

    cases = []
    for n, h in ((157, 40), (365, 28), (1095, 28), (41, 40)):
        data = pd.DataFrame({"v": np.arange(n)})
        train, score = split(data, h)
        cases.append([n, h, len(train), len(score), int(score["v"].iloc[0]), int(train["v"].iloc[-1])])
    out["split_cases"] = np.array(cases, dtype=np.int64)

    # ---- 01-data-generator.py:57-62, 135-181 ------------------------------------------------
    from dateutil import rrule
    from dateutil.relativedelta import relativedelta

    ns2 = {"datetime": datetime, "np": np, "pd": pd, "rrule": rrule, "relativedelta": relativedelta,
           "display": lambda *_a, **_k: None}
    exec(lines(gen, 57, 62), ns2)
    exec(lines(gen, 135, 181), ns2)
    dr = ns2["date_range"]
    out["gen_days"] = np.array([np.datetime64(d, "D") for d in dr["Date"]]).astype(np.int64)
    out["gen_helper"] = dr["Corona_Breakpoint_Helper"].to_numpy(dtype=np.int64)
    out["gen_corona_factor"] = dr["Corona_Factor"].to_numpy(dtype=np.float64)
    out["gen_week"] = np.asarray(dr["Week"], dtype=np.int64)
    out["gen_factor_xmas"] = dr["Factor_XMas"].to_numpy(dtype=np.float64)

    np.savez_compressed(os.path.join(HERE, "reference_fixtures.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("covid/christmas/new_year ones on the weekly calendar:", out["exo_weekly"].sum(axis=0))


if __name__ == "__main__":
    main()
