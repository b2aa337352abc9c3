"""Run every GPU parity check without stopping at the first failure; print a table and write
gpurun_out/report.json.  Usage (on the B200 box):  python scripts/gpu_report.py [ops] [unet] [full]"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    what = set(sys.argv[1:]) or {"ops", "unet"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    report = dict(ops=[], unet=[])
    if "ops" in what:
        import parity_checks as pc
        for thunk in pc.all_op_checks():
            t0 = time.time()
            try:
                r = thunk()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                r = dict(name="<exception>", ok=False, error=f"{type(e).__name__}: {e}",
                         tb=traceback.format_exc()[-600:])
            r["sec"] = round(time.time() - t0, 3)
            report["ops"].append(r)
            print(("PASS " if r.get("ok") else "FAIL ") + json.dumps({k: (round(v, 6) if isinstance(v, float) else v)
                                                                    for k, v in r.items() if k != "tb"}), flush=True)
    if "unet" in what:
        import unet_checks as uc
        from rcdms_b200.unet_spec import full_config, tiny_config
        cases = [("tiny", tiny_config(), (2, 5, 8, 8, 7), 981, torch.float16, True),
                 ("tiny", tiny_config(), (2, 5, 8, 8, 7), 981, torch.float16, False),
                 ("tiny", tiny_config(), (2, 5, 16, 16, 85), 501, torch.float16, False),
                 ("tiny", tiny_config(), (2, 5, 16, 16, 85), 501, torch.bfloat16, False),
                 ("tiny", tiny_config(), (4, 5, 32, 32, 91), 21, torch.float16, False)]
        if "full" in what:
            cases.append(("full", full_config(), (2, 5, 8, 8, 85), 981, torch.float16, False))
            cases.append(("full", full_config(), (2, 5, 32, 32, 85), 501, torch.float16, False))
        for name, cfg, shape, t, dt, simple in cases:
            t0 = time.time()
            try:
                res = uc.run_case(cfg, shape, t, dt, simple=simple, taps=True)
                r = dict(name=f"unet {name} {shape} t{t} {dt} simple{int(simple)}", **res["stats"],
                         floor_max=res["floor"]["max_abs"], floor_mean=res["floor"]["mean_abs"])
                r["ok"] = r["finite"] and r["max_abs"] <= max(3 * r["floor_max"], 5e-3)
                r["taps"] = {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                             for k, v in res["tap_stats"].items()}
                del res
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                r = dict(name=f"unet {name} {shape}", ok=False, error=f"{type(e).__name__}: {e}",
                         tb=traceback.format_exc()[-1500:])
            r["sec"] = round(time.time() - t0, 2)
            report["unet"].append(r)
            print(("PASS " if r.get("ok") else "FAIL ") + json.dumps({k: v for k, v in r.items() if k != "taps"}),
                  flush=True)
            for k, v in (r.get("taps") or {}).items():
                print("     tap", k, v, flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    n_fail = sum(not r.get("ok") for sec in report.values() for r in sec)
    print(f"SUMMARY: {sum(len(v) for v in report.values())} checks, {n_fail} failed")


if __name__ == "__main__":
    main()
