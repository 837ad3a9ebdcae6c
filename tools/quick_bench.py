"""Developer timing probe (not the contract bench): parity + per-stage times of the named workloads, one frame at a time.

    python tools/quick_bench.py [binned|direct] [c2_grid c4_views c1_sponza ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    args = sys.argv[1:]
    mode = args.pop(0) if args and args[0] in ("binned", "direct") else "binned"
    names = args or ["c2_grid", "c4_views", "c1_sponza", "c1_knot", "c3_knot"]
    golden = bench.load_golden()
    peak, _ = bench.peak_hbm()
    for name in names:
        out = bench.run_config(name, 0, mode, golden, peak)
        for v in out["views"]:
            st = {k: s["us"] for k, s in v["stages"].items()}
            par = v["parity"]["visbuffer"] + ("" if not isinstance(v["parity"]["colour"], dict) else f"/colour err {v['parity']['colour']['max_abs_err']}")
            print(f"{name} view={v['view']} [{mode}] parity={par}  {v['ms_per_frame'] * 1e3:.1f} us/frame (back to back {v['ms_per_frame_back_to_back'] * 1e3:.1f})  "
                  f"stages(us)={st}  rasterized {v['rasterized_Mtri_s'] / 1e3:.2f} Gtri/s  stats={v['draw_stats']}", flush=True)
