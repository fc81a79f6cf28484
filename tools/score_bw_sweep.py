"""Microbenchmark: achieved algorithmic GB/s of nrb_score_rank vs table size / pooling mode."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from news_recommendation_project_v2_b200 import ops

dev = torch.device("cuda", 0)
d = 1024
res = []
ROWS = [int(r) for r in os.environ.get("SWEEP_ROWS", "161013,1000000,2500000,10000000").split(",")]
for n_rows in ROWS:
    T = torch.randn(n_rows, d, device=dev, dtype=torch.bfloat16)
    X = torch.randn(n_rows, d, device=dev, dtype=torch.bfloat16)
    E = torch.rand(n_rows, d, device=dev, dtype=torch.bfloat16) + 0.5
    for mode, hmax in ((0, 50), (1, 50), (1, 200)):
        n_imp = 400_000
        hi, ho, ci, co, _, _, n_h, n_c = bench.make_device_impressions(n_imp, n_rows, hmax, 7, dev)
        scores = torch.empty(n_c, dtype=torch.float32, device=dev)
        ranks = torch.empty(n_c, dtype=torch.int32, device=dev)
        flag = ops.new_err_flag(dev)
        run = lambda: ops.score_rank(mode, X, E if mode == 0 else None, T, hi, ho, ci, co, n_c, err_flag=flag,
                                     out_scores=scores, out_ranks=ranks)
        for _ in range(2):
            run()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(3):
            run()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        r = 2 if mode == 0 else 1
        b = (r * n_h + n_c) * d * 2 + 4 * (n_h + n_c) + 8 * n_c
        res.append(dict(n_rows=n_rows, mode=mode, hmax=hmax, ms=round(ms, 3), gbs=round(b / ms / 1e6, 1)))
        print(res[-1], flush=True) if "SWEEP_QUIET" not in os.environ else None
    del T, X, E
if "SWEEP_QUIET" in os.environ:
    print(" ".join("m%d/h%d=%.0f" % (r["mode"], r["hmax"], r["gbs"]) for r in res), flush=True)
else:
    json.dump(res, open("gpurun_out/score_bw_sweep.json", "w"))
