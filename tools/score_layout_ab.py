"""A/B of the history-table layout for FinalAttention pooling: X and E as two [N, d] tables vs one [N, 2d]
table with X and E of a row adjacent (one 4 KB read per history slot instead of two 2 KB reads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from news_recommendation_project_v2_b200 import _lib, ops
from news_recommendation_project_v2_b200._lib import check, dtype_code, load, ptr, stream_ptr

dev = torch.device("cuda", 0)
d, n_rows = 1024, 161_013
n_imp = int(os.environ.get("N_IMP", "2400000"))
T = torch.randn(n_rows, d, device=dev, dtype=torch.bfloat16)
XE = torch.randn(n_rows, 2 * d, device=dev, dtype=torch.bfloat16)
XE[:, d:] = torch.rand(n_rows, d, device=dev).to(torch.bfloat16) + 0.5
X, E = XE[:, :d].contiguous(), XE[:, d:].contiguous()
hi, ho, ci, co, _, _, n_h, n_c = bench.make_device_impressions(n_imp, n_rows, 50, 7, dev)
scores = torch.empty(n_c, dtype=torch.float32, device=dev)
ranks = torch.empty(n_c, dtype=torch.int32, device=dev)
flag = ops.new_err_flag(dev)
lib = load()
es = 2


def run(px, pe, stride):
    check(lib.nrb_score_rank(0, dtype_code(torch.bfloat16), d, n_rows, px, pe, stride, ptr(T), d, None, 1.0, ptr(hi),
                             ptr(ho), ptr(ci), ptr(co), n_imp, None, ptr(scores), ptr(ranks), ptr(flag), stream_ptr()),
          "nrb_score_rank")


def timed(px, pe, stride):
    for _ in range(2):
        run(px, pe, stride)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(3):
        run(px, pe, stride)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / 3


b = (2 * n_h + n_c) * d * es + 4 * (n_h + n_c) + 8 * n_c
out = []
for rep in range(2):
    ms = timed(ptr(X), ptr(E), d)
    s_sep = scores.clone()
    out.append("separate=%.0f" % (b / ms / 1e6))
    ms = timed(XE.data_ptr(), XE.data_ptr() + d * es, 2 * d)
    out.append("adjacent=%.0f" % (b / ms / 1e6))
    assert torch.equal(s_sep, scores)
print(" ".join(out), flush=True)
