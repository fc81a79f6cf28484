"""Where the end-to-end step spends its time: CUDA-event timeline of one warm step (transform + score_host)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from news_recommendation_project_v2_b200 import synthetic as syn  # noqa: E402
from news_recommendation_project_v2_b200.engine import ScoringEngine, _mark  # noqa: E402
from news_recommendation_project_v2_b200.modeling_utils import FinalAttention  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = FinalAttention(bench.DIM, bench.HIDDEN, precision="bf16").eval()
model.load_state_dict(syn.make_final_attention_state_dict(bench.DIM, bench.HIDDEN, seed=1234))
model.to(dev)
table = syn.make_table(bench.N_ROWS, bench.DIM, seed=1234).pin_memory()
n_imp = 2_400_000
hist_idx, h_off, cand_idx, c_off, hist_len, cand_len, n_h, n_c = bench.make_device_impressions(n_imp, bench.N_ROWS, bench.H_MAX, 1234, dev)
pin = lambda x: x.cpu().pin_memory()
hi, ci, ho, co = pin(hist_idx), pin(cand_idx), pin(h_off), pin(c_off)
sc = torch.empty(n_c, dtype=torch.float32).pin_memory()
rk = torch.empty(n_c, dtype=torch.int16).pin_memory()
eng = ScoringEngine(table, model, precision="bf16", device=dev)
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else bench.E2E_CHUNKS
for rep in range(4):
    torch.cuda.synchronize()
    eng._trace = [] if rep == 3 else None
    cur = torch.cuda.current_stream()
    t0 = _mark(cur)
    eng.prepare_user_encoder(eng.cand)
    t1 = _mark(cur)
    eng.score_host(hi, ho, ci, co, scores_out=sc, ranks_out=rk, n_chunks=chunks)
    t2 = _mark(cur)
    torch.cuda.synchronize()
print("transform %.3f ms, score_host %.3f ms (total %.3f)" % (t0.elapsed_time(t1), t1.elapsed_time(t2), t0.elapsed_time(t2)))
prev = None
for label, ev in eng._trace:
    print("  %-28s at %8.3f ms" % (label, t0.elapsed_time(ev)))
