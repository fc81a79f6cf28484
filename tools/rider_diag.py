"""Diagnosis of the rider path: which rows / items differ between riding and stand-alone passes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from news_recommendation_project_v2_b200 import config, synthetic as syn  # noqa: E402
from news_recommendation_project_v2_b200.latent_attention import LatentAttentionModel  # noqa: E402

dim, L, heads, dh = 256, 40, 4, 64
m = LatentAttentionModel(dim=dim, num_latents=L, precision="bf16", heads=heads, dim_head=dh).eval()
m.load_state_dict(syn.make_latent_state_dict(dim, L, seed=11, heads=heads, dim_head=dh))
config.LATENT_MAX_TOKENS = 1 << 20
os.environ["NRB200_RIDER_MIN_TOKENS"] = "64"
for T in (1792, 1800, 2048, 1000, 5000):
    g = torch.Generator().manual_seed(T)
    tok = torch.randn(T, dim, generator=g).to(torch.bfloat16).cuda()
    lens = torch.randint(1, 30, (400,), generator=g)
    off = torch.zeros(401, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    n_items = int((off <= T).sum()) - 1
    off = off[:n_items + 1].clone()
    off[-1] = T
    res = {}
    for r in ("0", "1"):
        os.environ["NRB200_RIDERS"] = r
        res[r] = (m.forward_packed(tok, off).cpu(), m(tok.view(T, 1, dim), None).cpu().view(T, dim))
    bad_items = (res["0"][0] != res["1"][0]).any(1).nonzero().flatten().tolist()
    bad_rows = (res["0"][1] != res["1"][1]).any(1).nonzero().flatten().tolist()
    print(f"T={T}: items differing {bad_items[:20]} ({len(bad_items)}), rows differing (unpooled) {bad_rows[:20]} ({len(bad_rows)})")
    if bad_items:
        i = bad_items[0]
        print("   item", i, "rows", int(off[i]), int(off[i + 1]), "max abs diff", float((res['0'][0][i] - res['1'][0][i]).abs().max()))
    if bad_rows:
        rr = bad_rows[0]
        print("   row", rr, "max abs diff", float((res['0'][1][rr] - res['1'][1][rr]).abs().max()))
