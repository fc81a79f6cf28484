"""GPU parity at BASELINE size (configs[3]: 2.4 M MIND-large-shaped impressions, d=1024, bf16) through
size-independent properties, plus an oracle comparison on a sampled subset of the same run."""
import numpy as np
import pytest
import torch

from oracle import oracle
from news_recommendation_project_v2_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

N_ROWS, DIM, HIDDEN = 161_013, 1024, 4096


@pytest.fixture(scope="module")
def big():
    import bench
    from news_recommendation_project_v2_b200.engine import ScoringEngine
    from news_recommendation_project_v2_b200.modeling_utils import FinalAttention
    dev = torch.device("cuda", 0)
    model = FinalAttention(DIM, HIDDEN, precision="bf16").eval()
    model.load_state_dict(syn.make_final_attention_state_dict(DIM, HIDDEN, seed=1234))
    table = syn.make_table(N_ROWS, DIM, seed=1234)
    eng = ScoringEngine(table, model.to(dev), precision="bf16", device=dev)
    n_imp = 2_400_000
    hi, ho, ci, co, hl, cl, n_h, n_c = bench.make_device_impressions(n_imp, N_ROWS, 50, 1234, dev)
    _, scores, ranks = eng.score_device(hi, ho, ci, co, n_c)
    torch.cuda.synchronize()
    return dict(eng=eng, hi=hi, ho=ho, ci=ci, co=co, n_c=n_c, n_imp=n_imp, scores=scores, ranks=ranks, table=table,
                model=model)


def test_full_size_deterministic_and_shard_invariant(big):
    eng = big["eng"]
    _, s2, r2 = eng.score_device(big["hi"], big["ho"], big["ci"], big["co"], big["n_c"])
    assert torch.equal(s2, big["scores"]) and torch.equal(r2, big["ranks"])  # idempotent / deterministic
    # 8 impression shards (the multi-GPU partition) reproduce the single launch bit for bit
    s3 = torch.empty_like(s2)
    r3 = torch.empty_like(r2)
    n = big["n_imp"]
    for k in range(8):
        i0, i1 = n * k // 8, n * (k + 1) // 8
        eng.score_device(big["hi"], big["ho"][i0:i1 + 1], big["ci"], big["co"][i0:i1 + 1], big["n_c"],
                         out_scores=s3, out_ranks=r3)
    assert torch.equal(s3, big["scores"]) and torch.equal(r3, big["ranks"])


def test_full_size_rank_properties(big):
    """Dense-rank invariants over all 2.4 M impressions, checked with segment reductions on the GPU."""
    scores, ranks, co = big["scores"], big["ranks"].long(), big["co"]
    n_imp = big["n_imp"]
    seg = torch.repeat_interleave(torch.arange(n_imp, device=scores.device), co[1:] - co[:-1])
    assert torch.isfinite(scores).all() and scores.abs().max() <= 1.0 + 1e-3  # cosines
    cnt = co[1:] - co[:-1]
    rmin = torch.full((n_imp,), 1 << 30, device=scores.device).scatter_reduce(0, seg, ranks, "amin")
    rmax = torch.zeros(n_imp, dtype=torch.long, device=scores.device).scatter_reduce(0, seg, ranks, "amax")
    assert (rmin == 1).all() and (rmax <= cnt).all() and (rmax >= 1).all()
    # rank 1 <=> the impression's maximum score; a larger score never has a larger rank number
    smax = torch.full((n_imp,), -2.0, device=scores.device).scatter_reduce(0, seg, scores, "amax")
    assert torch.equal(ranks == 1, scores == smax[seg])
    # within an impression, sorting by rank sorts by score (descending): check adjacent pairs after a global sort
    key = seg.double() * 4096 + ranks.double()
    order = torch.argsort(key)
    s_sorted, seg_sorted, r_sorted = scores[order], seg[order], ranks[order]
    same = seg_sorted[1:] == seg_sorted[:-1]
    dr = r_sorted[1:] - r_sorted[:-1]
    assert ((dr == 0) | (dr == 1))[same].all()  # dense: consecutive integers
    assert (s_sorted[1:] < s_sorted[:-1])[same & (dr == 1)].all()
    assert (s_sorted[1:] == s_sorted[:-1])[same & (dr == 0)].all()


def test_full_size_sample_matches_oracle(big):
    """512 impressions sampled from the 2.4 M run against the fp64 oracle fed the same bf16 tables."""
    eng = big["eng"]
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(big["n_imp"], size=512, replace=False))
    ho, co = big["ho"].cpu().numpy(), big["co"].cpu().numpy()
    hi, ci = big["hi"].cpu().numpy(), big["ci"].cpu().numpy()
    X, E, T = eng.hist_x.cpu(), eng.hist_e.cpu(), eng.cand.cpu()
    bad_rank = 0
    for i in pick:
        r = torch.from_numpy(hi[ho[i]:ho[i + 1]]).long()
        u = (X[r].double() * E[r].double()).sum(0) / (E[r].double().sum(0) + 1e-10)
        c = torch.from_numpy(ci[co[i]:co[i + 1]]).long()
        want = oracle.cosine_scores(u[None], T, ci[co[i]:co[i + 1]], np.array([len(c)]), dtype=torch.float64).numpy()
        got = big["scores"][co[i]:co[i + 1]].cpu().numpy()
        np.testing.assert_allclose(got, want, atol=5e-6, rtol=0)
        gr = big["ranks"][co[i]:co[i + 1]].cpu().numpy()
        assert np.array_equal(gr, oracle.dense_rank_desc(got).astype(np.int32))  # bit-exact on own score bits
        s = np.sort(want)
        if len(s) < 2 or np.min(np.diff(s)) > 1e-5:
            bad_rank += int(not np.array_equal(gr, oracle.dense_rank_desc(want).astype(np.int32)))
    assert bad_rank == 0


def test_candidate_permutation_equivariance(big):
    """Permuting the candidates of an impression permutes scores and ranks the same way (bit-exact)."""
    eng = big["eng"]
    n = 50_000
    co = big["co"][: n + 1]
    n_c = int(co[-1])
    ci = big["ci"][:n_c].clone()
    seg = torch.repeat_interleave(torch.arange(n, device=ci.device), co[1:] - co[:-1])
    g = torch.Generator(device=ci.device).manual_seed(3)
    perm = torch.argsort(seg.double() + torch.rand(n_c, generator=g, device=ci.device, dtype=torch.float64) * 0.5)
    _, s, r = eng.score_device(big["hi"], big["ho"][: n + 1], ci[perm].contiguous(), co, n_c)
    assert torch.equal(s, big["scores"][:n_c][perm]) and torch.equal(r, big["ranks"][:n_c][perm])
