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


def test_full_size_bf16_metrics_equal_fp32_to_4_decimals(big):
    """The benchmarked precision (bf16 tables + tcgen05 row transform) held to the north-star metric bar on the
    benchmarked workload: AUC / MRR / nDCG@5 / nDCG@10 over all 2.4 M impressions equal to 4 decimals
    (|delta| < 5e-5) to the fp32 path -- the reference's own arithmetic, golden-tested against the reference to
    1e-5 in test_gpu_api.py.  Also the parity report SURVEY section 7 asks for: the impressions whose dense ranks
    differ, with the score gaps that explain them."""
    import json
    import os

    import bench
    from news_recommendation_project_v2_b200 import ops
    from news_recommendation_project_v2_b200.engine import ScoringEngine
    from news_recommendation_project_v2_b200.modeling_utils import FinalAttention
    dev = big["scores"].device
    m32 = FinalAttention(DIM, HIDDEN, precision="fp32").eval()
    m32.load_state_dict(big["model"].state_dict())
    eng32 = ScoringEngine(big["table"], m32.to(dev), precision="fp32", device=dev)
    _, s32, r32 = eng32.score_device(big["hi"], big["ho"], big["ci"], big["co"], big["n_c"])
    s16, r16, co, n_imp = big["scores"], big["ranks"], big["co"], big["n_imp"]
    labels = bench.make_device_labels(co, 99, dev)
    _, sums16 = ops.mind_metrics(r16, labels, co, want_per_impression=False)
    _, sums32 = ops.mind_metrics(r32, labels, co, want_per_impression=False)
    assert int(sums16[4]) == int(sums32[4]) == n_imp  # every impression has both classes and finite ranks
    m16, mm32 = (sums16[:4] / sums16[4]).cpu().numpy(), (sums32[:4] / sums32[4]).cpu().numpy()
    err = (s16 - s32).abs()
    max_err, mean_err = float(err.max()), float(err.mean())
    # ---- rank-mismatch report -----------------------------------------------------------------------
    cnt = co[1:] - co[:-1]
    seg = torch.repeat_interleave(torch.arange(n_imp, device=dev), cnt)
    differs = torch.zeros(n_imp, dtype=torch.int32, device=dev).index_add_(0, seg, (r16 != r32).int()) > 0
    order = torch.argsort(seg.double() * 4.0 + s32.double())  # per impression ascending fp32 score
    ss, sg = s32[order].double(), seg[order]
    gap = ss[1:] - ss[:-1]
    same = sg[1:] == sg[:-1]
    min_gap = torch.full((n_imp,), 9.0, dtype=torch.float64, device=dev)
    min_gap.scatter_reduce_(0, sg[1:][same], gap[same], "amin")
    tol = 3e-3  # the bf16 score tolerance written in test_gpu_api.py / DESIGN.md section 5
    hard = int((differs & (min_gap > 2 * tol)).sum())
    mg = min_gap[differs]
    q = torch.quantile(mg[:5_000_000], torch.tensor([0.5, 0.9, 0.99, 1.0], dtype=torch.float64, device=dev))
    report = {
        "workload": "2.4 M cfg-4 impressions, N=161,013, d=1024: bf16 path vs fp32 path, same inputs",
        "metrics_bf16": [round(float(v), 6) for v in m16], "metrics_fp32": [round(float(v), 6) for v in mm32],
        "metric_abs_delta": [float(abs(a - b)) for a, b in zip(m16, mm32)],
        "score_abs_err_max": max_err, "score_abs_err_mean": mean_err,
        "impressions_with_rank_mismatch": int(differs.sum()), "impressions": n_imp,
        "min_adjacent_fp32_score_gap_of_mismatched_impressions": {
            "median": float(q[0]), "p90": float(q[1]), "p99": float(q[2]), "max": float(q[3])},
        "mismatches_with_every_gap_above_2x_tolerance": hard,
    }
    print("PARITY_REPORT " + json.dumps(report))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_fullsize_bf16_vs_fp32.json"), "w") as f:
            json.dump(report, f, indent=1)
    assert max_err <= tol, max_err
    assert hard == 0  # a rank can only move where two reference scores are closer than twice the score error
    np.testing.assert_allclose(m16, mm32, atol=5e-5, rtol=0)  # "equal to 4 decimals"
