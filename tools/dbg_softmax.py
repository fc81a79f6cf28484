import sys, math, torch
sys.path.insert(0, '/root/repo')
from news_recommendation_project_v2_b200 import ops
for (M,N,K,group,valid) in [(300,512,128,256,256),(700,1024,256,512,512)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * (4.0 / math.sqrt(K))).to(torch.bfloat16)
    logits = (a.double() @ w.double().T).reshape(M, N // group, group)
    want = torch.softmax(logits, dim=-1).reshape(M, N)
    y = ops.linear(a.cuda(), w.cuda(), None, 5, None, torch.bfloat16, group=group, group_valid=valid).float().cpu().double()
    rs = y.reshape(M, N//group, group).sum(-1)
    print((M,N,K,group), 'row-sum min/max', rs.min().item(), rs.max().item(), 'max abs err', (y-want).abs().max().item())
    bad = ((y-want).abs() > 5e-3).nonzero()
    print('  n bad', len(bad), 'first', bad[:5].tolist(), 'rows with bad', sorted(set(bad[:,0].tolist()))[:20])
    r = bad[0,0].item() if len(bad) else 0
    print('  ratio y/want row', r, (y[r,:8]/want[r,:8]).tolist())
