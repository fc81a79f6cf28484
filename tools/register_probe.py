"""Does cudaHostRegister accept the mapped token section of a packed token file on this box?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from news_recommendation_project_v2_b200.token_store import PackedTokenFile, write_packed_tokens  # noqa: E402

torch.cuda.init()
for base in ("/dev/shm", "/tmp"):
    path = os.path.join(base, "nrb200_probe.nrbtok")
    g = torch.Generator().manual_seed(0)
    items = [torch.randn(64, 768, generator=g).to(torch.bfloat16) for _ in range(4096)]
    write_packed_tokens(path, items, 768)
    tf = PackedTokenFile(path)
    t = time.perf_counter()
    ok = tf.register()
    print(base, "register:", ok, tf.register_error, "%.1f ms for %.1f MB" % ((time.perf_counter() - t) * 1e3, tf.tokens_raw.nbytes / 1e6))
    if ok:
        src = torch.from_numpy(tf.tokens_raw[:])
        dst = torch.empty(src.shape, dtype=src.dtype, device="cuda")
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        print("   DMA from the mapping: %.1f GB/s" % (5 * src.numel() * 2 / (time.perf_counter() - t) / 1e9))
        tf.unregister()
    os.remove(path)
