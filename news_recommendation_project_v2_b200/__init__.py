"""B200-native embedding-and-scoring hot path for news_recommendation_project_v2.

Drop-in host API (same names / signatures / state_dict keys as the reference's
`news_rec_utils`) over hand-written sm_100a CUDA kernels behind a C-ABI shared
library (`include/nrb200.h`).  There is no CPU fallback: every compute entry
point raises if `libnrb200.so` is missing or no B200 is visible.
"""
__version__ = "0.1.0"
