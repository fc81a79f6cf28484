// Behaviour log -> CSR index builder (host code; the step in front of the hot path).
//
// Replaces data_utils.py:168-232 split_impressions_and_history: a per-row Python loop with a dict that
// assigns table row ids in first-appearance order over history-then-impression tokens and emits the flat
// int32 index arrays + length lists the scoring kernels consume.  Same semantics:
//   * a row with an empty history contributes nothing to the history arrays (no length entry);
//   * labels are present iff the FIRST impression row contains '-' (data_utils.py:172);
//   * impression tokens are "NEWSID-label" (split on '-') when labels are present.
// Inputs are two '\n'-separated UTF-8 buffers (one line per behaviour row).
#include "common.cuh"

#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace nrb {

struct CsrBuild {
  std::vector<std::string> news;
  std::vector<int32_t> hist_idx, hist_len, hist_owner, cand_idx, cand_len, cand_owner;
  std::vector<int8_t> labels;
  int label_present = 0;
  int64_t n_rows = 0;
};

}  // namespace nrb

using namespace nrb;

extern "C" void* nrb_csr_build(const char* impressions, int64_t imp_bytes, const char* history, int64_t hist_bytes,
                               int64_t n_rows) {
  if (!impressions || !history || n_rows <= 0) {
    set_error("nrb_csr_build: No Impressions given");
    return nullptr;
  }
  auto* b = new CsrBuild();
  b->n_rows = n_rows;
  b->news.reserve(1 << 16);
  // ids are interned as std::string in `news`; the map keys must stay valid when `news` reallocates,
  // so the map owns copies through a deque-like arena of strings
  std::unordered_map<std::string, int32_t> pos;
  pos.reserve(1 << 17);
  auto intern = [&](std::string_view tok) -> int32_t {
    auto it = pos.find(std::string(tok));
    if (it != pos.end()) return it->second;
    const int32_t id = (int32_t)b->news.size();
    b->news.emplace_back(tok);
    pos.emplace(std::string(tok), id);
    return id;
  };
  auto next_line = [](const char* base, int64_t n, int64_t& cur) -> std::string_view {
    if (cur > n) return std::string_view();
    int64_t e = cur;
    while (e < n && base[e] != '\n') ++e;
    std::string_view v(base + cur, (size_t)(e - cur));
    cur = e + 1;
    return v;
  };
  auto for_tokens = [](std::string_view line, auto&& fn) {
    size_t i = 0, n = line.size();
    while (i < n) {
      while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
      size_t j = i;
      while (j < n && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') ++j;
      if (j > i) fn(line.substr(i, j - i));
      i = j;
    }
  };
  int64_t ci = 0, ch = 0;
  {
    int64_t probe = 0;
    std::string_view first = next_line(impressions, imp_bytes, probe);
    b->label_present = first.find('-') != std::string_view::npos ? 1 : 0;
  }
  int32_t hist_rows = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    std::string_view imp = next_line(impressions, imp_bytes, ci);
    std::string_view hist = next_line(history, hist_bytes, ch);
    if (!hist.empty()) {
      int32_t cnt = 0;
      for_tokens(hist, [&](std::string_view t) {
        b->hist_idx.push_back(intern(t));
        b->hist_owner.push_back(hist_rows);
        ++cnt;
      });
      b->hist_len.push_back(cnt);
      ++hist_rows;
    }
    int32_t cnt = 0;
    bool bad = false;
    for_tokens(imp, [&](std::string_view t) {
      if (b->label_present) {
        const size_t d = t.find('-');
        if (d == std::string_view::npos) {
          bad = true;
          return;
        }
        b->labels.push_back((int8_t)atoi(std::string(t.substr(d + 1)).c_str()));
        t = t.substr(0, d);
      }
      b->cand_idx.push_back(intern(t));
      b->cand_owner.push_back((int32_t)r);
      ++cnt;
    });
    if (bad) {
      set_error("nrb_csr_build: impression token without '-label' in row %lld", (long long)r);
      delete b;
      return nullptr;
    }
    b->cand_len.push_back(cnt);
  }
  return b;
}

// sizes[0..4] = {n_news, sum_history, n_history_rows, sum_candidates, label_present}
extern "C" int nrb_csr_sizes(void* handle, int64_t* sizes) {
  NRB_REQUIRE(handle && sizes, "nrb_csr_sizes: null");
  auto* b = (CsrBuild*)handle;
  sizes[0] = (int64_t)b->news.size();
  sizes[1] = (int64_t)b->hist_idx.size();
  sizes[2] = (int64_t)b->hist_len.size();
  sizes[3] = (int64_t)b->cand_idx.size();
  sizes[4] = b->label_present;
  return NRB_OK;
}

// copies into caller-allocated HOST arrays (any may be NULL to skip)
extern "C" int nrb_csr_export(void* handle, int32_t* hist_idx, int32_t* hist_owner, int32_t* hist_len,
                              int32_t* cand_idx, int32_t* cand_owner, int32_t* cand_len, int8_t* labels) {
  NRB_REQUIRE(handle, "nrb_csr_export: null");
  auto* b = (CsrBuild*)handle;
  auto cp = [](auto* dst, const auto& v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  cp(hist_idx, b->hist_idx);
  cp(hist_owner, b->hist_owner);
  cp(hist_len, b->hist_len);
  cp(cand_idx, b->cand_idx);
  cp(cand_owner, b->cand_owner);
  cp(cand_len, b->cand_len);
  cp(labels, b->labels);
  return NRB_OK;
}

// news ids joined with '\n' into `out` (capacity `cap`); returns the number of bytes needed
extern "C" int64_t nrb_csr_news_ids(void* handle, char* out, int64_t cap) {
  if (!handle) return -1;
  auto* b = (CsrBuild*)handle;
  int64_t need = 0;
  for (auto& s : b->news) need += (int64_t)s.size() + 1;
  if (out && cap >= need) {
    char* p = out;
    for (auto& s : b->news) {
      memcpy(p, s.data(), s.size());
      p += s.size();
      *p++ = '\n';
    }
  }
  return need;
}

extern "C" void nrb_csr_free(void* handle) { delete (CsrBuild*)handle; }
