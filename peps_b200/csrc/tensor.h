// Walker-batched dense tensors, a caching device allocator and the einsum -> offset-table planner.
// Backend-agnostic host code (compiled with nvcc for the product, g++ for tests/hostsim).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "backend.h"

namespace peps {

// [W][d0][d1]...[d(rank-1)] row-major, walker stride = n (product of dims)
struct BT {
  double *p = nullptr;
  int rank = 0;
  int d[6] = {1, 1, 1, 1, 1, 1};
  long n = 0;
  bool valid() const { return p != nullptr; }
};

class Pool {
 public:
  ~Pool() { release_all(); }
  void *get(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    // best fit among the cached buffers of up to 25 % more than asked for: the many slightly different shapes of a
    // rank-adaptive chain reuse each other's buffers instead of growing the pool
    for (auto it = free_.lower_bound(bytes); it != free_.end() && it->first <= bytes + bytes / 4; ++it) {
      if (it->second.empty()) continue;
      void *p = it->second.back();
      it->second.pop_back();
      return p;
    }
    void *p = be_malloc(bytes);
    size_[p] = bytes;
    total_ += bytes;
    return p;
  }
  void put(void *p) {
    if (!p) return;
    auto it = size_.find(p);
    if (it == size_.end()) throw std::logic_error("Pool::put: unknown pointer");
    free_[it->second].push_back(p);
  }
  void release_all() {
    for (auto &kv : size_) be_free(kv.first);
    size_.clear();
    free_.clear();
    total_ = 0;
  }
  size_t total_bytes() const { return total_; }

 private:
  std::map<size_t, std::vector<void *>> free_;
  std::unordered_map<void *, size_t> size_;
  size_t total_ = 0;
};

// One cached contraction plan: tables on the device + the GettDesc pointing at them.
struct Plan {
  GettDesc d;
  std::vector<int> outdims;
  long outn = 0;
};

class Planner {
 public:
  ~Planner() { release_all(); }
  void release_all() {
    for (void *p : owned_) be_free(p);
    owned_.clear();
    cache_.clear();
    tabs_.clear();
  }
  // spec "abc,cd->abd"; dims of the two inputs given; a label's stride in its tensor is row-major from dims
  // unless explicit strides are passed. Output is row-major contiguous over the out labels (or out_strides).
  const Plan &get(const std::string &spec, const int *da, int ra, const int *db, int rb,
                  const long *sa = nullptr, const long *sb = nullptr, const long *sc = nullptr) {
    std::string key = spec;
    auto app = [&](const int *d, int r) { for (int i = 0; i < r; ++i) key += "," + std::to_string(d[i]); key += ";"; };
    auto apps = [&](const long *s, int r) { if (s) for (int i = 0; i < r; ++i) key += "s" + std::to_string(s[i]); key += ";"; };
    app(da, ra); app(db, rb); apps(sa, ra); apps(sb, rb);
    size_t arrow = spec.find("->"), comma = spec.find(',');
    std::string la = spec.substr(0, comma), lb = spec.substr(comma + 1, arrow - comma - 1), lc = spec.substr(arrow + 2);
    apps(sc, (int)lc.size());
    auto it = cache_.find(key);
    if (it != cache_.end()) return it->second;
    if ((int)la.size() != ra || (int)lb.size() != rb) throw std::logic_error("Planner: rank mismatch in " + spec);
    std::map<char, int> dim;
    std::map<char, long> stA, stB, stC;
    auto strides = [&](const std::string &l, const int *d, const long *s, std::map<char, long> &st) {
      long acc = 1;
      for (int i = (int)l.size() - 1; i >= 0; --i) {
        if (dim.count(l[i]) && dim[l[i]] != d[i]) throw std::logic_error("Planner: dim mismatch for label in " + spec);
        dim[l[i]] = d[i];
        st[l[i]] = s ? s[i] : acc;
        acc *= d[i];
      }
    };
    strides(la, da, sa, stA);
    strides(lb, db, sb, stB);
    Plan pl;
    {
      long acc = 1;
      pl.outdims.resize(lc.size());
      for (int i = (int)lc.size() - 1; i >= 0; --i) {
        if (!dim.count(lc[i])) throw std::logic_error("Planner: unknown output label in " + spec);
        pl.outdims[i] = dim[lc[i]];
        stC[lc[i]] = sc ? sc[i] : acc;
        acc *= dim[lc[i]];
      }
      pl.outn = acc;
    }
    // classify labels; enumerate M and N in output order, K in A order
    std::string ml, nl, kl;
    for (char c : lc) {
      if (stA.count(c) && !stB.count(c)) ml += c;
      else if (stB.count(c) && !stA.count(c)) nl += c;
      else throw std::logic_error("Planner: batch labels are not supported: " + spec);
    }
    for (char c : la) if (stB.count(c)) kl += c;
    for (char c : la) if (!stB.count(c) && lc.find(c) == std::string::npos) throw std::logic_error("Planner: dangling label in " + spec);
    auto table = [&](const std::string &labels, const std::map<char, long> &st) {
      long total = 1;
      for (char c : labels) total *= dim[c];
      std::vector<int32_t> tab((size_t)total);
      for (long i = 0; i < total; ++i) {
        long rem = i, off = 0;
        for (int j = (int)labels.size() - 1; j >= 0; --j) {
          int dd = dim[labels[j]];
          off += (rem % dd) * st.at(labels[j]);
          rem /= dd;
        }
        tab[(size_t)i] = (int32_t)off;
      }
      return tab;
    };
    auto up = [&](const std::vector<int32_t> &v) { return upload(v); };
    std::vector<int32_t> am = table(ml, stA), ak = table(kl, stA), bk = table(kl, stB), bn = table(nl, stB),
                         cm = table(ml, stC), cn = table(nl, stC);
    pl.d.M = (int)am.size(); pl.d.N = (int)bn.size(); pl.d.K = (int)ak.size();
    pl.d.am = up(am); pl.d.ak = up(ak); pl.d.bk = up(bk); pl.d.bn = up(bn); pl.d.cm = up(cm); pl.d.cn = up(cn);
    // fast-index hints: is the smallest-stride label of A (B) with dim > 1 a contracted (free) one?
    auto min_label = [&](const std::string &l, const std::map<char, long> &st) {
      char best = 0; long bs = -1;
      for (char c : l) if (dim[c] > 1 && (bs < 0 || st.at(c) < bs)) { bs = st.at(c); best = c; }
      return best;
    };
    char fa = min_label(la, stA), fb = min_label(lb, stB);
    pl.d.a_kfast = (fa && kl.find(fa) != std::string::npos) ? 1 : 0;
    pl.d.b_nfast = (fb && kl.find(fb) != std::string::npos) ? 0 : 1;
    pl.d.a_vec2 = pairs_ok(pl.d.a_kfast ? ak : am, pl.d.a_kfast ? am : ak);
    pl.d.b_vec2 = pairs_ok(pl.d.b_nfast ? bn : bk, pl.d.b_nfast ? bk : bn);
    pl.d.c_vec2 = pairs_ok(cn, cm);
    return cache_.emplace(key, pl).first->second;
  }

  // 16-byte access along `fast`: even length, pairs (2i, 2i+1) adjacent, every offset that starts a pair even
  static int pairs_ok(const std::vector<int32_t> &fast, const std::vector<int32_t> &other) {
    if (fast.size() < 2 || (fast.size() & 1)) return 0;
    for (size_t i = 0; i < fast.size(); i += 2)
      if ((fast[i] & 1) || fast[i + 1] != fast[i] + 1) return 0;
    for (int32_t o : other) if (o & 1) return 0;
    return 1;
  }

  // generic integer table upload, cached by content
  const int32_t *upload(const std::vector<int32_t> &v) {
    std::string key((const char *)v.data(), v.size() * sizeof(int32_t));
    auto it = tabs_.find(key);
    if (it != tabs_.end()) return it->second;
    int32_t *p = (int32_t *)be_malloc(v.size() * sizeof(int32_t) + 8);
    if (!v.empty()) be_h2d(p, v.data(), v.size() * sizeof(int32_t));
    owned_.push_back(p);
    tabs_[key] = p;
    return p;
  }

 private:
  std::unordered_map<std::string, Plan> cache_;
  std::unordered_map<std::string, const int32_t *> tabs_;
  std::vector<void *> owned_;
};

}  // namespace peps
