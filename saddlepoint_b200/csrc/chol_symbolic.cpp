// chol_symbolic.cpp -- host symbolic phase of the supernodal (multifrontal) Cholesky.
//
// The reference calls Eigen::SimplicialLLT::compute on every LM iteration
// (EigenSolverWrapper.h:58): AMD ordering + elimination tree + scalar up-looking numeric
// factorisation, all redone each time [Eigen-upstream].  Here ordering, tree, supernodes,
// row structures and every scatter/extend-add map are computed ONCE per pattern; the numeric
// phase (cholesky.cu) only streams values through them.
//
// Ordering: nested dissection by breadth-first level structures (George), because the GPU
// wants a shallow, balanced assembly tree whose nodes are big dense blocks:
//   * a connected region is split by the median level of a BFS from a pseudo-peripheral
//     vertex; the level is a vertex separator by construction;
//   * a disconnected region is split into two balanced groups of components (no separator);
//   * regions of <= kLeaf vertices, or whose level structure is too shallow to cut (dense
//     graphs), become leaves.
// Every leaf and every separator is ONE supernode whose pivot block is treated as dense
// (a little explicit fill for much more regular tensor-core work).  Row structures follow
// from the usual supernodal recurrence struct(s) = adj(cols(s)) U struct(children) above the
// pivot block; the assembly-tree parent is the supernode owning min(struct(s)).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <numeric>
#include <thread>

#include "chol_internal.h"

namespace spb {

namespace {

constexpr int kLeaf = 48;

// Nested dissection.  The recursion is a tree of independent subproblems (the two sides of a
// separator share no edge), so the top of it runs on host threads: a region above kParallelMin
// vertices is split by the calling thread, its first child handed to another thread (down to a
// depth that matches the machine), the results concatenated in the serial algorithm's order
// (first child, second child, separator) -- the ordering is identical to the single-threaded one.
// `region` and `dist` are shared scratch arrays; a task only touches the entries of its own
// vertices, except for reading region[] of neighbours, which is done with relaxed atomics (a
// neighbour outside the region can never carry this region's id).
struct NdBuilder {
  int n;
  const std::vector<int>& Ap;
  const std::vector<int>& Ai;
  std::vector<int> region;   // region id per vertex (current recursion owner)
  std::vector<int> dist;     // BFS scratch
  std::atomic<int> next_region{1};

  struct Out {
    std::vector<int> perm;      // order of this subproblem's vertices (perm[new] = old)
    std::vector<int> sn_sizes;  // supernode sizes in that order
    void emit(const std::vector<int>& verts) {
      if (verts.empty()) return;
      perm.insert(perm.end(), verts.begin(), verts.end());
      sn_sizes.push_back((int)verts.size());
    }
    void append(Out&& o) {
      perm.insert(perm.end(), o.perm.begin(), o.perm.end());
      sn_sizes.insert(sn_sizes.end(), o.sn_sizes.begin(), o.sn_sizes.end());
    }
  };

  NdBuilder(int n_, const std::vector<int>& p, const std::vector<int>& i) : n(n_), Ap(p), Ai(i), region(n_, 0), dist(n_, -1) {}

  int region_of(int v) const { return __atomic_load_n(&region[v], __ATOMIC_RELAXED); }
  void set_region(int v, int rid) { __atomic_store_n(&region[v], rid, __ATOMIC_RELAXED); }

  // BFS inside region `rid` from `src`; fills order (visit order) and dist; returns depth
  int bfs(int src, int rid, std::vector<int>& order) {
    order.clear();
    order.push_back(src);
    dist[src] = 0;
    int depth = 0;
    for (size_t h = 0; h < order.size(); ++h) {
      int v = order[h];
      depth = dist[v];
      for (int q = Ap[v]; q < Ap[v + 1]; ++q) {
        int w = Ai[q];
        if (region_of(w) == rid && dist[w] < 0) {
          dist[w] = dist[v] + 1;
          order.push_back(w);
        }
      }
    }
    return depth;
  }
  void clear_dist(const std::vector<int>& order) {
    for (int v : order) dist[v] = -1;
  }

  // One dissection step on `verts` (consumed).  Returns false when the region is a leaf (verts kept: emit it as
  // one supernode); otherwise fills the two children A, B and the separator S (empty for a disconnected region).
  bool split_step(std::vector<int>& verts, std::vector<int>& A, std::vector<int>& B, std::vector<int>& S) {
    const int sz = (int)verts.size();
    if (sz <= kLeaf) return false;
    std::vector<int> order, order2;
    const int rid = next_region.fetch_add(1, std::memory_order_relaxed);
    for (int v : verts) set_region(v, rid);
    // connected components
    std::vector<std::vector<int>> comps;
    for (int v : verts) {
      if (dist[v] >= 0) continue;
      bfs(v, rid, order);
      comps.push_back(order);
    }
    for (int v : verts) dist[v] = -1;
    if (comps.size() > 1) {
      std::sort(comps.begin(), comps.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
      for (auto& c : comps) {
        std::vector<int>& dst = A.size() <= B.size() ? A : B;
        dst.insert(dst.end(), c.begin(), c.end());
      }
      std::vector<int>().swap(verts);
      return true;  // no separator; the children are independent subtrees
    }
    // single component: pseudo-peripheral start (two sweeps), then median level cut
    bfs(verts[0], rid, order);
    int far = order.back();
    clear_dist(order);
    int depth = bfs(far, rid, order2);
    if (depth < 2) {  // too shallow to cut: dense-ish block
      clear_dist(order2);
      return false;
    }
    std::vector<int> cnt(depth + 1, 0);
    for (int v : order2) cnt[dist[v]]++;
    // choose the level in the middle third (by cumulative size) with the fewest vertices
    int best = -1;
    {
      int cum = 0;
      for (int l = 0; l <= depth; ++l) {
        int before = cum;
        cum += cnt[l];
        if (l == 0 || l == depth) continue;
        int after = sz - cum;
        if (before < sz / 4 || after < sz / 4) continue;
        if (best < 0 || cnt[l] < cnt[best]) best = l;
      }
      if (best < 0) {  // fall back to the median level
        cum = 0;
        for (int l = 0; l <= depth; ++l) {
          cum += cnt[l];
          if (cum * 2 >= sz) {
            best = std::min(std::max(l, 1), depth - 1);
            break;
          }
        }
      }
    }
    if (cnt[best] * 2 > sz) {  // separator would swallow the region
      clear_dist(order2);
      return false;
    }
    for (int v : order2) {
      int d = dist[v];
      (d < best ? A : d > best ? B : S).push_back(v);
    }
    clear_dist(order2);
    std::vector<int>().swap(verts);
    return true;
  }

  // serial dissection of one region (explicit stack: no deep recursion on path-like graphs)
  void dissect_serial(std::vector<int>& verts, Out& out) {
    struct Frame {
      std::vector<int> verts;
      std::vector<int> sep;  // emitted after the children
      int stage;
    };
    std::vector<Frame> stack;
    stack.push_back({std::move(verts), {}, 0});
    while (!stack.empty()) {
      Frame& F = stack.back();
      if (F.stage == 1) {  // children done: emit separator
        out.emit(F.sep);
        stack.pop_back();
        continue;
      }
      F.stage = 1;
      std::vector<int> A, B, S;
      if (!split_step(F.verts, A, B, S)) {
        out.emit(F.verts);
        stack.pop_back();
        continue;
      }
      F.sep = std::move(S);
      stack.push_back({std::move(B), {}, 0});  // note: invalidates F
      stack.push_back({std::move(A), {}, 0});
    }
  }

  static constexpr int kParallelMin = 20000;

  // parallel top of the recursion
  void dissect(std::vector<int>& verts, Out& out, int depth_left) {
    if (depth_left <= 0 || (int)verts.size() < kParallelMin) {
      dissect_serial(verts, out);
      return;
    }
    std::vector<int> A, B, S;
    if (!split_step(verts, A, B, S)) {
      out.emit(verts);
      return;
    }
    Out oa, ob;
    std::exception_ptr err;
    std::thread ta([&] {
      try {
        dissect(A, oa, depth_left - 1);
      } catch (...) {
        err = std::current_exception();
      }
    });
    try {
      dissect(B, ob, depth_left - 1);
    } catch (...) {
      ta.join();
      throw;
    }
    ta.join();
    if (err) std::rethrow_exception(err);
    out.append(std::move(oa));
    out.append(std::move(ob));
    out.emit(S);
  }
};

}  // namespace

void chol_symbolic(const HPattern& H, const SellLayout& S, CholSymbolic& C) {
  const int n = H.n;
  C.n = n;
  const bool timing = getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 1;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb analyze timing]   chol_symbolic: %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  // ---- ordering + supernode partition
  // Static condensation first.  Least-squares problems from meshes carry many low-degree "satellite" unknowns (the
  // seamless-integration structure: 8 field unknowns per face that couple with each other and with the face's 6 vertex
  // unknowns only, SIInitialSolutionTraits.h:238-245).  A breadth-first level of the full graph contains the satellites of
  // every face it crosses, so the separators come out ~9x larger than the mesh needs and the factorisation ~100x more
  // expensive.  Connected groups of low-degree vertices with a small outside neighbourhood are therefore eliminated FIRST,
  // each as one leaf supernode (no two groups are adjacent: all of them are independent fronts of the first level), and
  // the nested dissection runs on what is left, with the fill those eliminations cause (a clique over every group's
  // outside neighbours) added to its graph.  Patterns without such groups (LF: uniform degree) are untouched.
  std::vector<int> cond_perm, cond_sizes, rest;          // eliminated groups (original ids) / remaining vertices
  std::vector<int> rp, ri;                               // graph of the remaining vertices (compact ids), with fill
  {
    const int kMaxDeg = getenv("SPB_CHOL_CONDENSE_DEG") ? atoi(getenv("SPB_CHOL_CONDENSE_DEG")) : 16;  // 0 switches the condensation off (tests compare both)
    constexpr int kMaxGroup = 32, kMaxExt = 48;
    std::vector<int> grp(n, -1);  // -1 untouched, -2 rejected candidate, >= 0 group id
    std::vector<int> comp, ext, stackv;
    std::vector<int> seen(n, -1);
    std::vector<std::vector<int>> group_ext;
    if (kMaxDeg > 0) {
      for (int v0 = 0; v0 < n; ++v0) {
        if (grp[v0] != -1 || H.colptr[(size_t)v0 + 1] - H.colptr[v0] - 1 > kMaxDeg) continue;
        // component of low-degree vertices around v0 (abandoned once it is too big to be a satellite group)
        comp.clear();
        ext.clear();
        stackv.assign(1, v0);
        seen[v0] = v0;
        bool ok = true;
        while (!stackv.empty()) {
          const int v = stackv.back();
          stackv.pop_back();
          comp.push_back(v);
          if ((int)comp.size() > kMaxGroup) ok = false;
          for (int q = H.colptr[v]; q < H.colptr[(size_t)v + 1]; ++q) {
            const int w = H.rowidx[q];
            if (w == v || seen[w] == v0) continue;
            seen[w] = v0;
            if (H.colptr[(size_t)w + 1] - H.colptr[w] - 1 <= kMaxDeg) {
              if (ok) stackv.push_back(w);  // a rejected component is not explored further than needed to mark it
              else comp.push_back(w);
            } else {
              ext.push_back(w);
            }
          }
          if (!ok && comp.size() > 4096) break;  // giant low-degree region (a plain grid): stop walking it
        }
        ok = ok && (int)ext.size() <= kMaxExt && (int)ext.size() <= 2 * (int)comp.size() + 8;
        if (ok) {
          const int g = (int)group_ext.size();
          for (int v : comp) grp[v] = g;
          std::sort(ext.begin(), ext.end());
          group_ext.push_back(ext);
          std::sort(comp.begin(), comp.end());
          cond_perm.insert(cond_perm.end(), comp.begin(), comp.end());
          cond_sizes.push_back((int)comp.size());
        } else {
          for (int v : comp)
            if (grp[v] == -1) grp[v] = -2;
        }
      }
    }
    // a condensation that removes little is not worth a second graph
    if ((int64_t)cond_perm.size() * 8 < n) {
      cond_perm.clear();
      cond_sizes.clear();
      group_ext.clear();
      std::fill(grp.begin(), grp.end(), -1);
    }
    if (!cond_perm.empty()) {
      std::vector<int> newid(n, -1);
      for (int v = 0; v < n; ++v)
        if (grp[v] < 0) {
          newid[v] = (int)rest.size();
          rest.push_back(v);
        }
      // groups adjacent to every remaining vertex
      std::vector<std::vector<int>> groups_of(rest.size());
      for (int g = 0; g < (int)group_ext.size(); ++g)
        for (int w : group_ext[g]) groups_of[newid[w]].push_back(g);
      rp.assign(rest.size() + 1, 0);
      std::vector<int> mark(rest.size(), -1), row;
      for (size_t a = 0; a < rest.size(); ++a) {
        const int v = rest[a];
        row.clear();
        mark[a] = (int)a;
        row.push_back((int)a);  // diagonal: the graph arrays follow the H pattern's convention (full diagonal)
        for (int q = H.colptr[v]; q < H.colptr[(size_t)v + 1]; ++q) {
          const int b = newid[H.rowidx[q]];
          if (b >= 0 && mark[b] != (int)a) {
            mark[b] = (int)a;
            row.push_back(b);
          }
        }
        for (int g : groups_of[a])
          for (int w : group_ext[g]) {
            const int b = newid[w];
            if (mark[b] != (int)a) {
              mark[b] = (int)a;
              row.push_back(b);
            }
          }
        std::sort(row.begin(), row.end());
        ri.insert(ri.end(), row.begin(), row.end());
        rp[a + 1] = (int)ri.size();
      }
    }
  }
  lap("static condensation");
  if (!cond_perm.empty()) {
    const int nr = (int)rest.size();
    NdBuilder nd(nr, rp, ri);
    std::vector<int> all(nr);
    std::iota(all.begin(), all.end(), 0);
    NdBuilder::Out out;
    out.perm.reserve(nr);
    int depth = 0;
    {
      const char* e = getenv("SPB_HOST_THREADS");
      int threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      threads = std::max(1, std::min(threads, 32));
      while ((1 << depth) < 2 * threads && threads > 1) ++depth;
    }
    if (nr > 0) nd.dissect(all, out, depth);
    C.perm = std::move(cond_perm);
    for (int a : out.perm) C.perm.push_back(rest[a]);
    C.sn_col0.assign(1, 0);
    for (int sz : cond_sizes) C.sn_col0.push_back(C.sn_col0.back() + sz);
    for (int sz : out.sn_sizes) C.sn_col0.push_back(C.sn_col0.back() + sz);
  } else {
    NdBuilder nd(n, H.colptr, H.rowidx);
    std::vector<int> all(n);
    std::iota(all.begin(), all.end(), 0);
    NdBuilder::Out out;
    out.perm.reserve(n);
    // 2^depth concurrent subtrees at the bottom of the parallel part: about twice the host threads
    int depth = 0;
    {
      const char* e = getenv("SPB_HOST_THREADS");
      int threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      threads = std::max(1, std::min(threads, 32));
      while ((1 << depth) < 2 * threads && threads > 1) ++depth;
    }
    nd.dissect(all, out, depth);
    C.perm = std::move(out.perm);
    C.sn_col0.assign(1, 0);
    for (int sz : out.sn_sizes) C.sn_col0.push_back(C.sn_col0.back() + sz);
  }
  lap("nested dissection");
  if ((int)C.perm.size() != n) throw ArgFail{"internal: nested dissection lost vertices"};
  C.pinv.assign(n, 0);
  for (int k = 0; k < n; ++k) C.pinv[C.perm[k]] = k;
  const int ns = (int)C.sn_col0.size() - 1;
  C.nsuper = ns;
  std::vector<int> sn_of_col(n);
  for (int s = 0; s < ns; ++s)
    for (int c = C.sn_col0[s]; c < C.sn_col0[s + 1]; ++c) sn_of_col[c] = s;
  // ---- row structures, parents
  C.sn_parent.assign(ns, -1);
  C.struct_ptr.assign((size_t)ns + 1, 0);
  C.struct_idx.clear();
  std::vector<std::vector<int>> children(ns);
  std::vector<int> mark(n, -1), tmp;
  for (int s = 0; s < ns; ++s) {
    const int c0 = C.sn_col0[s], c1 = C.sn_col0[s + 1];
    tmp.clear();
    for (int c = c0; c < c1; ++c) {
      const int jo = C.perm[c];
      for (int q = H.colptr[jo]; q < H.colptr[(size_t)jo + 1]; ++q) {
        const int a = C.pinv[H.rowidx[q]];
        if (a >= c1 && mark[a] != s) {
          mark[a] = s;
          tmp.push_back(a);
        }
      }
    }
    for (int ch : children[s])
      for (int64_t t = C.struct_ptr[ch]; t < C.struct_ptr[(size_t)ch + 1]; ++t) {
        const int a = C.struct_idx[(size_t)t];
        if (a >= c1 && mark[a] != s) {
          mark[a] = s;
          tmp.push_back(a);
        }
      }
    std::sort(tmp.begin(), tmp.end());
    C.struct_idx.insert(C.struct_idx.end(), tmp.begin(), tmp.end());
    C.struct_ptr[(size_t)s + 1] = (int64_t)C.struct_idx.size();
    if (!tmp.empty()) {
      const int p = sn_of_col[tmp[0]];
      C.sn_parent[s] = p;
      children[p].push_back(s);
    }
  }
  lap("row structures");
  // ---- levels (leaves = 0)
  C.sn_level.assign(ns, 0);
  int nlevels = 0;
  for (int s = 0; s < ns; ++s) {
    int lv = 0;
    for (int ch : children[s]) lv = std::max(lv, C.sn_level[ch] + 1);
    C.sn_level[s] = lv;
    nlevels = std::max(nlevels, lv + 1);
  }
  C.nlevels = nlevels;
  C.level_ptr.assign((size_t)nlevels + 1, 0);
  for (int s = 0; s < ns; ++s) C.level_ptr[(size_t)C.sn_level[s] + 1]++;
  for (int l = 0; l < nlevels; ++l) C.level_ptr[(size_t)l + 1] += C.level_ptr[l];
  C.level_sn.resize(ns);
  {
    std::vector<int> cur(C.level_ptr.begin(), C.level_ptr.end() - 1);
    for (int s = 0; s < ns; ++s) C.level_sn[cur[C.sn_level[s]]++] = s;
  }
  // ---- storage offsets: L panel r x k (ld = r), update matrix u x u (ld = u)
  C.L_off.assign((size_t)ns + 1, 0);
  C.U_off.assign((size_t)ns + 1, 0);
  C.W_off.assign((size_t)ns + 1, 0);
  C.max_k = C.max_r = 0;
  C.flops = 0.0;
  for (int s = 0; s < ns; ++s) {
    const int64_t k = C.sn_col0[s + 1] - C.sn_col0[s];
    const int64_t u = C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s];
    const int64_t r = k + u;
    C.L_off[(size_t)s + 1] = C.L_off[s] + r * k;
    C.W_off[(size_t)s + 1] = C.W_off[s] + r;
    C.max_k = std::max<int>(C.max_k, (int)k);
    C.max_r = std::max<int>(C.max_r, (int)r);
    C.flops += (double)k * k * k / 3.0 + (double)u * k * k + (double)u * u * k;
  }
  // ---- update-matrix arena.  U_s is written at level(s) (extend-add from its children, then the
  // SYRK) and read once, by its parent's extend-add at level(parent).  Fronts are grouped by
  // (birth level, death level); a group is one contiguous block, allocated first-fit before its
  // birth level from space no live group occupies and released after its death level.  On nested-
  // dissection trees this keeps two or three levels' worth of update matrices resident instead of
  // all of them (C2: 9 GB -> about 1 GB; the 4.5 M-unknown C3 stand-in: 118 GB -> fits).
  {
    C.U_sum = 0;
    struct Group {
      int birth, death;
      int64_t size, off;
      std::vector<int> members;
    };
    std::vector<Group> groups;
    std::vector<std::vector<int>> by_birth((size_t)nlevels);
    {
      std::vector<int> gid((size_t)nlevels * nlevels, -1);
      for (int s = 0; s < ns; ++s) {
        const int64_t u = C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s];
        if (u == 0) continue;
        const int b = C.sn_level[s];
        const int d = C.sn_parent[s] >= 0 ? C.sn_level[C.sn_parent[s]] : b;
        int& g = gid[(size_t)b * nlevels + d];
        if (g < 0) {
          g = (int)groups.size();
          groups.push_back({b, d, 0, 0, {}});
          by_birth[b].push_back(g);
        }
        groups[g].members.push_back(s);
        groups[g].size += u * u;
        C.U_sum += u * u;
      }
    }
    // free list of (offset, length), sorted by offset; the arena grows at the end when nothing fits
    std::vector<std::pair<int64_t, int64_t>> freel;
    int64_t top = 0;
    auto release = [&](int64_t off, int64_t len) {
      if (len == 0) return;
      freel.emplace_back(off, len);
      std::sort(freel.begin(), freel.end());
      std::vector<std::pair<int64_t, int64_t>> merged;
      for (auto& f : freel) {
        if (!merged.empty() && merged.back().first + merged.back().second == f.first) merged.back().second += f.second;
        else merged.push_back(f);
      }
      if (!merged.empty() && merged.back().first + merged.back().second == top) {  // trailing free space: shrink
        top = merged.back().first;
        merged.pop_back();
      }
      freel.swap(merged);
    };
    C.u_zero_ptr.assign((size_t)nlevels + 1, 0);
    C.u_zero_range.clear();
    int64_t high = 0;
    for (int l = 0; l < nlevels; ++l) {
      for (int g : by_birth[l]) {
        Group& G = groups[g];
        bool placed = false;
        for (size_t f = 0; f < freel.size(); ++f)
          if (freel[f].second >= G.size) {
            G.off = freel[f].first;
            freel[f].first += G.size;
            freel[f].second -= G.size;
            if (freel[f].second == 0) freel.erase(freel.begin() + (long)f);
            placed = true;
            break;
          }
        if (!placed) {
          G.off = top;
          top += G.size;
          high = std::max(high, top);
        }
        int64_t o = G.off;
        for (int s : G.members) {
          const int64_t u = C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s];
          C.U_off[s] = o;
          o += u * u;
        }
        C.u_zero_range.push_back(G.off);
        C.u_zero_range.push_back(G.size);
      }
      C.u_zero_ptr[(size_t)l + 1] = (int)(C.u_zero_range.size() / 2);
      // groups read by this level's extend-add are dead once the level is done
      for (Group& G : groups)
        if (G.death == l && G.birth < l) release(G.off, G.size);
      for (int g : by_birth[l])
        if (groups[g].death == l) release(groups[g].off, groups[g].size);  // a root with a structure cannot occur, but stay safe
    }
    C.U_total = high;
    C.U_off[ns] = C.U_total;
  }
  C.nnzL = 0;
  for (int s = 0; s < ns; ++s) {
    const int64_t k = C.sn_col0[s + 1] - C.sn_col0[s];
    const int64_t u = C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s];
    C.nnzL += k * (k + 1) / 2 + u * k;
  }
  lap("levels, offsets, update-matrix arena");
  // ---- A scatter map: every lower entry (in the permuted order) of H -> position in its panel
  {
    const int64_t nlow = (H.nnz() + n) / 2;
    C.a_src.clear();
    C.a_dst.clear();
    C.a_src.reserve((size_t)nlow);
    C.a_dst.reserve((size_t)nlow);
    for (int jo = 0; jo < n; ++jo) {
      const int b = C.pinv[jo];
      const int s = sn_of_col[b];
      const int c0 = C.sn_col0[s], c1 = C.sn_col0[s + 1];
      const int64_t k = c1 - c0;
      const int64_t r = k + (C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s]);
      const int* sb = C.struct_idx.data() + C.struct_ptr[s];
      const int* se = C.struct_idx.data() + C.struct_ptr[(size_t)s + 1];
      int kk = 0;
      for (int q = H.colptr[jo]; q < H.colptr[(size_t)jo + 1]; ++q, ++kk) {
        const int a = C.pinv[H.rowidx[q]];
        if (a < b) continue;  // the symmetric partner covers it
        int64_t lrow;
        if (a < c1) {
          lrow = a - c0;
        } else {
          const int* it = std::lower_bound(sb, se, a);
          if (it == se || *it != a) throw ArgFail{"internal: H entry outside the symbolic structure"};
          lrow = k + (it - sb);
        }
        C.a_src.push_back(S.pos(jo, kk));  // row jo of the symmetric storage, entry kk == H(rowidx[q], jo)
        C.a_dst.push_back(C.L_off[s] + (int64_t)(b - c0) * r + lrow);
      }
    }
  }
  // ---- extend-add maps: child struct -> parent front local indices
  C.rel_ptr.assign((size_t)ns + 1, 0);
  C.rel_idx.clear();
  C.rel_idx.reserve(C.struct_idx.size());
  for (int s = 0; s < ns; ++s) {
    const int p = C.sn_parent[s];
    if (p >= 0) {
      const int pc0 = C.sn_col0[p], pc1 = C.sn_col0[p + 1];
      const int pk = pc1 - pc0;
      const int* pb = C.struct_idx.data() + C.struct_ptr[p];
      const int* pe = C.struct_idx.data() + C.struct_ptr[(size_t)p + 1];
      for (int64_t t = C.struct_ptr[s]; t < C.struct_ptr[(size_t)s + 1]; ++t) {
        const int a = C.struct_idx[(size_t)t];
        if (a < pc1) {
          C.rel_idx.push_back(a - pc0);
        } else {
          const int* it = std::lower_bound(pb, pe, a);
          if (it == pe || *it != a) throw ArgFail{"internal: child structure not contained in parent front"};
          C.rel_idx.push_back(pk + (int)(it - pb));
        }
      }
    }
    C.rel_ptr[(size_t)s + 1] = (int64_t)C.rel_idx.size();
  }
  C.child_ptr.assign((size_t)ns + 1, 0);
  C.child_idx.clear();
  for (int s = 0; s < ns; ++s) {
    for (int ch : children[s]) C.child_idx.push_back(ch);
    C.child_ptr[(size_t)s + 1] = (int)C.child_idx.size();
  }
  lap("scatter and extend-add maps");
  C.dinv_off.assign((size_t)ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    const int64_t k = C.sn_col0[s + 1] - C.sn_col0[s];
    C.dinv_off[(size_t)s + 1] = C.dinv_off[s] + ((k + 31) / 32) * 1024;
  }
  // ---- extend-add order and child-rank lists per level.  Children sorted by descending update-matrix size (stable): with
  // static condensation a parent has a few large children (its dissection subtrees) and up to a hundred tiny ones (the
  // satellite groups), and one launch per child rank was 216 launches per factorisation on the seamless-integration
  // structure.  Only the ranks that may hold a large child keep the rank-wise launch.
  C.rank_ptr.clear();
  C.rank_children.clear();
  C.level_rank_ptr.assign((size_t)nlevels + 1, 0);
  C.ea_ptr.assign((size_t)ns + 1, 0);
  C.ea_idx.clear();
  C.level_big_ranks.assign((size_t)nlevels, 0);
  C.small_parents_ptr.assign((size_t)nlevels + 1, 0);
  C.small_parents.clear();
  auto usize = [&](int c) { return (int)(C.struct_ptr[(size_t)c + 1] - C.struct_ptr[c]); };
  for (int s = 0; s < ns; ++s) {
    std::vector<int> ch(children[s]);
    std::stable_sort(ch.begin(), ch.end(), [&](int a, int b) { return usize(a) > usize(b); });
    C.ea_idx.insert(C.ea_idx.end(), ch.begin(), ch.end());
    C.ea_ptr[(size_t)s + 1] = (int)C.ea_idx.size();
  }
  for (int l = 0; l < nlevels; ++l) {
    int big = 0;
    for (int t = C.level_ptr[l]; t < C.level_ptr[(size_t)l + 1]; ++t) {
      const int p = C.level_sn[t];
      int nb = 0;
      for (int q = C.ea_ptr[p]; q < C.ea_ptr[(size_t)p + 1] && usize(C.ea_idx[q]) > kSmallChildU; ++q) ++nb;
      big = std::max(big, nb);
    }
    C.level_big_ranks[l] = big;
    for (int rk = 0; rk < big; ++rk) {
      C.rank_ptr.push_back((int)C.rank_children.size());
      for (int t = C.level_ptr[l]; t < C.level_ptr[(size_t)l + 1]; ++t) {
        const int p = C.level_sn[t];
        if (C.ea_ptr[(size_t)p + 1] - C.ea_ptr[p] > rk) C.rank_children.push_back(C.ea_idx[(size_t)C.ea_ptr[p] + rk]);
      }
    }
    C.level_rank_ptr[(size_t)l + 1] = (int)C.rank_ptr.size();
    for (int t = C.level_ptr[l]; t < C.level_ptr[(size_t)l + 1]; ++t) {
      const int p = C.level_sn[t];
      if (C.ea_ptr[(size_t)p + 1] - C.ea_ptr[p] > big) C.small_parents.push_back(p);
    }
    C.small_parents_ptr[(size_t)l + 1] = (int)C.small_parents.size();
  }
  C.rank_ptr.push_back((int)C.rank_children.size());
}

}  // namespace spb
