// Host side of the aggregation multigrid (amg.cuh): aggregates, coarse BSR patterns, transfer
// lists and -- on a partitioned run -- the owned/ghost numbering and halo plan of every level.
// Plain C++17, no CUDA in here, so tests/test_amg_host.py can drive it on the CPU with several
// simulated parts (threads) and check the Galerkin lists against a global computation.
//
// Partitioned hierarchy (replaces the role MueLu's distributed aggregation + repartitioning plays
// behind the reference's Teko/MueLu preconditioner, src/linear_solve.cpp:74-105):
//   * aggregates never cross parts: every part aggregates its owned nodes on its owned x owned graph;
//   * a distributed coarse level keeps the fine level's layout: rows = owned aggregates, columns =
//     owned aggregates followed by ghost aggregates (sorted by owner part, then by the owner's
//     numbering, so each neighbour fills one contiguous range).  Its halo plan follows from the fine
//     plan without extra messages: what a part sends to a neighbour are the aggregates of the nodes
//     it sends on the fine level;
//   * once the GLOBAL size of a level drops under `replicate_max_nodes` the level is replicated:
//     global numbering, every part holds the whole operator (its Galerkin contributions are summed
//     over the parts at set-up, the restricted right-hand side once per cycle) and the remaining
//     levels are computed redundantly with no communication at all (they are launch-latency bound).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <utility>
#include <vector>

namespace c8 {

struct HaloPlanHost {        // same meaning as the arguments of c8_set_halo_plan
  int n_owned = 0;
  std::vector<int> nbr_rank, send_ptr, send_nodes, recv_ptr;
  int n_send() const { return send_ptr.empty() ? 0 : send_ptr.back(); }
  int n_recv() const { return recv_ptr.empty() ? 0 : recv_ptr.back(); }
};

// collective services the build needs (host vectors; the device implementation stages them)
struct AmgCollectives {
  int rank = 0, nranks = 1;
  virtual ~AmgCollectives() {}
  virtual void allreduce(std::vector<double>& buf) = 0;            // sum over the parts
  virtual void halo(int level, std::vector<double>& vec) = 0;      // one double per node of the level
  virtual int add_level(const HaloPlanHost& plan) = 0;             // registers a plan, returns its id
  virtual HaloPlanHost plan(int level) const = 0;
};

struct AmgLevelHost {
  int n = 0;          // rows (owned nodes; all nodes of a serial or replicated level)
  int ld = 0;         // vector length: owned + ghost
  int nnzb = 0;
  int halo_level = -1;            // >= 0: partitioned level, id of its halo plan
  std::vector<int> rowptr, colind;  // level >= 1 only (level 0 is the context's pattern)
  // transfer to the next coarser level
  int nc_rows = 0;                // rows of the coarse level = aggregates listed in aggptr
  bool coarse_replicated = false; // the coarse level is global: sum its values and rhs over the parts
  std::vector<int> agg;           // [ld] index into the coarse VECTOR (owned + ghost, or global)
  std::vector<int> aggptr, aggmem;  // owned members of every coarse row
  std::vector<int> cptr, cmem;      // fine blocks summed into every coarse block
};

struct AmgBuildOptions {
  int coarsest_max_nodes = 40;
  int max_levels = 12;
  int max_aggregate_size = 8;
  int coarse_aggregate_size = 8;
  int replicate_max_nodes = 30000;
};

// Greedy aggregation on the node graph (columns must all be < n).  max_size <= 0: a root node + all
// its neighbours when none of them is taken yet.  max_size > 0: bounded compact aggregates -- a free
// root takes the free neighbours that share the most neighbours with it.  Leftovers join the
// neighbouring aggregate they are most connected to.
inline void amg_aggregate(int n, const std::vector<int>& rowptr, const std::vector<int>& colind,
                          std::vector<int>& agg, int& nc, int max_size) {
  agg.assign(n, -1);
  nc = 0;
  if (max_size <= 0) {
    for (int i = 0; i < n; ++i) {
      if (agg[i] != -1) continue;
      bool free_nbrs = true;
      for (int k = rowptr[i]; k < rowptr[i + 1] && free_nbrs; ++k)
        if (agg[colind[k]] != -1) free_nbrs = false;
      if (!free_nbrs) continue;
      agg[i] = nc;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) agg[colind[k]] = nc;
      ++nc;
    }
  } else {
    std::vector<std::pair<int, int>> cand;
    std::vector<char> mark(n, 0);
    for (int i = 0; i < n; ++i) {
      if (agg[i] != -1) continue;
      cand.clear();
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) mark[colind[k]] = 1;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colind[k];
        if (j == i || agg[j] != -1) continue;
        int common = 0;
        for (int k2 = rowptr[j]; k2 < rowptr[j + 1]; ++k2) common += mark[colind[k2]];
        cand.emplace_back(-common, j);
      }
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) mark[colind[k]] = 0;
      if (int(cand.size()) + 1 < (max_size + 1) / 2) continue;  // too few free neighbours: leftover
      std::sort(cand.begin(), cand.end());
      agg[i] = nc;
      for (int q = 0; q < int(cand.size()) && q < max_size - 1; ++q) agg[cand[q].second] = nc;
      ++nc;
    }
  }
  std::vector<int> agg2 = agg;
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    int best = -1, best_cnt = 0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int a = agg[colind[k]];
      if (a == -1) continue;
      int c = 0;
      for (int k2 = rowptr[i]; k2 < rowptr[i + 1]; ++k2) c += (agg[colind[k2]] == a);
      if (c > best_cnt || (c == best_cnt && a < best)) { best_cnt = c; best = a; }
    }
    agg2[i] = best;
  }
  agg.swap(agg2);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    agg[i] = nc;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (agg[colind[k]] == -1) agg[colind[k]] = nc;
    ++nc;
  }
}

// Builds the level list.  Level 0: rows [0, n0) of the pattern (rowptr0, colind0) with columns in
// [0, ld0).  comm == nullptr (or one part): serial hierarchy; ghost columns (>= n0) of a part are
// then dropped, i.e. the hierarchy acts on the owned x owned block.  lv[0] carries only the transfer
// lists (its pattern is the caller's).
inline void amg_build_host(int n0, int ld0, const int* rowptr0, const int* colind0,
                           AmgCollectives* comm, int halo_level0, const AmgBuildOptions& opt,
                           std::vector<AmgLevelHost>& lv) {
  lv.clear();
  bool dist = comm != nullptr && comm->nranks > 1;
  const int rank = dist ? comm->rank : 0, nranks = dist ? comm->nranks : 1;
  std::vector<int> rowptr(n0 + 1, 0), colind, blk;   // blk: block of this level's value array
  colind.reserve(rowptr0[n0]);
  blk.reserve(rowptr0[n0]);
  for (int i = 0; i < n0; ++i) {
    for (int k = rowptr0[i]; k < rowptr0[i + 1]; ++k) {
      const int j = colind0[k];
      if (dist || j < n0) { colind.push_back(j); blk.push_back(k); }
    }
    rowptr[i + 1] = int(colind.size());
  }
  AmgLevelHost L0;
  L0.n = n0; L0.ld = dist ? ld0 : n0; L0.nnzb = rowptr0[n0];
  L0.halo_level = dist ? halo_level0 : -1;
  lv.push_back(L0);
  int n = n0, ld = L0.ld, hl = L0.halo_level;
  auto global_sum = [&](int mine, std::vector<long long>* off) -> long long {
    if (!dist) { if (off) { off->assign(2, 0); (*off)[1] = mine; } return mine; }
    std::vector<double> cnt(nranks, 0.0);
    cnt[rank] = mine;
    comm->allreduce(cnt);
    long long s = 0;
    if (off) off->assign(nranks + 1, 0);
    for (int r = 0; r < nranks; ++r) { s += (long long)cnt[r]; if (off) (*off)[r + 1] = s; }
    return s;
  };
  long long n_glob = global_sum(n, nullptr);
  while (n_glob > opt.coarsest_max_nodes && int(lv.size()) < opt.max_levels) {
    // ---- aggregates of the owned nodes (owned x owned graph)
    std::vector<int> agg;
    int nc = 0;
    const int agg_size = lv.size() == 1 ? opt.max_aggregate_size : opt.coarse_aggregate_size;
    if (ld > n) {
      std::vector<int> orp(n + 1, 0), oci;
      oci.reserve(colind.size());
      for (int i = 0; i < n; ++i) {
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          if (colind[k] < n) oci.push_back(colind[k]);
        orp[i + 1] = int(oci.size());
      }
      amg_aggregate(n, orp, oci, agg, nc, agg_size);
    } else {
      amg_aggregate(n, rowptr, colind, agg, nc, agg_size);
    }
    std::vector<long long> off;
    const long long nc_glob = global_sum(nc, &off);
    if (nc_glob >= n_glob) break;  // no coarsening possible (the same decision on every part)
    const bool replicate = dist && nc_glob <= opt.replicate_max_nodes;
    AmgLevelHost C;
    std::vector<int> aggx(ld, 0);  // aggregate of every local column, in the coarse vector's numbering
    if (!dist) {
      aggx = agg;
      C.n = C.ld = nc;
    } else if (replicate) {
      std::vector<double> v(ld, 0.0);
      for (int i = 0; i < n; ++i) v[i] = double(off[rank] + agg[i]);
      comm->halo(hl, v);
      for (int i = 0; i < ld; ++i) aggx[i] = int(v[i]);
      C.n = C.ld = int(nc_glob);
    } else {
      std::vector<double> v(ld, 0.0);
      for (int i = 0; i < n; ++i) v[i] = double(agg[i]);
      comm->halo(hl, v);
      const HaloPlanHost P = comm->plan(hl);
      HaloPlanHost Pc;
      Pc.n_owned = nc;
      Pc.nbr_rank = P.nbr_rank;
      Pc.send_ptr.assign(1, 0);
      Pc.recv_ptr.assign(1, 0);
      for (int i = 0; i < n; ++i) aggx[i] = agg[i];
      std::vector<int> u;
      for (size_t k = 0; k < P.nbr_rank.size(); ++k) {
        u.clear();
        for (int q = P.send_ptr[k]; q < P.send_ptr[k + 1]; ++q) u.push_back(agg[P.send_nodes[q]]);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        Pc.send_nodes.insert(Pc.send_nodes.end(), u.begin(), u.end());
        Pc.send_ptr.push_back(int(Pc.send_nodes.size()));
        u.clear();
        for (int g = P.recv_ptr[k]; g < P.recv_ptr[k + 1]; ++g) u.push_back(int(v[n + g]));
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        for (int g = P.recv_ptr[k]; g < P.recv_ptr[k + 1]; ++g)
          aggx[n + g] = nc + Pc.recv_ptr.back() +
                        int(std::lower_bound(u.begin(), u.end(), int(v[n + g])) - u.begin());
        Pc.recv_ptr.push_back(Pc.recv_ptr.back() + int(u.size()));
      }
      C.n = nc;
      C.ld = nc + Pc.recv_ptr.back();
      C.halo_level = comm->add_level(Pc);
    }
    // ---- coarse pattern of this part's rows: unique (agg[i], agg[j]) pairs + the fine blocks of each
    const int row0 = replicate ? int(off[rank]) : 0;   // first coarse row of this part
    std::vector<std::pair<uint64_t, int>> keys;
    keys.reserve(colind.size());
    for (int i = 0; i < n; ++i)
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
        keys.emplace_back((uint64_t(uint32_t(aggx[i])) << 32) | uint32_t(aggx[colind[k]]), blk[k]);
    std::sort(keys.begin(), keys.end());
    std::vector<int> own_rowlen(nc, 0), own_colind, own_cptr(1, 0), cmem;
    cmem.reserve(keys.size());
    for (size_t q = 0; q < keys.size();) {
      size_t e = q;
      while (e < keys.size() && keys[e].first == keys[q].first) { cmem.push_back(keys[e].second); ++e; }
      own_colind.push_back(int(keys[q].first & 0xffffffffu));
      own_rowlen[int(keys[q].first >> 32) - row0] += 1;
      own_cptr.push_back(int(cmem.size()));
      q = e;
    }
    std::vector<int> c_rowptr(C.n + 1, 0), c_colind, cptr;
    if (!replicate) {
      for (int I = 0; I < nc; ++I) c_rowptr[I + 1] = c_rowptr[I] + own_rowlen[I];
      c_colind.swap(own_colind);
      cptr.swap(own_cptr);
    } else {
      // the parts' rows are disjoint and contiguous in the global numbering: summing zero-padded
      // arrays over the parts concatenates them
      std::vector<double> rl(C.n, 0.0);
      for (int I = 0; I < nc; ++I) rl[row0 + I] = own_rowlen[I];
      comm->allreduce(rl);
      for (int I = 0; I < C.n; ++I) c_rowptr[I + 1] = c_rowptr[I] + int(rl[I]);
      const int nnz = c_rowptr[C.n], base = c_rowptr[row0];
      std::vector<double> cg(nnz, 0.0);
      for (size_t q = 0; q < own_colind.size(); ++q) cg[base + q] = double(own_colind[q]);
      comm->allreduce(cg);
      c_colind.resize(nnz);
      for (int k = 0; k < nnz; ++k) c_colind[k] = int(cg[k]);
      cptr.assign(nnz + 1, 0);
      const int n_own = int(own_colind.size());
      for (int k = 0; k <= nnz; ++k)
        cptr[k] = k < base ? 0 : (k - base <= n_own ? own_cptr[k - base] : own_cptr[n_own]);
    }
    // ---- owned members of every coarse row
    std::vector<int> aggptr(C.n + 1, 0), aggmem(n);
    for (int i = 0; i < n; ++i) aggptr[aggx[i] + 1] += 1;
    for (int I = 0; I < C.n; ++I) aggptr[I + 1] += aggptr[I];
    {
      std::vector<int> pos(aggptr.begin(), aggptr.end() - 1);
      for (int i = 0; i < n; ++i) aggmem[pos[aggx[i]]++] = i;
    }
    AmgLevelHost& F = lv.back();
    F.nc_rows = C.n;
    F.coarse_replicated = replicate;
    F.agg.swap(aggx);
    F.aggptr.swap(aggptr);
    F.aggmem.swap(aggmem);
    F.cptr.swap(cptr);
    F.cmem.swap(cmem);
    C.nnzb = int(c_colind.size());
    C.rowptr = c_rowptr;
    C.colind = c_colind;
    lv.push_back(C);
    // next level's graph: its blocks are the coarse blocks themselves
    rowptr.swap(c_rowptr);
    colind.swap(c_colind);
    blk.resize(colind.size());
    std::iota(blk.begin(), blk.end(), 0);
    n = C.n; ld = C.ld;
    if (replicate) { dist = false; hl = -1; } else { hl = C.halo_level; }
    n_glob = nc_glob;
  }
}

}  // namespace c8
