// CPU check of the partitioned multigrid hierarchy build (calibr8_b200/csrc/amg_host.hpp).
// P simulated parts (threads + barriers standing in for the halo copy and the allreduce) build the
// hierarchy of a structured 3-D node graph; the main thread then verifies, level by level and in
// GLOBAL numbering, that
//   * the halo plans of every level are consistent (matching counts, ghost = the owner's node),
//   * the aggregate of a ghost node is the aggregate its owner gave it,
//   * every owned node is listed exactly once among the members of its aggregate,
//   * the coarse values each part computes from its (cptr, cmem) lists -- summed over the parts
//     when the level is replicated -- equal the global Galerkin product P^T A P entry by entry.
// usage: amg_host_check nx ny nz px py pz replicate_max_nodes   -> prints "ok ..." or aborts
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <thread>

#include "amg_host.hpp"

using namespace c8;
typedef long long ll;

#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      std::fprintf(stderr, "CHECK failed %s:%d: ", __FILE__, __LINE__); \
      std::fprintf(stderr, __VA_ARGS__);                  \
      std::fprintf(stderr, "\n");                         \
      std::exit(1);                                       \
    }                                                     \
  } while (0)

struct Shared {
  int P;
  std::barrier<> bar;
  std::vector<std::vector<HaloPlanHost>> plans;   // [rank][level]
  std::vector<std::vector<double>*> ptr;
  explicit Shared(int p) : P(p), bar(p), plans(p), ptr(p, nullptr) {}
};

struct ThreadComm : AmgCollectives {
  Shared* sh;
  void allreduce(std::vector<double>& buf) override {
    sh->ptr[rank] = &buf;
    sh->bar.arrive_and_wait();
    std::vector<double> sum(buf.size(), 0.0);
    for (int r = 0; r < nranks; ++r) {
      CHECK(sh->ptr[r]->size() == buf.size(), "allreduce size mismatch");
      for (size_t i = 0; i < buf.size(); ++i) sum[i] += (*sh->ptr[r])[i];
    }
    sh->bar.arrive_and_wait();
    buf.swap(sum);
  }
  void halo(int level, std::vector<double>& vec) override {
    sh->ptr[rank] = &vec;
    sh->bar.arrive_and_wait();
    const HaloPlanHost& P = sh->plans[rank][level];
    CHECK(int(vec.size()) == P.n_owned + P.n_recv(), "halo vector length %zu != %d + %d", vec.size(),
          P.n_owned, P.n_recv());
    std::vector<double> recv(P.n_recv());
    for (size_t k = 0; k < P.nbr_rank.size(); ++k) {
      const int q = P.nbr_rank[k];
      const HaloPlanHost& Q = sh->plans[q][level];
      int kk = -1;
      for (size_t t = 0; t < Q.nbr_rank.size(); ++t)
        if (Q.nbr_rank[t] == rank) kk = int(t);
      CHECK(kk >= 0, "neighbour relation not symmetric");
      const int cnt = P.recv_ptr[k + 1] - P.recv_ptr[k];
      CHECK(Q.send_ptr[kk + 1] - Q.send_ptr[kk] == cnt, "level %d: rank %d expects %d from %d which sends %d",
            level, rank, cnt, q, Q.send_ptr[kk + 1] - Q.send_ptr[kk]);
      for (int t = 0; t < cnt; ++t) recv[P.recv_ptr[k] + t] = (*sh->ptr[q])[Q.send_nodes[Q.send_ptr[kk] + t]];
    }
    sh->bar.arrive_and_wait();
    for (int g = 0; g < P.n_recv(); ++g) vec[P.n_owned + g] = recv[g];
    sh->bar.arrive_and_wait();
  }
  int add_level(const HaloPlanHost& plan) override {
    sh->plans[rank].push_back(plan);
    return int(sh->plans[rank].size()) - 1;
  }
  HaloPlanHost plan(int level) const override { return sh->plans[rank][level]; }
};

static double aval(ll i, ll j) {  // deterministic, non-symmetric matrix entry
  unsigned long long h = (unsigned long long)i * 1000003ull + (unsigned long long)j * 7919ull + 12345ull;
  h ^= h >> 13; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
  return 0.5 + double(h % 100000) / 100000.0 + (i == j ? 30.0 : 0.0);
}

int main(int argc, char** argv) {
  CHECK(argc >= 8, "usage: nx ny nz px py pz replicate_max_nodes");
  const int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
  const int px = atoi(argv[4]), py = atoi(argv[5]), pz = atoi(argv[6]);
  AmgBuildOptions opt;
  opt.replicate_max_nodes = atoi(argv[7]);
  if (argc > 8) opt.coarsest_max_nodes = atoi(argv[8]);
  const int P = px * py * pz;
  const ll N = ll(nx) * ny * nz;
  auto gidx = [&](int x, int y, int z) { return (ll(z) * ny + y) * nx + x; };
  auto owner_of = [&](int x, int y, int z) {
    return (std::min(z * pz / nz, pz - 1) * py + std::min(y * py / ny, py - 1)) * px + std::min(x * px / nx, px - 1);
  };
  // global graph (Kuhn-tet-like 15-point stencil: neighbours whose offset has all components of
  // one sign) + owner of every node
  std::vector<std::vector<ll>> nbrs(N);
  std::vector<int> owner(N);
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        const ll i = gidx(x, y, z);
        owner[i] = owner_of(x, y, z);
        for (int dz = -1; dz <= 1; ++dz)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
              const bool pos = dx >= 0 && dy >= 0 && dz >= 0, neg = dx <= 0 && dy <= 0 && dz <= 0;
              if (!pos && !neg) continue;
              const int X = x + dx, Y = y + dy, Z = z + dz;
              if (X < 0 || Y < 0 || Z < 0 || X >= nx || Y >= ny || Z >= nz) continue;
              nbrs[i].push_back(gidx(X, Y, Z));
            }
        std::sort(nbrs[i].begin(), nbrs[i].end());
      }
  // local numbering of every part: owned by gid, ghosts by (owner, gid)
  std::vector<std::vector<ll>> l2g(P);
  std::vector<std::map<ll, int>> g2l(P);
  std::vector<int> n_owned(P, 0);
  Shared sh(P);
  for (ll i = 0; i < N; ++i) { g2l[owner[i]][i] = int(l2g[owner[i]].size()); l2g[owner[i]].push_back(i); }
  for (int r = 0; r < P; ++r) n_owned[r] = int(l2g[r].size());
  for (int r = 0; r < P; ++r) {
    std::vector<std::pair<int, ll>> ghosts;
    for (int i = 0; i < n_owned[r]; ++i)
      for (ll j : nbrs[l2g[r][i]])
        if (owner[j] != r) ghosts.emplace_back(owner[j], j);
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
    HaloPlanHost pl;
    pl.n_owned = n_owned[r];
    pl.recv_ptr.assign(1, 0);
    for (size_t g = 0; g < ghosts.size(); ++g) {
      if (g == 0 || ghosts[g].first != ghosts[g - 1].first) {
        if (g) pl.recv_ptr.push_back(int(g));
        pl.nbr_rank.push_back(ghosts[g].first);
      }
      g2l[r][ghosts[g].second] = int(l2g[r].size());
      l2g[r].push_back(ghosts[g].second);
    }
    if (!ghosts.empty()) pl.recv_ptr.push_back(int(ghosts.size()));
    sh.plans[r].push_back(pl);
  }
  for (int r = 0; r < P; ++r) {   // send lists: what the neighbour q ghosts of r, in r's order
    HaloPlanHost& pl = sh.plans[r][0];
    pl.send_ptr.assign(1, 0);
    for (int q : pl.nbr_rank) {
      for (int i = 0; i < n_owned[r]; ++i) {
        bool need = false;
        for (ll j : nbrs[l2g[r][i]]) need = need || owner[j] == q;
        if (need) pl.send_nodes.push_back(i);
      }
      pl.send_ptr.push_back(int(pl.send_nodes.size()));
    }
  }
  // local level-0 patterns (rows = owned) and values
  std::vector<std::vector<int>> rp(P), ci(P);
  std::vector<std::vector<double>> vals0(P);
  for (int r = 0; r < P; ++r) {
    rp[r].assign(n_owned[r] + 1, 0);
    for (int i = 0; i < n_owned[r]; ++i) {
      std::vector<int> cols;
      for (ll j : nbrs[l2g[r][i]]) cols.push_back(g2l[r][j]);
      std::sort(cols.begin(), cols.end());
      for (int c : cols) { ci[r].push_back(c); vals0[r].push_back(aval(l2g[r][i], l2g[r][c])); }
      rp[r][i + 1] = int(ci[r].size());
    }
  }
  // ---- build on P threads
  std::vector<std::vector<AmgLevelHost>> lv(P);
  std::vector<ThreadComm> comms(P);
  std::vector<std::thread> th;
  for (int r = 0; r < P; ++r) {
    comms[r].rank = r; comms[r].nranks = P; comms[r].sh = &sh;
    th.emplace_back([&, r]() {
      amg_build_host(n_owned[r], int(l2g[r].size()), rp[r].data(), ci[r].data(), P > 1 ? &comms[r] : nullptr, 0,
                     opt, lv[r]);
    });
  }
  for (auto& t : th) t.join();
  const int nl = int(lv[0].size());
  for (int r = 0; r < P; ++r) CHECK(int(lv[r].size()) == nl, "level counts differ");
  // ---- verify level by level in global numbering
  std::vector<std::vector<ll>> gid = l2g;               // [rank][local] -> global id of this level
  std::map<std::pair<ll, ll>, double> M;                // global operator of this level
  for (ll i = 0; i < N; ++i)
    for (ll j : nbrs[i]) M[{i, j}] = aval(i, j);
  std::vector<std::vector<double>> vals = vals0;
  int n_dist = 0, n_repl = 0;
  for (int l = 0; l + 1 < nl; ++l) {
    const bool fine_dist = lv[0][l].halo_level >= 0;
    const bool repl = lv[0][l].coarse_replicated;
    const bool coarse_dist = lv[0][l + 1].halo_level >= 0;
    n_dist += coarse_dist; n_repl += repl;
    // global ids of the coarse level's local vectors
    std::vector<std::vector<ll>> cg(P);
    std::vector<ll> off(P + 1, 0);
    if (coarse_dist)
      for (int r = 0; r < P; ++r) off[r + 1] = off[r] + lv[r][l + 1].n;
    for (int r = 0; r < P; ++r) {
      const AmgLevelHost& C = lv[r][l + 1];
      cg[r].resize(C.ld);
      if (!coarse_dist) { for (int i = 0; i < C.ld; ++i) cg[r][i] = i; continue; }
      for (int i = 0; i < C.n; ++i) cg[r][i] = off[r] + i;
    }
    if (coarse_dist)
      for (int r = 0; r < P; ++r) {
        const HaloPlanHost& Pl = sh.plans[r][lv[r][l + 1].halo_level];
        CHECK(Pl.n_owned == lv[r][l + 1].n && Pl.n_owned + Pl.n_recv() == lv[r][l + 1].ld, "plan sizes");
        for (size_t k = 0; k < Pl.nbr_rank.size(); ++k) {
          const int q = Pl.nbr_rank[k];
          const HaloPlanHost& Q = sh.plans[q][lv[q][l + 1].halo_level];
          int kk = -1;
          for (size_t t = 0; t < Q.nbr_rank.size(); ++t) if (Q.nbr_rank[t] == r) kk = int(t);
          CHECK(kk >= 0, "coarse neighbour relation");
          CHECK(Q.send_ptr[kk + 1] - Q.send_ptr[kk] == Pl.recv_ptr[k + 1] - Pl.recv_ptr[k], "coarse counts");
          for (int t = 0; t < Pl.recv_ptr[k + 1] - Pl.recv_ptr[k]; ++t)
            cg[r][Pl.n_owned + Pl.recv_ptr[k] + t] = off[q] + Q.send_nodes[Q.send_ptr[kk] + t];
        }
      }
    // global aggregate map from the owners; ghost aggregates must agree with it
    std::map<ll, ll> agg_g;
    const int n_parts_here = (fine_dist ? P : 1);
    for (int r = 0; r < n_parts_here; ++r)
      for (int i = 0; i < lv[r][l].n; ++i) {
        CHECK(agg_g.count(gid[r][i]) == 0, "node owned twice");
        agg_g[gid[r][i]] = cg[r][lv[r][l].agg[i]];
      }
    for (int r = 0; r < P; ++r) {
      const AmgLevelHost& F = lv[r][l];
      CHECK(int(F.agg.size()) == F.ld, "agg length %zu != ld %d (level %d)", F.agg.size(), F.ld, l);
      for (int i = 0; i < F.ld; ++i)
        CHECK(cg[r][F.agg[i]] == agg_g[gid[r][i]], "level %d rank %d: aggregate of local node %d disagrees", l, r, i);
      // members
      std::vector<int> seen(F.n, 0);
      CHECK(int(F.aggptr.size()) == F.nc_rows + 1 && F.nc_rows == lv[r][l + 1].n, "aggptr size");
      for (int I = 0; I < F.nc_rows; ++I)
        for (int k = F.aggptr[I]; k < F.aggptr[I + 1]; ++k) {
          CHECK(F.agg[F.aggmem[k]] == I, "member list");
          seen[F.aggmem[k]] += 1;
        }
      for (int i = 0; i < F.n; ++i) CHECK(seen[i] == 1, "node %d listed %d times", i, seen[i]);
    }
    // reference Galerkin product
    std::map<std::pair<ll, ll>, double> Mc;
    for (auto& e : M) Mc[{agg_g[e.first.first], agg_g[e.first.second]}] += e.second;
    // the parts' coarse values
    std::vector<std::vector<double>> cv(P);
    for (int r = 0; r < P; ++r) {
      const AmgLevelHost& F = lv[r][l];
      const AmgLevelHost& C = lv[r][l + 1];
      CHECK(int(F.cptr.size()) == C.nnzb + 1, "cptr size");
      cv[r].assign(C.nnzb, 0.0);
      for (int K = 0; K < C.nnzb; ++K)
        for (int m = F.cptr[K]; m < F.cptr[K + 1]; ++m) cv[r][K] += vals[r][F.cmem[m]];
    }
    if (repl) {
      std::vector<double> sum(cv[0].size(), 0.0);
      for (int r = 0; r < P; ++r) {
        CHECK(cv[r].size() == sum.size(), "replicated nnz differ");
        for (size_t k = 0; k < sum.size(); ++k) sum[k] += cv[r][k];
      }
      for (int r = 0; r < P; ++r) cv[r] = sum;
    }
    size_t count = 0;
    for (int r = 0; r < P; ++r) {
      const AmgLevelHost& C = lv[r][l + 1];
      CHECK(int(C.rowptr.size()) == C.n + 1 && int(C.colind.size()) == C.nnzb, "coarse pattern sizes");
      for (int I = 0; I < C.n; ++I)
        for (int k = C.rowptr[I]; k < C.rowptr[I + 1]; ++k) {
          CHECK(C.colind[k] >= 0 && C.colind[k] < C.ld, "coarse column out of range");
          CHECK(k == C.rowptr[I] || C.colind[k] > C.colind[k - 1], "columns not sorted");
          auto it = Mc.find({cg[r][I], cg[r][C.colind[k]]});
          CHECK(it != Mc.end(), "level %d rank %d: entry (%d,%d) not in the global product", l + 1, r, I, C.colind[k]);
          CHECK(std::fabs(it->second - cv[r][k]) <= 1e-11 * std::fabs(it->second),
                "level %d rank %d: value %.17g vs %.17g", l + 1, r, cv[r][k], it->second);
          if (coarse_dist || r == 0) ++count;
        }
    }
    CHECK(count == Mc.size(), "level %d: %zu entries vs %zu in the global product", l + 1, count, Mc.size());
    M.swap(Mc);
    gid.swap(cg);
    vals.swap(cv);
  }
  ll n_last = 0;
  for (int r = 0; r < (lv[0][nl - 1].halo_level >= 0 ? P : 1); ++r) n_last += lv[r][nl - 1].n;
  std::printf("ok levels=%d distributed_coarse=%d replicated_at=%d coarsest_nodes=%lld parts=%d\n", nl, n_dist,
              n_repl, n_last, P);
  return 0;
}
