// Partition communication for one-context-per-GPU runs.  Replaces, for this path, the
// OWNED<->GHOST Tpetra import/export of the reference (src/linear_alg.cpp:53-118) and its PCU
// reductions (PCU_Add_Double / PCU_Add_Int on J, gradients and the local-solve status).
//
// Because every part assembles the complete rows of its owned nodes (halo elements are evaluated
// redundantly), only two operations exist:
//   halo copy   ghost entries of a nodal vector <- the owner's values (pack kernel -> neighbour
//               exchange -> received straight into the contiguous ghost range of each neighbour)
//   allreduce   fp64 sums of a few scalars
// Transports:
//   NCCL        ncclSend/ncclRecv grouped per neighbour + ncclAllReduce, enqueued on the context's
//               stream (NVLink 5 / NVSwitch).  libnccl.so.2 is bound at run time with dlopen so
//               that the library already loaded in the process (e.g. torch's) is the one used.
//   host-staged pack -> pinned host -> caller's exchange function (MPI / gloo) -> ghost range.
//               This is what a calibr8 MPI rank without NCCL binds (PCU/MPI callbacks).
//   NVLink push EXPERIMENTAL, off unless C8_P2P=1 (written at the end of round 1, not yet run on a
//               multi-GPU box): the halo copy as two small kernels over CUDA-IPC peer memory -- pack
//               and store straight into the neighbour's staging buffer + release a flag; wait for the
//               neighbours' flags and unpack -- in place of the pack kernel + ncclSend/Recv group
//               (~20 us per halo copy, 9 per Arnoldi iteration of a partitioned run), and the small
//               allreduces (<= 128 doubles: Gram-Schmidt dots, norms, scalars) as one single-CTA
//               push / wait / sum-in-rank-order kernel (C8_P2P_ALLREDUCE=0 keeps NCCL for those).
//               Sequence numbers live on the device, so the kernels replay inside the iteration graphs.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "c8b200.h"
#include "comm.cuh"
#include "context.cuh"

namespace c8 {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define C8_SYM(field, name)                                                \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name));             \
  if (!field) { err = std::string("dlsym ") + name; lib = nullptr; return false; }
    C8_SYM(GetUniqueId, "ncclGetUniqueId");
    C8_SYM(CommInitRank, "ncclCommInitRank");
    C8_SYM(CommDestroy, "ncclCommDestroy");
    C8_SYM(AllReduce, "ncclAllReduce");
    C8_SYM(Send, "ncclSend");
    C8_SYM(Recv, "ncclRecv");
    C8_SYM(GroupStart, "ncclGroupStart");
    C8_SYM(GroupEnd, "ncclGroupEnd");
    C8_SYM(GetErrorString, "ncclGetErrorString");
#undef C8_SYM
    return true;
  }
};
static NcclApi g_nccl;

struct Comm {
  c8_ctx* ctx = nullptr;
  // neighbour plan
  int n_nbr = 0;
  std::vector<int> nbr_rank, send_ptr, recv_ptr;  // node counts (prefix sums)
  int n_send = 0, n_recv = 0;
  int* d_send_nodes = nullptr;
  double* d_sendbuf = nullptr;   // [n_send][NBMAX]
  // NCCL transport
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
  // host-staged transport
  c8_host_exchange_fn exchange = nullptr;
  c8_host_allreduce_fn host_allreduce = nullptr;
  void* user = nullptr;
  double* h_send = nullptr;
  double* h_recv = nullptr;
  std::string last;
  // plans of the coarser multigrid levels (amg_host.hpp); level l >= 1 is levels[l - 1].  They share
  // the neighbour list and the send/staging buffers of the level-0 plan (a coarse level never sends
  // more nodes to a neighbour than the fine level does)
  HaloPlanHost plan0;
  struct Level { HaloPlanHost plan; int* d_send_nodes = nullptr; };
  std::vector<Level> levels;
  bool library_transport = false;   // NCCL or host-staged (not the caller's own c8_set_comm hooks)
  struct P2P* p2p = nullptr;        // NVLink push halo (experimental, C8_P2P=1)
  // statistics
  long long n_halo = 0, n_allreduce = 0, halo_bytes = 0;
};
static std::map<c8_ctx*, Comm> g_comm;
static const int NBMAX = 4;

__global__ void k_halo_pack(const double* __restrict__ v, const int* __restrict__ nodes,
                            double* __restrict__ out, int n, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nb) return;
  const int k = i / nb, c = i - k * nb;
  out[i] = v[size_t(__ldg(&nodes[k])) * nb + c];
}


// ---- NVLink push halo (experimental, C8_P2P=1) --------------------------------------------------
constexpr int P2P_MAXN = 16;          // neighbours per part
constexpr size_t P2P_HEADER = 1024;   // bytes before the staging area: halo flags[16] | seq @256 | done[2] @264 |
                                      // allreduce flags[16] @512 | allreduce seq @640
struct P2PPlan { int n; int send_ptr[P2P_MAXN + 1]; int recv_ptr[P2P_MAXN + 1]; };
struct P2PPeers {
  double* stage[P2P_MAXN];              // neighbour's staging area (peer mapping)
  unsigned long long* flag[P2P_MAXN];   // this part's flag inside the neighbour's header
  long long stride[P2P_MAXN];           // doubles per parity slot of the neighbour's staging area
  int dst_off[P2P_MAXN];                // node offset of this part's message there (the neighbour's level-0 recv_ptr)
  int recv_off[P2P_MAXN];               // own level-0 recv_ptr
};
constexpr int P2P_AR_MAX = 128;       // doubles per small allreduce (Gram-Schmidt dots, scalars)
struct P2PAll {                       // every rank (not only the halo neighbours), for the small allreduce
  int nranks, rank;
  double* area[P2P_MAXN];             // rank r's allreduce area [2 parities][nranks senders][P2P_AR_MAX]
  unsigned long long* flag[P2P_MAXN]; // this rank's flag in rank r's header (offset 512)
};
struct P2P {
  bool ready = false;
  bool ar_ready = false;
  char* base = nullptr;                 // own header | halo staging (2 parity slots x n_recv0 x NBMAX doubles) | allreduce area
  long long stride = 0;
  std::vector<void*> peer_base;         // opened IPC mappings, one per rank (nullptr: not opened / self)
  P2PPeers peers{};
  P2PAll all{};
};

// pack the send nodes of every neighbour straight into its staging slot (parity = seq & 1); the last
// CTA to finish releases this part's flag at every neighbour with the value seq + 1
__global__ void k_halo_push(const double* __restrict__ vec, const int* __restrict__ send_nodes, P2PPlan pl,
                            P2PPeers pe, int nb, int maxnb, const unsigned long long* seq,
                            unsigned int* done) {
  const unsigned long long s = *reinterpret_cast<const volatile unsigned long long*>(seq);
  const int total = pl.send_ptr[pl.n] * nb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int node_i = i / nb, c = i - node_i * nb;
    int k = 0;
    while (node_i >= pl.send_ptr[k + 1]) ++k;
    const int j = node_i - pl.send_ptr[k];
    pe.stage[k][(s & 1ull) * pe.stride[k] + size_t(pe.dst_off[k]) * maxnb + size_t(j) * nb + c] =
        vec[size_t(__ldg(&send_nodes[node_i])) * nb + c];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int k = 0; k < pl.n; ++k) *reinterpret_cast<volatile unsigned long long*>(pe.flag[k]) = s + 1;
    }
  }
}

// wait until every neighbour has released seq + 1, copy the staging slot into the ghost range, and
// let the last CTA advance the sequence number
__global__ void k_halo_pull(double* __restrict__ ghost, const double* stage, const unsigned long long* flags,
                            P2PPlan pl, P2PPeers pe, long long stride, int nb, int maxnb,
                            unsigned long long* seq, unsigned int* done) {
  const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(seq);
  if (threadIdx.x < pl.n) {
    const volatile unsigned long long* f = flags + threadIdx.x;
    while (*f < s + 1) { }
  }
  __syncthreads();
  __threadfence_system();
  const int total = pl.recv_ptr[pl.n] * nb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int node_i = i / nb, c = i - node_i * nb;
    int k = 0;
    while (node_i >= pl.recv_ptr[k + 1]) ++k;
    const int j = node_i - pl.recv_ptr[k];
    // the slot was written by a peer through NVLink: read it from L2, not from a stale L1 line
    ghost[size_t(node_i) * nb + c] =
        __ldcg(&stage[(s & 1ull) * stride + size_t(pe.recv_off[k]) * maxnb + size_t(j) * nb + c]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(seq) = s + 1;
    }
  }
}

// Small allreduce (n <= P2P_AR_MAX) in ONE launch of one CTA: push this part's values into every
// rank's area (parity = seq & 1), release the flags, wait for every rank's flag, sum in rank order
// (the same order everywhere, so every part gets the same bits) and advance the sequence number.
__global__ void k_allreduce_push(double* __restrict__ buf, int n, P2PAll pa, const double* my_area,
                                 const unsigned long long* my_flags, unsigned long long* seq) {
  const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(seq);
  const int t = threadIdx.x;
  const size_t slot = (s & 1ull) * size_t(pa.nranks) * P2P_AR_MAX;
  if (t < n) {
    const double v = buf[t];
    for (int r = 0; r < pa.nranks; ++r) pa.area[r][slot + size_t(pa.rank) * P2P_AR_MAX + t] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (t < pa.nranks) {
    *reinterpret_cast<volatile unsigned long long*>(pa.flag[t]) = s + 1;
    const volatile unsigned long long* f = my_flags + t;
    while (*f < s + 1) { }
  }
  __syncthreads();
  __threadfence_system();
  if (t < n) {
    double acc = 0.0;
    for (int r = 0; r < pa.nranks; ++r) acc += __ldcg(&my_area[slot + size_t(r) * P2P_AR_MAX + t]);
    buf[t] = acc;
  }
  __syncthreads();
  if (t == 0) *reinterpret_cast<volatile unsigned long long*>(seq) = s + 1;
}

// one halo copy over NCCL for any level's plan: pack -> grouped send/recv straight into the ghost range
static void halo_nccl_plan(Comm& c, const int* d_send_nodes, const std::vector<int>& send_ptr,
                           const std::vector<int>& recv_ptr, int n_owned, double* vec, int nb) {
  if (c.n_nbr == 0) return;
  cudaStream_t s = c.ctx->stream;
  const int n_send = send_ptr[c.n_nbr], n_recv = recv_ptr[c.n_nbr];
  if (c.p2p && c.p2p->ready) {
    P2P& q = *c.p2p;
    P2PPlan pl;
    pl.n = c.n_nbr;
    for (int k = 0; k <= c.n_nbr; ++k) { pl.send_ptr[k] = send_ptr[k]; pl.recv_ptr[k] = recv_ptr[k]; }
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(q.base);
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(q.base + 256);
    unsigned int* done = reinterpret_cast<unsigned int*>(q.base + 264);
    const double* stage = reinterpret_cast<const double*>(q.base + P2P_HEADER);
    auto grid = [](int work) { const int g = (work + 255) / 256; return g < 1 ? 1 : (g > 148 ? 148 : g); };
    k_halo_push<<<grid(n_send * nb), 256, 0, s>>>(vec, d_send_nodes, pl, q.peers, nb, NBMAX, seq, done);
    k_halo_pull<<<grid(n_recv * nb), 256, 0, s>>>(vec + size_t(n_owned) * nb, stage, flags, pl, q.peers, q.stride,
                                                  nb, NBMAX, seq, done + 1);
    ++c.n_halo;
    c.halo_bytes += (long long)(n_send + n_recv) * nb * 8;
    return;
  }
  if (n_send)
    k_halo_pack<<<(n_send * nb + 255) / 256, 256, 0, s>>>(vec, d_send_nodes, c.d_sendbuf, n_send, nb);
  g_nccl.GroupStart();
  for (int k = 0; k < c.n_nbr; ++k) {
    const int ns = send_ptr[k + 1] - send_ptr[k], nr = recv_ptr[k + 1] - recv_ptr[k];
    if (ns) g_nccl.Send(c.d_sendbuf + size_t(send_ptr[k]) * nb, size_t(ns) * nb, ncclDouble, c.nbr_rank[k], c.nccl, s);
    if (nr)
      g_nccl.Recv(vec + (size_t(n_owned) + recv_ptr[k]) * nb, size_t(nr) * nb, ncclDouble, c.nbr_rank[k], c.nccl, s);
  }
  ncclResult_t r = g_nccl.GroupEnd();
  if (r != ncclSuccess) c.last = g_nccl.GetErrorString(r);
  ++c.n_halo;
  c.halo_bytes += (long long)(n_send + n_recv) * nb * 8;
}

static void p2p_release(Comm& c) {
  if (!c.p2p) return;
  for (void* p : c.p2p->peer_base)
    if (p) cudaIpcCloseMemHandle(p);
  if (c.p2p->base) cudaFree(c.p2p->base);
  delete c.p2p;
  c.p2p = nullptr;
}

// Maps every neighbour's staging buffer into this process (CUDA IPC, one box) and learns where this
// part's messages go there.  The 64-byte handles and the per-rank tables travel by one ncclAllReduce
// of zero-padded doubles (every rank fills its own segment).  Any failure on any rank leaves NCCL
// send/recv in use on every rank.
static void p2p_setup(Comm& c) {
  // Every rank of the communicator takes part in BOTH collectives below, whatever happened locally: a
  // rank that cannot set up (no neighbours, too many, an allocation or IPC failure) contributes a zero
  // segment to the table and a fail flag to the agreement, and P2P is released everywhere.
  if (c.nranks < 2) return;
  cudaStream_t s = c.ctx->stream;
  bool ok = c.n_nbr >= 1 && c.n_nbr <= P2P_MAXN && c.nranks <= P2P_MAXN;
  P2P* q = new P2P();
  c.p2p = q;
  q->stride = (long long)(c.n_recv > 0 ? c.n_recv : 1) * NBMAX;
  const size_t ar_bytes = size_t(2) * c.nranks * P2P_AR_MAX * sizeof(double);
  const size_t bytes = P2P_HEADER + size_t(2) * q->stride * sizeof(double) + ar_bytes;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&q->base, bytes) != cudaSuccess) { cudaGetLastError(); q->base = nullptr; ok = false; }
  if (ok) cudaMemsetAsync(q->base, 0, bytes, s);
  if (ok && cudaIpcGetMemHandle(&mine, q->base) != cudaSuccess) { cudaGetLastError(); ok = false; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  // per rank: 64 handle bytes | stride | for every sender: recv offset + 1 | flag slot + 1
  const int W = 64 + 1 + 2 * c.nranks;
  std::vector<double> tab(size_t(W) * c.nranks, 0.0);
  if (ok) {
    double* me = tab.data() + size_t(W) * c.rank;
    const unsigned char* hb = reinterpret_cast<const unsigned char*>(&mine);
    for (int i = 0; i < 64; ++i) me[i] = double(hb[i]);
    me[64] = double(q->stride);
    for (int k = 0; k < c.n_nbr; ++k) {
      me[65 + c.nbr_rank[k]] = double(c.recv_ptr[k] + 1);
      me[65 + c.nranks + c.nbr_rank[k]] = double(k + 1);
    }
  }
  // collective 0: can every rank hold the table?  (an allreduce needs the same count everywhere, so a
  // rank without the buffer cannot enter collective 1; agree on skipping it instead)
  double* d_tab = nullptr;
  const bool have_tab = cudaMalloc(&d_tab, tab.size() * sizeof(double)) == cudaSuccess;
  if (!have_tab) { cudaGetLastError(); ok = false; }
  double pre = have_tab ? 0.0 : 1.0, *d_pre = nullptr;
  if (cudaMalloc(&d_pre, sizeof(double)) != cudaSuccess) {
    // not even one double: no collective is possible; end the job loudly rather than hang the others
    fprintf(stderr, "[c8b200] C8_P2P=1: out of device memory during set-up\n");
    std::abort();
  }
  cudaMemcpyAsync(d_pre, &pre, sizeof(double), cudaMemcpyHostToDevice, s);
  g_nccl.AllReduce(d_pre, d_pre, 1, ncclDouble, ncclSum, c.nccl, s);
  cudaMemcpyAsync(&pre, d_pre, sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  if (pre != 0.0) {   // agreed everywhere: nobody enters collective 1, nobody uses P2P
    if (d_tab) cudaFree(d_tab);
    cudaFree(d_pre);
    p2p_release(c);
    return;
  }
  // collective 1: the table (a rank that failed locally contributes a zero segment)
  cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, s);
  g_nccl.AllReduce(d_tab, d_tab, tab.size(), ncclDouble, ncclSum, c.nccl, s);
  cudaMemcpyAsync(tab.data(), d_tab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) ok = false;
  cudaFree(d_tab);
  // map every rank's buffer (the small allreduce talks to all of them, the halo to the neighbours)
  q->peer_base.assign(c.nranks, nullptr);
  std::vector<char*> base_of(c.nranks, nullptr);
  for (int r = 0; r < c.nranks && ok; ++r) {
    if (r == c.rank) { base_of[r] = q->base; continue; }
    const double* row = tab.data() + size_t(W) * r;
    if (row[64] == 0.0) { ok = false; break; }   // that rank contributed nothing: it failed locally
    cudaIpcMemHandle_t h;
    unsigned char* b = reinterpret_cast<unsigned char*>(&h);
    for (int i = 0; i < 64; ++i) b[i] = (unsigned char)(row[i] + 0.5);
    if (cudaIpcOpenMemHandle(&q->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      q->peer_base[r] = nullptr;
      ok = false;
      break;
    }
    base_of[r] = static_cast<char*>(q->peer_base[r]);
  }
  for (int k = 0; k < c.n_nbr && ok; ++k) {
    const int r = c.nbr_rank[k];
    const double* row = tab.data() + size_t(W) * r;
    const int dst_off = int(row[65 + c.rank] + 0.5) - 1, slot = int(row[65 + c.nranks + c.rank] + 0.5) - 1;
    if (dst_off < 0 || slot < 0) { ok = false; break; }   // the neighbour relation is not symmetric
    q->peers.stage[k] = reinterpret_cast<double*>(base_of[r] + P2P_HEADER);
    q->peers.flag[k] = reinterpret_cast<unsigned long long*>(base_of[r]) + slot;
    q->peers.stride[k] = (long long)(row[64] + 0.5);
    q->peers.dst_off[k] = dst_off;
    q->peers.recv_off[k] = c.recv_ptr[k];
  }
  q->all.nranks = c.nranks;
  q->all.rank = c.rank;
  for (int r = 0; r < c.nranks && ok; ++r) {
    const long long stride_r = (long long)(tab[size_t(W) * r + 64] + 0.5);
    q->all.area[r] = reinterpret_cast<double*>(base_of[r] + P2P_HEADER + size_t(2) * stride_r * sizeof(double));
    q->all.flag[r] = reinterpret_cast<unsigned long long*>(base_of[r] + 512) + c.rank;
  }
  // collective 2: every rank must agree before anyone pushes (a part that failed would never release
  // its flags)
  double flag_ok = ok ? 0.0 : 1.0;
  cudaMemcpyAsync(d_pre, &flag_ok, sizeof(double), cudaMemcpyHostToDevice, s);
  g_nccl.AllReduce(d_pre, d_pre, 1, ncclDouble, ncclSum, c.nccl, s);
  cudaMemcpyAsync(&flag_ok, d_pre, sizeof(double), cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) flag_ok = 1.0;
  cudaFree(d_pre);
  if (flag_ok != 0.0) {
    if (c.rank == 0)
      fprintf(stderr, "[c8b200] C8_P2P=1: peer mapping not available on every part, using NCCL send/recv\n");
    p2p_release(c);
    return;
  }
  q->ready = true;
  const char* e = getenv("C8_P2P_ALLREDUCE");
  q->ar_ready = !(e && e[0] == '0');
}

static void halo_nccl(void* user, double* vec, int nb) {
  Comm& c = *static_cast<Comm*>(user);
  halo_nccl_plan(c, c.d_send_nodes, c.send_ptr, c.recv_ptr, c.ctx->n_owned_nodes, vec, nb);
}

static void allreduce_nccl(void* user, double* buf, int n) {
  Comm& c = *static_cast<Comm*>(user);
  if (c.p2p && c.p2p->ar_ready && n <= P2P_AR_MAX) {
    P2P& q = *c.p2p;
    const double* my_area = reinterpret_cast<const double*>(q.base + P2P_HEADER + size_t(2) * q.stride * sizeof(double));
    k_allreduce_push<<<1, P2P_AR_MAX, 0, c.ctx->stream>>>(
        buf, n, q.all, my_area, reinterpret_cast<const unsigned long long*>(q.base + 512),
        reinterpret_cast<unsigned long long*>(q.base + 640));
    ++c.n_allreduce;
    return;
  }
  ncclResult_t r = g_nccl.AllReduce(buf, buf, size_t(n), ncclDouble, ncclSum, c.nccl, c.ctx->stream);
  if (r != ncclSuccess) c.last = g_nccl.GetErrorString(r);
  ++c.n_allreduce;
}

static void halo_host(void* user, double* vec, int nb) {
  Comm& c = *static_cast<Comm*>(user);
  if (c.n_nbr == 0) return;
  cudaStream_t s = c.ctx->stream;
  if (c.n_send) {
    k_halo_pack<<<(c.n_send * nb + 255) / 256, 256, 0, s>>>(vec, c.d_send_nodes, c.d_sendbuf, c.n_send, nb);
    cudaMemcpyAsync(c.h_send, c.d_sendbuf, size_t(c.n_send) * nb * sizeof(double), cudaMemcpyDeviceToHost, s);
  }
  cudaStreamSynchronize(s);
  c.exchange(c.user, c.h_send, c.h_recv, nb);
  if (c.n_recv)
    cudaMemcpyAsync(vec + size_t(c.ctx->n_owned_nodes) * nb, c.h_recv, size_t(c.n_recv) * nb * sizeof(double),
                    cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);  // h_recv is reused by the next call
  ++c.n_halo;
  c.halo_bytes += (long long)(c.n_send + c.n_recv) * nb * 8;
}

static void allreduce_host(void* user, double* buf, int n) {
  Comm& c = *static_cast<Comm*>(user);
  cudaStream_t s = c.ctx->stream;
  std::vector<double> h(n);
  cudaMemcpyAsync(h.data(), buf, n * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  c.host_allreduce(c.user, h.data(), n);
  cudaMemcpyAsync(buf, h.data(), n * sizeof(double), cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  ++c.n_allreduce;
}

// a coarse level through the host-staged transport: the caller's exchange function only knows the
// level-0 message sizes, so the (shorter) coarse messages travel in the leading part of each
// neighbour's level-0 slot
static void halo_host_level(Comm& c, const Comm::Level& L, double* vec, int nb) {
  if (c.n_nbr == 0) return;
  cudaStream_t s = c.ctx->stream;
  const HaloPlanHost& P = L.plan;
  const int n_send = P.n_send(), n_recv = P.n_recv();
  std::vector<double> tmp(size_t(std::max(n_send, n_recv) + 1) * nb);
  if (n_send) {
    k_halo_pack<<<(n_send * nb + 255) / 256, 256, 0, s>>>(vec, L.d_send_nodes, c.d_sendbuf, n_send, nb);
    cudaMemcpyAsync(tmp.data(), c.d_sendbuf, size_t(n_send) * nb * sizeof(double), cudaMemcpyDeviceToHost, s);
  }
  cudaStreamSynchronize(s);
  for (int k = 0; k < c.n_nbr; ++k)
    std::memcpy(c.h_send + size_t(c.send_ptr[k]) * nb, tmp.data() + size_t(P.send_ptr[k]) * nb,
                size_t(P.send_ptr[k + 1] - P.send_ptr[k]) * nb * sizeof(double));
  c.exchange(c.user, c.h_send, c.h_recv, nb);
  for (int k = 0; k < c.n_nbr; ++k)
    std::memcpy(tmp.data() + size_t(P.recv_ptr[k]) * nb, c.h_recv + size_t(c.recv_ptr[k]) * nb,
                size_t(P.recv_ptr[k + 1] - P.recv_ptr[k]) * nb * sizeof(double));
  if (n_recv)
    cudaMemcpyAsync(vec + size_t(P.n_owned) * nb, tmp.data(), size_t(n_recv) * nb * sizeof(double),
                    cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  ++c.n_halo;
  c.halo_bytes += (long long)(n_send + n_recv) * nb * 8;
}

static void drop_levels(Comm& c) {
  for (Comm::Level& L : c.levels)
    if (L.d_send_nodes) cudaFree(L.d_send_nodes);
  c.levels.clear();
}

// ---- services for the partitioned multigrid (comm.cuh)
bool comm_library_transport(c8_ctx* ctx) {
  auto it = g_comm.find(ctx);
  return it != g_comm.end() && it->second.library_transport && it->second.nranks > 1 &&
         ctx->halo_cb != nullptr && ctx->allreduce_cb != nullptr;
}
int comm_rank(c8_ctx* ctx) { auto it = g_comm.find(ctx); return it == g_comm.end() ? 0 : it->second.rank; }
int comm_nranks(c8_ctx* ctx) { auto it = g_comm.find(ctx); return it == g_comm.end() ? 1 : it->second.nranks; }
HaloPlanHost comm_plan(c8_ctx* ctx, int level) {
  Comm& c = g_comm[ctx];
  return level == 0 ? c.plan0 : c.levels[level - 1].plan;
}
int comm_add_level(c8_ctx* ctx, const HaloPlanHost& plan) {
  Comm& c = g_comm[ctx];
  Comm::Level L;
  L.plan = plan;
  if (plan.n_send()) {
    if (cudaMalloc(&L.d_send_nodes, plan.n_send() * sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpy(L.d_send_nodes, plan.send_nodes.data(), plan.n_send() * sizeof(int), cudaMemcpyHostToDevice);
  }
  c.levels.push_back(L);
  return int(c.levels.size());
}
void comm_drop_levels(c8_ctx* ctx) {
  auto it = g_comm.find(ctx);
  if (it != g_comm.end()) drop_levels(it->second);
}
void comm_halo_level(c8_ctx* ctx, int level, double* vec, int nb) {
  if (level == 0) { if (ctx->halo_cb) ctx->halo_cb(ctx->comm_user, vec, nb); return; }
  Comm& c = g_comm[ctx];
  const Comm::Level& L = c.levels[level - 1];
  if (c.nccl) halo_nccl_plan(c, L.d_send_nodes, L.plan.send_ptr, L.plan.recv_ptr, L.plan.n_owned, vec, nb);
  else halo_host_level(c, L, vec, nb);
}

void comm_release(c8_ctx* ctx) {
  auto it = g_comm.find(ctx);
  if (it == g_comm.end()) return;
  Comm& c = it->second;
  drop_levels(c);
  p2p_release(c);
  if (c.nccl && g_nccl.lib) g_nccl.CommDestroy(c.nccl);
  if (c.d_send_nodes) cudaFree(c.d_send_nodes);
  if (c.d_sendbuf) cudaFree(c.d_sendbuf);
  if (c.h_send) cudaFreeHost(c.h_send);
  if (c.h_recv) cudaFreeHost(c.h_recv);
  g_comm.erase(it);
}

}  // namespace c8

using namespace c8;

extern "C" {

int c8_set_partition(c8_ctx* ctx, int n_owned_nodes, int n_owned_elems) {
  C8_REQUIRE(ctx, n_owned_nodes >= 0 && n_owned_nodes <= ctx->n_nodes && n_owned_elems >= 0 &&
                      n_owned_elems <= ctx->n_elems, "owned counts exceed the local mesh");
  ctx->n_owned_nodes = n_owned_nodes;
  ctx->n_owned_elems = n_owned_elems;
  c8_linalg_invalidate(ctx);
  return C8_OK;
}

int c8_get_partition(c8_ctx* ctx, int* n_owned_nodes, int* n_owned_elems) {
  if (n_owned_nodes) *n_owned_nodes = ctx->n_owned_nodes;
  if (n_owned_elems) *n_owned_elems = ctx->n_owned_elems;
  return C8_OK;
}

int c8_set_comm(c8_ctx* ctx, c8_halo_fn halo, c8_allreduce_fn allreduce, void* user) {
  ctx->halo_cb = halo;
  ctx->allreduce_cb = allreduce;
  ctx->comm_user = user;
  ctx->comm_capturable = false;
  auto it = g_comm.find(ctx);
  if (it != g_comm.end()) it->second.library_transport = false;   // the caller's own hooks
  c8_linalg_invalidate(ctx);
  return C8_OK;
}

int c8_set_halo_plan(c8_ctx* ctx, int n_nbr, const int32_t* nbr_rank, const int32_t* send_ptr,
                     const int32_t* send_nodes, const int32_t* recv_ptr) {
  // validate against the arguments first; the stored plan is only touched once everything passed
  C8_REQUIRE(ctx, n_nbr >= 0, "negative neighbour count");
  C8_REQUIRE(ctx, n_nbr == 0 || (nbr_rank && send_ptr && recv_ptr), "halo plan: null array");
  int n_send = 0, n_recv = 0;
  if (n_nbr) {
    C8_REQUIRE(ctx, send_ptr[0] == 0 && recv_ptr[0] == 0, "halo plan: prefix arrays must start at 0");
    for (int k = 0; k < n_nbr; ++k) {
      C8_REQUIRE(ctx, send_ptr[k + 1] >= send_ptr[k] && recv_ptr[k + 1] >= recv_ptr[k],
                 "halo plan: prefix arrays must be non-decreasing");
      C8_REQUIRE(ctx, nbr_rank[k] >= 0, "halo plan: negative neighbour rank");
    }
    n_send = send_ptr[n_nbr];
    n_recv = recv_ptr[n_nbr];
  }
  C8_REQUIRE(ctx, n_recv == ctx->n_nodes - ctx->n_owned_nodes,
             "halo plan: received node count differs from the ghost count (c8_set_partition first)");
  C8_REQUIRE(ctx, n_send == 0 || send_nodes != nullptr, "halo plan: null send list");
  for (int i = 0; i < n_send; ++i)
    C8_REQUIRE(ctx, send_nodes[i] >= 0 && send_nodes[i] < ctx->n_owned_nodes,
               "halo plan: a send node is not owned");
  C8_CUDA(ctx, cudaSetDevice(ctx->device));
  // allocate the new buffers before dropping the old ones
  int* d_send_nodes = nullptr;
  double *d_sendbuf = nullptr, *h_send = nullptr, *h_recv = nullptr;
  auto drop_new = [&]() {
    if (d_send_nodes) cudaFree(d_send_nodes);
    if (d_sendbuf) cudaFree(d_sendbuf);
    if (h_send) cudaFreeHost(h_send);
    if (h_recv) cudaFreeHost(h_recv);
  };
  bool ok = true;
  if (n_send) {
    ok = ok && cuda_ok(ctx, cudaMalloc(&d_send_nodes, n_send * sizeof(int)), "cudaMalloc(send nodes)");
    ok = ok && cuda_ok(ctx, cudaMemcpy(d_send_nodes, send_nodes, n_send * sizeof(int), cudaMemcpyHostToDevice),
                       "cudaMemcpy(send nodes)");
    ok = ok && cuda_ok(ctx, cudaMalloc(&d_sendbuf, size_t(n_send) * NBMAX * sizeof(double)), "cudaMalloc(send buffer)");
  }
  ok = ok && cuda_ok(ctx, cudaMallocHost(&h_send, size_t(n_send + 1) * NBMAX * sizeof(double)), "cudaMallocHost");
  ok = ok && cuda_ok(ctx, cudaMallocHost(&h_recv, size_t(n_recv + 1) * NBMAX * sizeof(double)), "cudaMallocHost");
  if (!ok) { drop_new(); return C8_ERR_CUDA; }
  // commit
  Comm& c = g_comm[ctx];
  c.ctx = ctx;
  p2p_release(c);   // its staging buffer was sized from the old plan (re-created by the next c8_nccl_init)
  drop_levels(c);   // a new plan invalidates the multigrid's per-level plans
  if (c.d_send_nodes) cudaFree(c.d_send_nodes);
  if (c.d_sendbuf) cudaFree(c.d_sendbuf);
  if (c.h_send) cudaFreeHost(c.h_send);
  if (c.h_recv) cudaFreeHost(c.h_recv);
  c.d_send_nodes = d_send_nodes; c.d_sendbuf = d_sendbuf; c.h_send = h_send; c.h_recv = h_recv;
  c.n_nbr = n_nbr;
  c.nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
  c.send_ptr.assign(send_ptr, send_ptr + (n_nbr ? n_nbr + 1 : 0));
  c.recv_ptr.assign(recv_ptr, recv_ptr + (n_nbr ? n_nbr + 1 : 0));
  c.send_ptr.resize(n_nbr + 1, 0);   // n_nbr == 0: one zero entry
  c.recv_ptr.resize(n_nbr + 1, 0);
  c.n_send = n_send;
  c.n_recv = n_recv;
  c.plan0.n_owned = ctx->n_owned_nodes;
  c.plan0.nbr_rank = c.nbr_rank;
  c.plan0.send_ptr = c.send_ptr;
  c.plan0.recv_ptr = c.recv_ptr;
  c.plan0.send_nodes.assign(send_nodes, send_nodes + n_send);
  c8_linalg_invalidate(ctx);
  if (c.nccl) {   // re-plan on a live communicator: ranks are known now, and the push halo is rebuilt
    for (int k = 0; k < n_nbr; ++k)
      if (nbr_rank[k] >= c.nranks) return fail(ctx, C8_ERR_USAGE, "halo plan: neighbour rank out of range");
    const char* e = getenv("C8_P2P");
    if (e && e[0] == '1') p2p_setup(c);
  }
  return C8_OK;
}

int c8_nccl_unique_id(char* out128) {
  if (!g_nccl.load()) return C8_ERR_USAGE;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return C8_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(out128, &id, 128);
  return C8_OK;
}

int c8_nccl_init(c8_ctx* ctx, const char* id128, int rank, int nranks) {
  if (!g_nccl.load()) return fail(ctx, C8_ERR_USAGE, g_nccl.err);
  auto it = g_comm.find(ctx);
  C8_REQUIRE(ctx, it != g_comm.end(), "c8_set_halo_plan must be called before c8_nccl_init");
  Comm& c = it->second;
  C8_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int k = 0; k < c.n_nbr; ++k)
    C8_REQUIRE(ctx, c.nbr_rank[k] < nranks && c.nbr_rank[k] != rank, "halo plan: neighbour rank out of range");
  p2p_release(c);
  if (c.nccl) { g_nccl.CommDestroy(c.nccl); c.nccl = nullptr; }   // a second init replaces the communicator
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c.nccl, nranks, id, rank);
  if (r != ncclSuccess) return fail(ctx, C8_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  c.rank = rank; c.nranks = nranks;
  const int rc = c8_set_comm(ctx, &halo_nccl, &allreduce_nccl, &c);
  ctx->comm_capturable = true;
  c.library_transport = true;
  p2p_release(c);
  const char* e = getenv("C8_P2P");
  if (e && e[0] == '1') p2p_setup(c);   // experimental NVLink push halo, see the header of this file
  return rc;
}

int c8_set_comm_host(c8_ctx* ctx, c8_host_exchange_fn exchange, c8_host_allreduce_fn allreduce,
                     void* user) {
  auto it = g_comm.find(ctx);
  C8_REQUIRE(ctx, it != g_comm.end(), "c8_set_halo_plan must be called before c8_set_comm_host");
  Comm& c = it->second;
  c.exchange = exchange; c.host_allreduce = allreduce; c.user = user;
  const int rc = c8_set_comm(ctx, &halo_host, &allreduce_host, &c);
  c.library_transport = true;
  return rc;
}

int c8_set_comm_rank(c8_ctx* ctx, int rank, int nranks) {
  C8_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "rank must be in [0, nranks)");
  auto it = g_comm.find(ctx);
  C8_REQUIRE(ctx, it != g_comm.end(), "c8_set_halo_plan must be called before c8_set_comm_rank");
  it->second.rank = rank; it->second.nranks = nranks;
  c8_linalg_invalidate(ctx);
  return C8_OK;
}

int c8_halo(c8_ctx* ctx, double* vec_dev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  if (ctx->halo_cb) ctx->halo_cb(ctx->comm_user, vec_dev, ctx->kt->nb);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_halo_nb(c8_ctx* ctx, double* vec_dev, int nb) {
  C8_REQUIRE(ctx, nb >= 1 && nb <= NBMAX, "halo width must be 1..4");
  if (ctx->halo_cb) ctx->halo_cb(ctx->comm_user, vec_dev, nb);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_allreduce(c8_ctx* ctx, double* buf_dev, int n) {
  if (ctx->allreduce_cb) ctx->allreduce_cb(ctx->comm_user, buf_dev, n);
  C8_CUDA(ctx, cudaGetLastError());
  return C8_OK;
}

int c8_comm_stats(c8_ctx* ctx, int64_t* out3) {
  auto it = g_comm.find(ctx);
  if (it == g_comm.end()) { out3[0] = out3[1] = out3[2] = 0; return C8_OK; }
  out3[0] = it->second.n_halo; out3[1] = it->second.n_allreduce; out3[2] = it->second.halo_bytes;
  return C8_OK;
}

int c8_comm_p2p_active(c8_ctx* ctx) {
  auto it = g_comm.find(ctx);
  if (it == g_comm.end() || !it->second.p2p) return 0;
  return (it->second.p2p->ready ? 1 : 0) | (it->second.p2p->ar_ready ? 2 : 0);
}

void c8_comm_release(c8_ctx* ctx) { comm_release(ctx); }

}  // extern "C"
