// nw_exchange.cu -- transports of the sharded search's per-level all-gather (see nw_exchange.h).
#include "nw_exchange.h"

#include <dlfcn.h>

#include <condition_variable>
#include <mutex>
#include <vector>

namespace nw {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time.  Only the handful of entry points used here; the ABI of these has been stable since
// NCCL 2.0 (ncclUniqueId is 128 opaque bytes passed by value, data types / reduction ops are small enums).
// ---------------------------------------------------------------------------------------------------------------
struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_group)(void);
constexpr int NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_SUM = 0;

struct NcclApi {
  void* handle = nullptr;
  std::string why;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_error_string error_string = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  NcclApi() {
    // a process that already holds NCCL (torch.distributed) shares it: the loader resolves the soname to the loaded
    // object; otherwise the system library is searched
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      const char* e = dlerror();
      why = std::string("libnccl.so.2 cannot be loaded: ") + (e ? e : "unknown error");
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(handle, s);
      if (!p && why.empty()) why = std::string("libnccl misses ") + s;
      return p;
    };
    get_unique_id = (fn_get_unique_id)sym("ncclGetUniqueId");
    comm_init_rank = (fn_comm_init_rank)sym("ncclCommInitRank");
    comm_destroy = (fn_comm_destroy)sym("ncclCommDestroy");
    error_string = (fn_error_string)sym("ncclGetErrorString");
    all_gather = (fn_all_gather)sym("ncclAllGather");
    all_reduce = (fn_all_reduce)sym("ncclAllReduce");
    group_start = (fn_group)sym("ncclGroupStart");
    group_end = (fn_group)sym("ncclGroupEnd");
    if (!why.empty()) handle = nullptr;
  }
};
NcclApi& nccl() {
  static NcclApi api;
  return api;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    const char* e = nccl().error_string ? nccl().error_string(rc) : "?";
    throw ExchangeError{std::string(what) + ": " + e};
  }
}
void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw ExchangeError{std::string(what) + ": " + cudaGetErrorString(e)};
}

class NcclExchange final : public Exchange {
  NcclComm comm_ = nullptr;
  int rank_, world_;
  int64_t* dev_ = nullptr;
  int64_t* pin_ = nullptr;

 public:
  NcclExchange(const void* id128, int rank, int world) : rank_(rank), world_(world) {
    NcclApi& N = nccl();
    if (!N.handle) throw ExchangeError{N.why};
    NcclId id;
    memcpy(id.internal, id128, sizeof id.internal);
    nccl_check(N.comm_init_rank(&comm_, world, id, rank), "ncclCommInitRank");
    cuda_check(cudaMalloc(&dev_, 64 * sizeof(int64_t)), "cudaMalloc");
    cuda_check(cudaMallocHost(&pin_, 64 * sizeof(int64_t)), "cudaMallocHost");
  }
  ~NcclExchange() override {
    if (comm_) nccl().comm_destroy(comm_);
    if (dev_) cudaFree(dev_);
    if (pin_) cudaFreeHost(pin_);
  }
  int rank() const override { return rank_; }
  int world() const override { return world_; }
  void allgather3(int32_t* k3, int32_t* cost, int32_t* total, uint32_t chunk, cudaStream_t st) override {
    NcclApi& N = nccl();
    const size_t at = (size_t)rank_ * chunk;
    nccl_check(N.group_start(), "ncclGroupStart");  // one fused launch for the three arrays
    nccl_check(N.all_gather(k3 + at, k3, chunk, NCCL_INT32, comm_, st), "ncclAllGather");
    nccl_check(N.all_gather(cost + at, cost, chunk, NCCL_INT32, comm_, st), "ncclAllGather");
    nccl_check(N.all_gather(total + at, total, chunk, NCCL_INT32, comm_, st), "ncclAllGather");
    nccl_check(N.group_end(), "ncclGroupEnd");
  }
  void allreduce_sum(int64_t* vals, int n, cudaStream_t st) override {
    if (n > 64) throw ExchangeError{"allreduce_sum: too many values"};
    memcpy(pin_, vals, (size_t)n * 8);
    cuda_check(cudaMemcpyAsync(dev_, pin_, (size_t)n * 8, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
    nccl_check(nccl().all_reduce(dev_, dev_, (size_t)n, NCCL_INT64, NCCL_SUM, comm_, st), "ncclAllReduce");
    cuda_check(cudaMemcpyAsync(pin_, dev_, (size_t)n * 8, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    memcpy(vals, pin_, (size_t)n * 8);
  }
};

}  // namespace

bool nccl_available(std::string* why) {
  NcclApi& N = nccl();
  if (!N.handle && why) *why = N.why;
  return N.handle != nullptr;
}
void nccl_unique_id(void* id128) {
  NcclApi& N = nccl();
  if (!N.handle) throw ExchangeError{N.why};
  NcclId id;
  nccl_check(N.get_unique_id(&id), "ncclGetUniqueId");
  memcpy(id128, id.internal, sizeof id.internal);
}
Exchange* nccl_exchange_create(const void* id128, int rank, int world) { return new NcclExchange(id128, rank, world); }

// ---------------------------------------------------------------------------------------------------------------
// process-local transport: a generation barrier that can be aborted, and peer copies between the ranks' arrays
// ---------------------------------------------------------------------------------------------------------------
struct LocalGroup {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  bool failed = false;
  struct Slot {
    int32_t *k3 = nullptr, *cost = nullptr, *total = nullptr;
    int device = 0;
    std::vector<int64_t> vals;
  };
  std::vector<Slot> slot;
  explicit LocalGroup(int w) : world(w), slot(w) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) throw ExchangeError{"another rank of the sharded search failed"};
    const uint64_t g = gen;
    if (++arrived == world) {
      arrived = 0;
      ++gen;
      cv.notify_all();
      return;
    }
    cv.wait(lk, [&] { return gen != g || failed; });
    if (gen == g) throw ExchangeError{"another rank of the sharded search failed"};
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
};

namespace {
class LocalExchange final : public Exchange {
  std::shared_ptr<LocalGroup> g_;
  int rank_, device_;

 public:
  LocalExchange(std::shared_ptr<LocalGroup> g, int rank, int device) : g_(std::move(g)), rank_(rank), device_(device) {
    g_->slot[rank].device = device;
  }
  int rank() const override { return rank_; }
  int world() const override { return g_->world; }
  void abort() override { g_->fail(); }
  void allgather3(int32_t* k3, int32_t* cost, int32_t* total, uint32_t chunk, cudaStream_t st) override {
    LocalGroup& G = *g_;
    LocalGroup::Slot& me = G.slot[rank_];
    me.k3 = k3;
    me.cost = cost;
    me.total = total;
    cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize");  // own segment complete before anybody reads it
    G.barrier();
    const size_t bytes = (size_t)chunk * 4;
    for (int p = 0; p < G.world; ++p) {
      if (p == rank_) continue;
      const LocalGroup::Slot& o = G.slot[p];
      const size_t at = (size_t)p * chunk;
      cuda_check(cudaMemcpyPeerAsync(k3 + at, device_, o.k3 + at, o.device, bytes, st), "cudaMemcpyPeerAsync");
      cuda_check(cudaMemcpyPeerAsync(cost + at, device_, o.cost + at, o.device, bytes, st), "cudaMemcpyPeerAsync");
      cuda_check(cudaMemcpyPeerAsync(total + at, device_, o.total + at, o.device, bytes, st), "cudaMemcpyPeerAsync");
    }
    cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    G.barrier();  // every rank has read its peers: the arrays may be reused
  }
  void allreduce_sum(int64_t* vals, int n, cudaStream_t) override {
    LocalGroup& G = *g_;
    G.slot[rank_].vals.assign(vals, vals + n);
    G.barrier();
    for (int k = 0; k < n; ++k) {
      int64_t s = 0;
      for (int p = 0; p < G.world; ++p) s += G.slot[p].vals[k];
      vals[k] = s;
    }
    G.barrier();
  }
};
}  // namespace

std::shared_ptr<LocalGroup> local_group_create(int world) { return std::make_shared<LocalGroup>(world); }
Exchange* local_exchange_create(std::shared_ptr<LocalGroup> g, int rank, int device) {
  return new LocalExchange(std::move(g), rank, device);
}

}  // namespace nw
