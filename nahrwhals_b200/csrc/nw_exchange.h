// nw_exchange.h -- the one exchange step of the frontier-sharded search (BASELINE config 3).
//
// The reference's level loop (solver.jl:667-684) is serial.  Here every rank runs the same deterministic expansion,
// dedup and selection; the banded DP of a level's fresh candidates is divided between the ranks by contiguous ranges
// of new-rank, and ONE all-gather per level completes the k3 / cost / total arrays on every rank.  Two transports:
//   * NCCL (ncclAllGather on the context's stream, no host synchronisation; multi-process or multi-device), loaded at
//     run time from libnccl.so.2 so that the library itself has no link-time dependency on it;
//   * a process-local group (one host thread per context; peer copies between the contexts' arrays), which is what
//     nw_solve_sharded_ctxs uses and how the GPU tests emulate several ranks on one device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <string>

namespace nw {

struct ExchangeError {
  std::string msg;
};

class Exchange {
 public:
  virtual ~Exchange() {}
  virtual int rank() const = 0;
  virtual int world() const = 0;
  // In-place all-gather over three int32 arrays of world * chunk items each: on return (stream order) segment p of
  // every array holds rank p's values.  This rank's own segment [rank * chunk, (rank + 1) * chunk) is the input.
  virtual void allgather3(int32_t* k3, int32_t* cost, int32_t* total, uint32_t chunk, cudaStream_t st) = 0;
  // sum of n int64 values over the ranks (host values; used once per search for the statistics)
  virtual void allreduce_sum(int64_t* vals, int n, cudaStream_t st) = 0;
  // a rank that fails tells the others (local transport) so that nobody waits forever
  virtual void abort() {}
};

// ---- NCCL transport ----
constexpr int NCCL_ID_BYTES = 128;
bool nccl_available(std::string* why);
void nccl_unique_id(void* id128);                                             // throws ExchangeError
Exchange* nccl_exchange_create(const void* id128, int rank, int world);       // collective; current device; throws
// ---- process-local transport: `world` exchanges sharing one rendezvous object ----
struct LocalGroup;
std::shared_ptr<LocalGroup> local_group_create(int world);
Exchange* local_exchange_create(std::shared_ptr<LocalGroup> g, int rank, int device);

}  // namespace nw
