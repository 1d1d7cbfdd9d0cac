// nw_launch.h -- host-callable launch wrappers implemented in nw_kernels.cu.  Each returns the number of CUDA
// kernels it launched (the driver sums them into nw_stats.kernel_launches).
#pragma once
#include "nw_kernels.cuh"

namespace nw {

double launch_int_peak(int mode, int* out, int iters, cudaStream_t st);
void init_kernel_attrs();  // once per device
int launch_moveset(const RegionDev* regions_dev, int n_regions, int nx_max, cudaStream_t st);
int launch_insert_roots(const unsigned char* blob, uint32_t pstride, int LS, const int32_t* len, int n_regions,
                        uint64_t* table, uint64_t tmask, cudaStream_t st);
int launch_rehash(const uint64_t* oldT, uint64_t old_cap, uint64_t* newT, uint64_t new_mask, cudaStream_t st);
int launch_expand(bool emit, const ExpandArgs& A, cudaStream_t st);
int launch_hash_insert(const ExpandArgs& A, uint32_t G, cudaStream_t st);
int launch_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* bsum_scratch, uint32_t* total_out,
                cudaStream_t st);
uint64_t scan_scratch_items(uint64_t n);
int launch_resolve(const ExpandArgs& A, uint32_t G, uint32_t* fresh_mask, uint32_t* wcount, uint32_t* owner,
                   cudaStream_t st);
int launch_gather_rank(const uint32_t* mask, const uint32_t* wprefix, const uint32_t* idx, uint32_t n, uint32_t limit,
                       uint32_t fill, uint32_t* dst, cudaStream_t st);
int launch_assign(const AssignArgs& A, cudaStream_t st);
int launch_verify_dups(const ExpandArgs& A, const VerifyTable& V, const uint32_t* fresh_mask, const uint32_t* owner, uint32_t first,
                       uint32_t end, unsigned int* mismatches, cudaStream_t st);
int launch_len_hist(const uint16_t* new_len, const uint16_t* new_rid, const RegionDev* regions, int lcap, uint32_t n,
                    int force, uint32_t* hist, cudaStream_t st);
constexpr uint32_t BINS_SMALL_LIMIT = 8192;  // launch_bins_small handles up to this many (region, length) bins
int launch_bins_small(const uint32_t* hist, uint32_t nb, BinRec* out /* 1 + nb records */, cudaStream_t st);
int launch_bin_flags(const uint32_t* hist, uint32_t nb, uint32_t* flag, uint32_t* nsum, cudaStream_t st);
int launch_bin_compact(const uint32_t* flag, const uint32_t* rank, const uint32_t* nsum, uint32_t nb, BinRec* out,
                       cudaStream_t st);
int launch_bucket_fill(const BinRec* ent, uint32_t n, const uint32_t* hist, uint32_t* bstart, uint32_t* cursor,
                       cudaStream_t st);
int launch_bucket_scatter(const uint16_t* new_len, const uint16_t* new_rid, const RegionDev* regions, int lcap,
                          uint32_t n_new, int force, const OwnRange& own, const uint32_t* bstart, uint32_t* cursor,
                          uint32_t* work, cudaStream_t st);
int launch_max_k3(const int32_t* k3, uint32_t n, int* best, cudaStream_t st);
int generic_threads();
int launch_score_generic(const ScoreArgs& A, cudaStream_t st);
int launch_valid_flags(const int32_t* k3, uint32_t n, uint32_t* flag, cudaStream_t st);
int launch_compact_ranks(const uint32_t* flag, const uint32_t* pre, uint32_t n, uint32_t* sel, cudaStream_t st);
int launch_make_keys(const int32_t* k3, const uint16_t* new_rid, const uint32_t* flag, const uint32_t* pre, uint32_t n,
                     uint64_t* keys, uint64_t nvalid, uint64_t npad, cudaStream_t st);
int launch_take_sorted(const uint64_t* keys, uint32_t n, const uint32_t* seg_start, const uint32_t* take,
                       const uint32_t* out_start, uint32_t* sel, cudaStream_t st);
int launch_gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t n, uint32_t limit, uint32_t fill, uint32_t* dst,
                      cudaStream_t st);
int launch_qual_append(const int32_t* k3, const uint16_t* nrid, const int32_t* k3_min, int32_t id_base, uint32_t n, void* out,
                       uint32_t cap, uint32_t* counter, cudaStream_t st);
int launch_set_needed(const int32_t* gids, uint32_t n, uint8_t* needed, cudaStream_t st);
int launch_sort_u64(uint64_t* keys, uint64_t npad, cudaStream_t st);
int launch_topk_keys(const int32_t* k3, uint32_t n, uint32_t w, uint32_t* hist, uint32_t* thr, uint32_t* flag, uint32_t* tie_rank,
                     uint32_t* bsum_scratch, uint32_t* total_out, uint64_t* keys, uint32_t npad, cudaStream_t st);
int launch_rank_select(const uint64_t* keys, uint32_t n, uint32_t* sel, cudaStream_t st);
bool rank_select_fits(uint64_t n);
int launch_keys_to_sel(const uint64_t* keys, uint32_t n, uint32_t* sel, cudaStream_t st);
int launch_materialise(const MaterialiseArgs& A, cudaStream_t st);
int launch_good_flags(const int32_t* k3, uint32_t n, int32_t id_base, int k3_min, int k3_star, uint32_t quota,
                      const uint32_t* tie_rank, uint8_t* needed, cudaStream_t st);
int launch_tie_flags(const int32_t* k3, uint32_t n, int k3_star, uint32_t* flag, cudaStream_t st);
int launch_k3_hist(const int32_t* k3, uint32_t n, int k3_floor, uint32_t* hist, cudaStream_t st);
int launch_edge_append(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                       const uint8_t* needed, const uint32_t* nrank, int lvl, int maxdepth, uint32_t gbase, void* out,
                       uint32_t cap, uint32_t* counter, cudaStream_t st);
int launch_id_write(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                    const uint8_t* needed, const uint32_t* nrank, void* out, cudaStream_t st);
int launch_needed_flags(const uint8_t* needed, uint32_t n, uint32_t* flag, cudaStream_t st);
int launch_id_append(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                     const uint8_t* needed, void* out, uint32_t cap, uint32_t* counter, cudaStream_t st);
int launch_mark_parents(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                        uint8_t* needed, int pass, int lvl, int maxdepth, cudaStream_t st);
int launch_edge_flags(const int32_t* cand_id, uint32_t G, const uint8_t* needed, uint32_t* flag, int lvl, int maxdepth,
                      cudaStream_t st);
int launch_edge_compact(const int32_t* cand_id, const uint64_t* desc, const int32_t* front_id, uint32_t G,
                        const uint32_t* flag, const uint32_t* pre, uint32_t gbase, void* out, cudaStream_t st);
int launch_id_flags(const uint8_t* needed, int32_t id_base, uint32_t n, uint32_t* flag, cudaStream_t st);
int launch_id_compact(const int32_t* k3, const int32_t* cost, const int32_t* total, int32_t id_base, uint32_t n,
                      const uint32_t* flag, const uint32_t* pre, void* out, cudaStream_t st);

}  // namespace nw
