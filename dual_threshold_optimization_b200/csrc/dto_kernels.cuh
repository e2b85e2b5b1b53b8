// dto_kernels.cuh -- launcher interface between the engine (dto_engine.cu) and the kernels (dto_kernels.cu)
#pragma once

#include "dto_device.cuh"

namespace dto {

#ifndef DTO_SCAN_THREADS
#define DTO_SCAN_THREADS 384
#endif
constexpr int kScanThreads = DTO_SCAN_THREADS;  // max threads per CTA of the scan kernel (12 warps = 12 permutations)
constexpr int kScanCtasPerSm = 2;   // occupancy the scan kernel is compiled for: 2 x 12 warps -> 80 registers/thread.  The
                                    // row loop is a function of its own (scan_rows), so ptxas gives it a private
                                    // allocation: its column state stays in registers down to 64 registers/thread.
                                    // Measured per 100 000 permutations at N = 20 000: 20 / 24 / 26 / 28 warps per SM =
                                    // 19.4 / 18.3 / 18.7 / 18.5 ms (with the loop inlined in the kernel, 24 warps spilled
                                    // the column state and ran at 24 ms)
constexpr int kSigmaThreads = 1024; // one CTA per permutation in the pairing kernel
constexpr int kCandCap = 64;        // per-warp shared-memory candidate buffer (entries)
constexpr int kTaskStatWords = 8;   // per-task diagnostics words (option task_stats)
constexpr int kChunk = 128;         // positions per cp.async chunk (8 B per lane)
constexpr int kRing = 4 * kChunk;   // per-warp staging ring of partner slots (u16)

int pick_ch(int T2);
int pick_bucket_bits(uint32_t n);
size_t scan_smem_bytes(int CH, int T1, int warps);
size_t sigma_smem_bytes(const Problem &P, int B1, int B2, bool ba_in_smem);
size_t sigma_scratch_words(const Problem &P);

cudaError_t launch_build_kcrit(const Problem &P, uint16_t *kcrit, uint32_t *counts, uint2 *meta, cudaStream_t st);
cudaError_t launch_scan_counts(uint32_t *counts, int n, unsigned long long *total_out, cudaStream_t st);
cudaError_t launch_fill_lptab(const Problem &P, const uint32_t *offsets, uint2 *meta, double *lptab, cudaStream_t st);
cudaError_t launch_sigma_sort(const Problem &P, uint64_t seed, const uint64_t *seeds, uint32_t seg, uint64_t first_id,
                              int n_tasks, uint16_t *pb, uint32_t *pairing_out, uint32_t *scratch, size_t smem_limit,
                              int sm_count, int ctas_per_sm, cudaStream_t st);
cudaError_t launch_compose(const Problem &P, const uint32_t *perm1, const uint32_t *perm2, const int32_t *slot2_maps,
                           int n_tasks, uint32_t *inv_scratch, int *err_flag, uint16_t *pb, cudaStream_t st);
cudaError_t launch_scan(const Problem &P, const uint16_t *pb, int n_tasks, int n_plain, uint32_t flags, dto_b200_record *out,
                        const ScanOut &status, unsigned long long *counters, uint32_t *task_stats, int grid, int warps,
                        cudaStream_t st);
cudaError_t launch_full_grid(const Problem &P, const uint16_t *pbrow, uint32_t *H, double *pv, double *logp,
                             cudaStream_t st);
cudaError_t launch_full_argmin(const Problem &P, const uint32_t *H, const double *pv, uint32_t flags,
                               dto_b200_record *out, uint32_t *best_cell, cudaStream_t st);
cudaError_t launch_full_collect(const Problem &P, const uint32_t *H, const double *pv, const uint32_t *best_cell,
                                uint32_t *count, uint2 *cells_out, cudaStream_t st);
cudaError_t launch_patch_records(const uint32_t *idx, const dto_b200_record *patch, int n, dto_b200_record *records,
                                 cudaStream_t st);
cudaError_t launch_pvalues(const double *lf, const uint64_t *N, const uint64_t *K, const uint64_t *n,
                           const uint64_t *k, size_t count, double *out, cudaStream_t st);
cudaError_t launch_fp64_probe(double *out, int blocks, int threads, int iters, cudaStream_t st);
cudaError_t launch_hbm_probe(const void *src, void *dst, size_t bytes, int blocks, cudaStream_t st);

}  // namespace dto
