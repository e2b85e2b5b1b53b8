// dto_internal.hpp -- shared between the engine (CUDA) and the host layer (C++): error plumbing only.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/dto_b200.h"

namespace dto {
extern thread_local std::string g_last_error;
// records a printf-style message for dto_b200_last_error() and returns `code`
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
// frees the per-batch device buffers of a context that exceed `keep_bytes` each (used before a context goes back to the
// host layer's pool, so an idle context does not sit on gigabytes of partner-slot rows)
void trim_batch_buffers(dto_b200_ctx *ctx, size_t keep_bytes);
// batched list pairs sharing the rank structure of the problem loaded in ctx (dto_engine.cu)
int run_pair_group(dto_b200_ctx *ctx, const int32_t *slot_maps, const uint64_t *seeds, size_t G, size_t perms,
                   dto_b200_record *records_out);
bool same_rank_structure(const dto_b200_ctx *ctx, const uint32_t *ranks1, size_t n1, const uint32_t *thr1, size_t T1,
                         const uint32_t *ranks2, size_t n2, const uint32_t *thr2, size_t T2, uint64_t population);
bool identical_gene_sets(const dto_b200_ctx *ctx);
size_t group_task_capacity(const dto_b200_ctx *ctx);  // tasks one launch should carry (0 without a problem)
}  // namespace dto
