// dto_internal.hpp -- shared between the engine (CUDA) and the host layer (C++): error plumbing only.
#pragma once

#include <string>

#include "../../include/dto_b200.h"

namespace dto {
extern thread_local std::string g_last_error;
// records a printf-style message for dto_b200_last_error() and returns `code`
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
// frees the per-batch device buffers of a context that exceed `keep_bytes` each (used before a context goes back to the
// host layer's pool, so an idle context does not sit on gigabytes of partner-slot rows)
void trim_batch_buffers(dto_b200_ctx *ctx, size_t keep_bytes);
}  // namespace dto
