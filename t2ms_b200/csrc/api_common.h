// Helpers shared by the C-ABI translation units (api.cu, train_api.cu).
#pragma once
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/t2s_b200.h"

namespace t2s_api {
extern thread_local char g_err[512];
int fail(int code, const char* fmt, const char* a = "", const char* b = "");
int ensure_init();
}  // namespace t2s_api

#define CUDA_OK(expr)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) return t2s_api::fail(T2S_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(expr)                      \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != T2S_OK) return rc_; \
    } while (0)
