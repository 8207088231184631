// Internal interface between lee.cu (split-K plan, CUDA-core kernel, reduction) and lee_tc.cu
// (tcgen05 3xTF32 kernel).
#pragma once
#include "common.cuh"

namespace sc {

constexpr int kLeeMaxSplits = 64;
constexpr int kLeeSubChunk = 512;  // cells accumulated in FP32 before folding into FP64

struct LeePlan {
  int ldt;        // padded tile-aligned gene count (multiple of 128); partial is f64 [splits][ldt][ldt]
  int splits;     // K splits
  int64_t chunk;  // cells per split (multiple of 64)
};

LeePlan lee_plan(int64_t n, int g);
int lee_reduce(const double* partial, const LeePlan& p, int g, float* L, int64_t ldl, cudaStream_t st);

bool lee_tc_supported(int64_t n, int g, int64_t lda, int64_t ldb);
size_t lee_tc_extra_workspace_bytes(int64_t n, int g);
// Runs split + tcgen05 contraction + FP64 reduction and writes L (f32 [g, ldl]).
int lee_tc_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n, int g,
                  const LeePlan& p, double* partial, void* extra_ws, float* L, int64_t ldl,
                  cudaStream_t st);

}  // namespace sc
