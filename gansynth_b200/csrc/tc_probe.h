/* Development self-test of the tcgen05 operand addressing (csrc/tc_probe.cu), built into libgansynth_b200_probe.so --
 * NOT part of the product ABI (include/gansynth_b200.h).  One 128 x n bf16 UMMA accumulation from operands staged in
 * the SWIZZLE_NONE core-matrix layout of the tensor-core convolutions; tests/test_tc_gpu.py, tools/mma_timing.py. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
int gs_tc_probe(const float* a, const float* b, float* d, int k, int n, int rows_a, int rows_b, int shift, int gstride,
                int mode, void* stream);
int gs_tc_probe_time(const float* a, const float* b, float* d, int k, int n, int rows_a, int rows_b, int shift,
                     int gstride, int mode, int reps, long long* cycles, void* stream);
#ifdef __cplusplus
}
#endif
