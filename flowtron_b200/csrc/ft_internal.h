// Internal declarations shared by the .cu files of libflowtron_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ft {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// error plumbing (api.cu)
int ft_set_error(const char* msg);            // records msg, returns -1
int ft_check_launch(const char* what);        // cudaGetLastError -> ft_set_error
int* ft_status_word();                        // device int written by kernel watchdogs (0 = ok)
void ft_count_launch(int n);

enum { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };

struct GemmArgs {
    int M = 0, N = 0, K = 0;
    const void* A = nullptr; long long lda = 0; int a_fmt = 0; int a_mn = 0;
    const void* B = nullptr; long long ldb = 0; int b_fmt = 0; int b_mn = 0;
    const float* bias = nullptr; const float* bias2 = nullptr;
    int act = 0; int beta = 0; float alpha = 1.0f;
    float* C32 = nullptr; long long ldc32 = 0;
    void* C16 = nullptr; long long ldc16 = 0; int c16_fmt = 0;
};
int launch_gemm(const GemmArgs& g, cudaStream_t st);
int make_tmap_2d(CUtensorMap* out, const void* ptr, int fmt, long long rows, long long cols, long long ld,
                 int box_cols, int box_rows);

int launch_lstm_fwd(int T, int B, const float* xproj, const void* whh16, const int* lens, void* hseq16, long long ldh,
                    void* gates16, float* cstate, float* h32, long long ldh32, int* flags, cudaStream_t st);
int launch_lstm_bwd(int T, int B, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                    const float* cstate, const int* lens, void* dG16, int* flags, cudaStream_t st);

}  // namespace ft
