/* flowtron_b200 — C ABI of the B200-native Flowtron hot path (libflowtron_b200.so).
 *
 * The reference (NVIDIA/flowtron) has no FFI: the hot path sits behind Python nn.Modules.  These entry
 * points are what a binding for that path replaces; each cites the reference interface it stands for.
 * Conventions: all pointers are DEVICE pointers owned by the caller (torch's allocator in the shipped
 * host layer); no entry point allocates, synchronises the stream, or touches the default stream unless
 * `stream` is 0; return 0 on success, negative on error (message in ft_last_error()).  Weights are
 * passed in PyTorch's native layouts ([4H, in], gate order i,f,g,o) so checkpoints need no conversion.
 */
#ifndef FLOWTRON_B200_H
#define FLOWTRON_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element formats */
#define FT_F16 0
#define FT_BF16 1
#define FT_TF32 2   /* fp32 storage, tf32 tensor-core math */

int ft_version(void);
const char* ft_last_error(void);
/* 0 = ok; non-zero = code written by a device-side watchdog (a bounded spin-wait expired). Synchronises. */
int ft_device_status(void);
/* number of kernels this library has launched (bench.py's gpu_launches) */
long long ft_launch_count(void);
void ft_reset_launch_count(void);

/* C[M,N] = act(alpha * A[M,K] B[N,K]^T + bias + bias2) (+ C32 if beta).  tcgen05/TMA GEMM.
 * a_mn / b_mn = 1: operand is given MN-major, i.e. A points at a row-major [K, M] tensor.
 * Replaces: nn.Linear / Conv1d(k=1) / LSTM input projections (flowtron.py:278-288, 651, 654-655) and
 * their autograd dgrad / wgrad. */
int ft_gemm(int M, int N, int K, const void* A, long long lda, int a_fmt, int a_mn, const void* B, long long ldb,
            int b_fmt, int b_mn, const float* bias, const float* bias2, int act, int beta, float alpha, float* C32,
            long long ldc32, void* C16, long long ldc16, int c16_fmt, void* stream);

/* One nn.LSTM layer recurrence, hidden size 1024 (flowtron.py:654-655 `lstm` / `attention_lstm`, run via
 * run_padded_sequence :671-695).  xproj = W_ih x + b_ih + b_hh for all T*B rows (row = t*B + b), fp32
 * [T*B, 4096], gate order i,f,g,o.  whh16: fp16 copy of weight_hh [4096,1024].  Writes h (fp16, zero at
 * t >= lens[b]) to hseq16 [T*B, ldh] and, when non-null, the post-activation gates (fp16 [T*B,4096]) and
 * cell state (fp32 [T*B,1024]) needed by ft_lstm_bwd.  flags: int scratch [T*16].  B <= 128. */
int ft_lstm_fwd(int T, int B, const float* xproj, const void* whh16, const int* lens, void* hseq16, long long ldh,
                void* gates16, float* cstate, float* h32, long long ldh32, int* flags, void* stream);

/* BPTT of the same layer (autograd of nn.LSTM).  dh_ext: gradient w.r.t. the layer outputs, fp32
 * [T*B, ldd] (ignored at t >= lens[b]).  whhT16: bf16 copy of weight_hh^T [1024,4096].  Writes dG (bf16
 * [T*B,4096]), the gradient w.r.t. the gate pre-activations; dW_ih, dW_hh, db and dx are GEMMs / column
 * sums over dG done by the caller.  flags: int scratch [T*64]. */
int ft_lstm_bwd(int T, int B, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                const float* cstate, const int* lens, void* dG16, int* flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif
