// Device side of the reference's data layer for the hot path: the beta-binomial attention prior
// (data.py:31-41, thresholded :137-139) and DataCollate's zero-padding of mel / gate target / prior (data.py:197-246).
//
// The reference builds the prior on a CPU DataLoader worker with one scipy.stats.betabinom.pmf call per mel frame and
// ships [B,T,L] floats to the GPU every step (19.7 MB of the 30.1 MB step input at cfg 2).  Here the padded tensor is
// produced where it is consumed:
//   prior[b, i-1, k] = C(n,k) B(k+a, n-k+b) / B(a,b),  n = P-1, a = s*i, b = s*(M+1-i)       (pmf of BetaBinomial)
// evaluated in log space with fp64 lgamma (fp32 lgamma at arguments ~1e3 would cost 1e-3 relative, the whole parity
// budget); the three lgamma terms that depend on k alone are computed once per utterance row block, the four that depend
// on the frame alone once per block.  HBM-bound: 4 B written per element, nothing read.
#include "ft_internal.h"
#include "../../include/flowtron_b200.h"

namespace ft {

constexpr int PRIOR_THREADS = 256;

__global__ void __launch_bounds__(PRIOR_THREADS)
attn_prior_kernel(const int* __restrict__ in_lens, const int* __restrict__ out_lens, int T, int L, double scaling,
                  float threshold, int frames_per_block, float* __restrict__ prior) {
    extern __shared__ double s_logc[];                 // [L] lgamma(n+1) - lgamma(k+1) - lgamma(n-k+1)
    __shared__ double s_row;                           // frame-only terms of the current row
    const int b = blockIdx.y;
    const int P = in_lens[b], Mx = out_lens[b];
    const double n = static_cast<double>(P - 1);
    for (int k = threadIdx.x; k < P; k += blockDim.x)
        s_logc[k] = lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0);
    __syncthreads();
    const int t0 = blockIdx.x * frames_per_block;
    for (int t = t0; t < t0 + frames_per_block && t < T; ++t) {
        float* dst = prior + (static_cast<long long>(b) * T + t) * L;
        if (t >= Mx) {                                 // padded frame rows are zero (data.py:222-224)
            for (int k = threadIdx.x; k < L; k += blockDim.x) dst[k] = 0.f;
            continue;
        }
        const double a = scaling * (t + 1), bb = scaling * (Mx - t);      // i = t+1: a = s*i, b = s*(M+1-i)
        if (threadIdx.x == 0) s_row = lgamma(a + bb) - lgamma(a) - lgamma(bb) - lgamma(n + a + bb);
        __syncthreads();
        const double row = s_row;
        for (int k = threadIdx.x; k < L; k += blockDim.x) {
            float v = 0.f;
            if (k < P) {
                const double lp = s_logc[k] + lgamma(k + a) + lgamma(n - k + bb) + row;
                v = static_cast<float>(exp(lp));
                if (threshold > 0.f && v < threshold) v = 0.f;
            }
            dst[k] = v;
        }
        __syncthreads();                               // s_row is rewritten by the next frame
    }
}

// mel_packed: per-utterance [n_mel, F_u] blocks (ft_mel_spectrogram's output layout) -> mel_padded [B, n_mel, T] with row i
// taken from utterance order[i]; gate_padded [B, T] = 1 from frame F-1 on (data.py:235); out_lens[i] = F.
__global__ void collate_mel_kernel(const float* __restrict__ mel_packed, const long long* __restrict__ frame_offsets,
                                   const int* __restrict__ order, int n_mel, int T, float* __restrict__ mel_padded,
                                   float* __restrict__ gate_padded, int* __restrict__ out_lens) {
    const int b = blockIdx.y, u = order[b];
    const long long f0 = frame_offsets[u];
    const int F = static_cast<int>(frame_offsets[u + 1] - f0);
    const float* src = mel_packed + f0 * n_mel;
    float* dst = mel_padded + static_cast<long long>(b) * n_mel * T;
    const int total = n_mel * T;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int m = i / T, t = i - m * T;
        dst[i] = t < F ? src[static_cast<long long>(m) * F + t] : 0.f;
    }
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t < T; t += blockDim.x) gate_padded[static_cast<long long>(b) * T + t] = t >= F - 1 ? 1.f : 0.f;
        if (threadIdx.x == 0) out_lens[b] = F;
    }
}

}  // namespace ft

extern "C" {

int ft_attn_prior(const int* in_lens, const int* out_lens, int B, int T, int L, float scaling, float threshold, float* prior,
                  void* stream) {
    using namespace ft;
    if (!in_lens || !out_lens || !prior) return ft_set_error("ft_attn_prior: NULL argument");
    if (B <= 0 || T <= 0 || L <= 0) return 0;
    if (L > 4096) return ft_set_error("ft_attn_prior: L > 4096 not supported");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int fpb = 8;
    dim3 grid((T + fpb - 1) / fpb, B);
    TimeScope ts("attn_prior", T, B, L, st);
    attn_prior_kernel<<<grid, PRIOR_THREADS, sizeof(double) * L, st>>>(in_lens, out_lens, T, L, static_cast<double>(scaling),
                                                                       threshold, fpb, prior);
    ft_count_launch(1);
    return ft_check_launch("attn_prior_kernel");
}

int ft_collate_mel(const float* mel_packed, const long long* frame_offsets, const int* order, int B, int n_mel, int T,
                   float* mel_padded, float* gate_padded, int* out_lens, void* stream) {
    using namespace ft;
    if (!mel_packed || !frame_offsets || !order || !mel_padded || !gate_padded || !out_lens) return ft_set_error("ft_collate_mel: NULL argument");
    if (B <= 0 || T <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int bx = (n_mel * T + 255) / 256;
    if (bx > 64) bx = 64;
    TimeScope ts("collate_mel", T, B, n_mel, st);
    collate_mel_kernel<<<dim3(bx, B), 256, 0, st>>>(mel_packed, frame_offsets, order, n_mel, T, mel_padded, gate_padded, out_lens);
    ft_count_launch(1);
    return ft_check_launch("collate_mel_kernel");
}

}  // extern "C"
