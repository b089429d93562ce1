// Autoregressive inference of one flow (AR_Step.infer, flowtron.py:775-828) as ONE persistent cooperative
// kernel: the per-frame chain  attention_lstm -> Q -> score/softmax/context -> lstm x2 -> dense x2 -> conv ->
// (z - b)/exp(log_s) -> gate  runs for all T frames without returning to the host (the reference issues ~25
// launches and one host sync per frame, and re-projects K/V every frame; here K/V are projected once).
//
// Parallelisation: every matrix-vector phase is split into warp tasks of 4 weight rows (for the LSTMs: the 4
// gates of one hidden unit, so the cell update happens in the same warp); a task streams its rows from L2
// (fp16, 128-bit loads), the activations of the phase are staged once per CTA in shared memory (fp16, batch
// tile of 8), accumulation is fp32, and 9 grid-wide barriers per frame order the phases.  Same operand
// precision as the training kernels (fp16 operands, fp32 state), so forward(infer(z)) closes to ~1e-3.
#include "ptx.cuh"
#include "ft_internal.h"
#include "../../include/flowtron_b200.h"

namespace ft {

constexpr int IH = 1024, IG = 4096;
// batch tile IBT is a template parameter (1, 2, 4 or 8): B = 1 must not pay for 8 rows of staging and FMAs
constexpr int INF_THREADS = 256;
constexpr int KMAX = 1664 + 1024;      // widest phase input: [d ; h0]

struct InferParams {
    int T, B, L, M, A, E, D;
    // fp16 weights (row-major, PyTorch layouts)
    const __half *w_ih_a, *w_hh_a, *w_ih0, *w_hh0, *w_ih1, *w_hh1, *wq, *w1, *w2, *wc;
    // fp32 biases / small vectors
    const float *b_ih_a, *b_hh_a, *b_ih0, *b_hh0, *b_ih1, *b_hh1, *b1, *b2, *bc, *v, *wg, *bg;
    const float* Kp; const float* Vp;      // [L*B, A] projected once
    const float* residual;                 // [T,B,M] flow-time order
    const float* prior;                    // [B,T,L] (row i used at frame i) or null
    const float* attn_forced;              // [T,B,L] forced alignments (flowtron.py:585-588, 797): scoring is skipped
    float inv_temperature, gate_threshold;
    int has_gate;
    // outputs
    float* out;                            // [T,B,M]
    float* attn_out;                       // [T,B,L]
    int* n_frames;                         // [B]
    // state (global scratch, zero-initialised by the launcher)
    float *hA[2], *cA, *h0[2], *c0, *h1[2], *c1, *xprev, *q, *e, *d, *y1, *y2;
    int* alive;                            // [B] 1 while the sample is still generating
    int* barrier;                          // monotonic grid barrier counter
    int* status;
};

// Grid barrier: one release-add per CTA on a monotonic counter, one acquiring poller per CTA.  The release (gpu scope) after the
// CTA barrier is cumulative over every thread's earlier writes, so no separate __threadfence() is needed (r1 had one: a second
// full fence on the critical path of each of the 9 barriers per frame).
__device__ __forceinline__ void grid_sync(const InferParams& p, int& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        red_release_add(p.barrier, 1);
        wait_flag_ge(p.barrier, epoch * static_cast<int>(gridDim.x), p.status, 301);
    }
    __syncthreads();
}

// stage x[b][0:K] (fp32 global, row pitch ld, batch rows b0..b0+IBT) as fp16 into smem at column offset c0
template <int IBT>
__device__ __forceinline__ void stage_x(__half* sx, int c0, const float* src, long long ld, int K, int b0, int B) {
    for (int i = threadIdx.x; i < IBT * K; i += INF_THREADS) {
        const int bb = i / K, k = i % K;
        const int b = b0 + bb;
        sx[bb * KMAX + c0 + k] = __float2half_rn(b < B ? src[static_cast<long long>(b) * ld + k] : 0.f);
    }
}

// acc[r][bb] += sum_k W[rows[r], k] * x[bb][c0 + k]   (one warp; K multiple of 8)
template <int IBT>
__device__ __forceinline__ void rows4_dot(float (&acc)[4][IBT], const __half* W, int ldw, const int (&rows)[4], int K,
                                          const __half* sx, int c0, int lane) {
    for (int k = lane * 8; k < K; k += 256) {
        float wf[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint4 pk = __ldg(reinterpret_cast<const uint4*>(W + static_cast<long long>(rows[r]) * ldw + k));
            const __half2* h = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                wf[r][2 * j] = f.x; wf[r][2 * j + 1] = f.y;
            }
        }
#pragma unroll
        for (int bb = 0; bb < IBT; ++bb) {
            const uint4 pk = *reinterpret_cast<const uint4*>(sx + bb * KMAX + c0 + k);
            const __half2* h = reinterpret_cast<const __half2*>(&pk);
            float xf[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                xf[2 * j] = f.x; xf[2 * j + 1] = f.y;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[r][bb] = fmaf(wf[r][j], xf[j], acc[r][bb]);
        }
    }
}

template <int IBT>
__device__ __forceinline__ void reduce_acc(float (&acc)[4][IBT]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int bb = 0; bb < IBT; ++bb) {
            float x = acc[r][bb];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            acc[r][bb] = x;
        }
}

// One LSTM layer step for all hidden units: x = [xa (Ka) ; xb (IH)], weights W_ih [4H,Ka], W_hh [4H,IH]
template <int IBT>
__device__ void lstm_phase(const InferParams& p, __half* sx, const __half* W_ih, int Ka, const float* xa, long long lda,
                           const __half* W_hh, const float* hprev, const float* b_ih, const float* b_hh, float* c,
                           float* hnew) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (INF_THREADS / 32) + warp, nw = gridDim.x * (INF_THREADS / 32);
    for (int b0 = 0; b0 < p.B; b0 += IBT) {
        __syncthreads();
        stage_x<IBT>(sx, 0, xa, lda, Ka, b0, p.B);
        stage_x<IBT>(sx, Ka, hprev, IH, IH, b0, p.B);
        __syncthreads();
        for (int u = gw; u < IH; u += nw) {
            float acc[4][IBT];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int bb = 0; bb < IBT; ++bb) acc[r][bb] = 0.f;
            const int rows[4] = {u, IH + u, 2 * IH + u, 3 * IH + u};
            rows4_dot<IBT>(acc, W_ih, Ka, rows, Ka, sx, 0, lane);
            rows4_dot<IBT>(acc, W_hh, IH, rows, IH, sx, Ka, lane);
            reduce_acc<IBT>(acc);
            if (lane < IBT && b0 + lane < p.B) {
                const int b = b0 + lane;
                float a[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float s = 0.f;
#pragma unroll
                    for (int bb = 0; bb < IBT; ++bb) s = (bb == lane) ? acc[r][bb] : s;
                    a[r] = s + b_ih[rows[r]] + b_hh[rows[r]];
                }
                const float gi = sigmoid_f(a[0]), gf = sigmoid_f(a[1]), gg = tanh_f(a[2]), go = sigmoid_f(a[3]);
                const float cn = gf * c[b * IH + u] + gi * gg;
                c[b * IH + u] = cn;
                hnew[b * IH + u] = go * tanh_f(cn);
            }
        }
    }
}

// y[b, r] = act(W[r,:] x[b,:] + bias[r]) for r < R (R multiple of 4), K multiple of 8
template <int IBT>
__device__ void dense_phase(const InferParams& p, __half* sx, const __half* W, int K, int R, const float* x, long long ldx,
                            const float* bias, int act, float* y, long long ldy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (INF_THREADS / 32) + warp, nw = gridDim.x * (INF_THREADS / 32);
    for (int b0 = 0; b0 < p.B; b0 += IBT) {
        __syncthreads();
        stage_x<IBT>(sx, 0, x, ldx, K, b0, p.B);
        __syncthreads();
        for (int t = gw; t < R / 4; t += nw) {
            float acc[4][IBT];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int bb = 0; bb < IBT; ++bb) acc[r][bb] = 0.f;
            const int rows[4] = {4 * t, 4 * t + 1, 4 * t + 2, 4 * t + 3};
            rows4_dot<IBT>(acc, W, K, rows, K, sx, 0, lane);
            reduce_acc<IBT>(acc);
            if (lane < IBT && b0 + lane < p.B) {
                const int b = b0 + lane;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float s = 0.f;
#pragma unroll
                    for (int bb = 0; bb < IBT; ++bb) s = (bb == lane) ? acc[r][bb] : s;
                    s += bias ? bias[rows[r]] : 0.f;
                    if (act) s = tanh_f(s);
                    y[static_cast<long long>(b) * ldy + rows[r]] = s;
                }
            }
        }
    }
}

template <int IBT>
__global__ void __launch_bounds__(INF_THREADS, 1)
infer_kernel(InferParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __half* sx = reinterpret_cast<__half*>(smem_raw);          // [IBT][KMAX] fp16
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (INF_THREADS / 32) + warp, nw = gridDim.x * (INF_THREADS / 32);
    int epoch = 0;

    for (int i = 0; i < p.T; ++i) {
        const int cur = i & 1, prv = cur ^ 1;
        // all samples stopped?  (alive is only written in phase 3 of the previous frame, ordered by grid barriers)
        bool any = false;
        for (int b = 0; b < p.B; ++b) any |= (ld_acquire(&p.alive[b]) != 0);
        if (!any) break;

        // ---- P1 attention_lstm step on the previous output frame (zeros at i == 0)
        lstm_phase<IBT>(p, sx, p.w_ih_a, p.M, p.xprev, p.M, p.w_hh_a, p.hA[prv], p.b_ih_a, p.b_hh_a, p.cA, p.hA[cur]);
        grid_sync(p, epoch);
        // forced alignments (`attn` given, flowtron.py:585-588): no query / score / softmax / prior, context = attn . V
        const bool forced = p.attn_forced != nullptr;
        // ---- P2a query projection (no bias)
        if (!forced) {
        dense_phase<IBT>(p, sx, p.wq, IH, p.A, p.hA[cur], IH, nullptr, 0, p.q, p.A);
        grid_sync(p, epoch);
        }
        // ---- P2b scores e[b,l] = v . tanh(q[b] + K[l,b]) / temperature   (no key mask in inference, flowtron.py:800-803)
        for (int t = gw; t < p.B * p.L && !forced; t += nw) {
            const int b = t / p.L, l = t % p.L;
            const float* kr = p.Kp + (static_cast<long long>(l) * p.B + b) * p.A;
            float s = 0.f;
            for (int a = lane; a < p.A; a += 32) s = fmaf(p.v[a], tanh_f(p.q[b * p.A + a] + kr[a]), s);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) p.e[b * p.L + l] = s * p.inv_temperature;
        }
        if (!forced) grid_sync(p, epoch);
        // ---- P2c softmax (+ prior posterior) and context; d = [hA ; ctx]
        {
            const int chunks = (p.A + 31) / 32;
            for (int t = gw; t < p.B * (chunks + 1); t += nw) {
                const int b = t / (chunks + 1), ch = t % (chunks + 1);
                // softmax over L in registers: lane holds l = lane + 32 j
                float w[8];
                float m = -INFINITY;
                if (forced) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int l = lane + 32 * j;
                        w[j] = (l < p.L) ? p.attn_forced[(static_cast<long long>(i) * p.B + b) * p.L + l] : 0.f;
                    }
                } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int l = lane + 32 * j;
                    w[j] = (l < p.L) ? p.e[b * p.L + l] : -INFINITY;
                    m = fmaxf(m, w[j]);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { w[j] = (lane + 32 * j < p.L) ? expf(w[j] - m) : 0.f; s += w[j]; }
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float inv = 1.f / s;
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] *= inv;
                }
                if (p.prior && !forced) {
                    float m2 = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int l = lane + 32 * j;
                        if (l < p.L) {
                            const float pr = p.prior[(static_cast<long long>(b) * p.T + i) * p.L + l];
                            w[j] = logf(w[j] + 1e-20f) + logf(pr + 1e-20f);
                            m2 = fmaxf(m2, w[j]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
                    float s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { w[j] = (lane + 32 * j < p.L) ? expf(w[j] - m2) : 0.f; s2 += w[j]; }
#pragma unroll
                    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    const float inv2 = 1.f / s2;
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] *= inv2;
                }
                if (ch == chunks) {                      // this task publishes the attention weights and the hA half of d
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int l = lane + 32 * j;
                        if (l < p.L) p.attn_out[(static_cast<long long>(i) * p.B + b) * p.L + l] = w[j];
                    }
                    for (int k = lane; k < IH; k += 32) p.d[b * p.D + k] = p.hA[cur][b * IH + k];
                } else {
                    const int a = ch * 32 + lane;
                    float c = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int lmax = min(32, p.L - 32 * j);                 // warp-uniform
                        for (int ll = 0; ll < lmax; ++ll) {
                            const float wl = __shfl_sync(0xffffffffu, w[j], ll);
                            if (a < p.A) c = fmaf(wl, p.Vp[(static_cast<long long>(32 * j + ll) * p.B + b) * p.A + a], c);
                        }
                    }
                    if (a < p.A) p.d[b * p.D + IH + a] = c;
                }
            }
        }
        grid_sync(p, epoch);
        // ---- P3 lstm layer 0 on d; the gate decision for this frame rides along (one warp per sample)
        if (p.has_gate) {
            for (int b = gw; b < p.B; b += nw) {
                float s = 0.f;
                for (int k = lane; k < p.D; k += 32) s = fmaf(p.wg[k], p.d[b * p.D + k], s);
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && p.alive[b]) {
                    p.n_frames[b] = i + 1;                               // the frame that trips the gate IS emitted (:823-826)
                    if (sigmoid_f(s + p.bg[0]) > p.gate_threshold) p.alive[b] = 0;
                }
            }
        } else if (gw == 0 && lane < p.B) {
            p.n_frames[lane] = i + 1;
        }
        lstm_phase<IBT>(p, sx, p.w_ih0, p.D, p.d, p.D, p.w_hh0, p.h0[prv], p.b_ih0, p.b_hh0, p.c0, p.h0[cur]);
        grid_sync(p, epoch);
        // ---- P4 lstm layer 1
        lstm_phase<IBT>(p, sx, p.w_ih1, IH, p.h0[cur], IH, p.w_hh1, p.h1[prv], p.b_ih1, p.b_hh1, p.c1, p.h1[cur]);
        grid_sync(p, epoch);
        // ---- P5/P6 dense layers
        dense_phase<IBT>(p, sx, p.w1, IH, IH, p.h1[cur], IH, p.b1, 1, p.y1, IH);
        grid_sync(p, epoch);
        dense_phase<IBT>(p, sx, p.w2, IH, IH, p.y1, IH, p.b2, 1, p.y2, IH);
        grid_sync(p, epoch);
        // ---- P7 conv + inverse affine: out = (residual - b) / exp(log_s)
        for (int b0 = 0; b0 < p.B; b0 += IBT) {
            __syncthreads();
            stage_x<IBT>(sx, 0, p.y2, IH, IH, b0, p.B);
            __syncthreads();
            for (int t = gw; t < p.M / 2; t += nw) {
                float acc[4][IBT];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int bb = 0; bb < IBT; ++bb) acc[r][bb] = 0.f;
                const int rows[4] = {2 * t, 2 * t + 1, p.M + 2 * t, p.M + 2 * t + 1};
                rows4_dot<IBT>(acc, p.wc, IH, rows, IH, sx, 0, lane);
                reduce_acc<IBT>(acc);
                if (lane < IBT && b0 + lane < p.B) {
                    const int b = b0 + lane;
                    float o[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        float s = 0.f;
#pragma unroll
                        for (int bb = 0; bb < IBT; ++bb) s = (bb == lane) ? acc[r][bb] : s;
                        o[r] = s + p.bc[rows[r]];
                    }
                    const long long ro = (static_cast<long long>(i) * p.B + b) * p.M;
                    // a sample that stopped at an earlier frame emits zeros; the frame that trips the gate is
                    // still emitted, which n_frames (written in P3, ordered by the barriers since) encodes
                    const bool emit = (i < p.n_frames[b]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int m = 2 * t + e;
                        const float val = (p.residual[ro + m] - o[2 + e]) / expf(o[e]);
                        p.out[ro + m] = emit ? val : 0.f;
                        p.xprev[b * p.M + m] = val;
                    }
                }
            }
        }
        grid_sync(p, epoch);
    }
}

// ------------------------------------------------------------------------------------------------ host
struct InferScratch {
    uint16_t *w_ih_a, *w_hh_a, *w_ih0, *w_hh0, *w_ih1, *w_hh1, *wq, *wk, *wv, *w1, *w2, *wc, *text16;
    float *Kp, *Vp, *state;
    int* ints;
    size_t state_floats, total;
};

static InferScratch plan_infer(const FtArStepDesc& d, uint8_t* base) {
    InferScratch s;
    size_t off = 0;
    auto get = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return base ? base + o : nullptr; };
    const int M = d.n_mel, A = d.n_attn, E = d.n_text, D = IH + A;
    s.w_ih_a = reinterpret_cast<uint16_t*>(get(size_t(IG) * M * 2));
    s.w_hh_a = reinterpret_cast<uint16_t*>(get(size_t(IG) * IH * 2));
    s.w_ih0 = reinterpret_cast<uint16_t*>(get(size_t(IG) * D * 2));
    s.w_hh0 = reinterpret_cast<uint16_t*>(get(size_t(IG) * IH * 2));
    s.w_ih1 = reinterpret_cast<uint16_t*>(get(size_t(IG) * IH * 2));
    s.w_hh1 = reinterpret_cast<uint16_t*>(get(size_t(IG) * IH * 2));
    s.wq = reinterpret_cast<uint16_t*>(get(size_t(A) * IH * 2));
    s.wk = reinterpret_cast<uint16_t*>(get(size_t(A) * E * 2));
    s.wv = reinterpret_cast<uint16_t*>(get(size_t(A) * E * 2));
    s.w1 = reinterpret_cast<uint16_t*>(get(size_t(IH) * IH * 2));
    s.w2 = reinterpret_cast<uint16_t*>(get(size_t(IH) * IH * 2));
    s.wc = reinterpret_cast<uint16_t*>(get(size_t(2 * M) * IH * 2));
    s.text16 = reinterpret_cast<uint16_t*>(get(size_t(d.L) * d.B * E * 2));
    s.Kp = reinterpret_cast<float*>(get(size_t(d.L) * d.B * A * 4));
    s.Vp = reinterpret_cast<float*>(get(size_t(d.L) * d.B * A * 4));
    // state: hA[2], cA, h0[2], c0, h1[2], c1 (9 x B*IH), xprev (B*M), q (B*A), e (B*L), d (B*D), y1, y2 (B*IH)
    s.state_floats = size_t(d.B) * (9 * IH + M + A + d.L + D + 2 * IH);
    s.state = reinterpret_cast<float*>(get(s.state_floats * 4));
    s.ints = reinterpret_cast<int*>(get((size_t(d.B) + 64) * 4));
    s.total = off;
    return s;
}

}  // namespace ft

extern "C" {

size_t ft_ar_step_infer_scratch_bytes(const FtArStepDesc* d) { return ft::plan_infer(*d, nullptr).total + 256; }

int ft_ar_step_infer(const FtArStepDesc* d, const FtArStepWeights* w, const float* residual, const float* text,
                     const float* attn_prior, const float* attn_forced, float gate_threshold, float* out, float* attn_out,
                     int* n_frames, void* scratch, void* stream) {
    using namespace ft;
    if (!d || !w || !residual || !text || !out || !attn_out || !n_frames || !scratch) return ft_set_error("ft_ar_step_infer: NULL argument");
    if (d->n_hidden != IH) return ft_set_error("infer: n_hidden must be 1024");
    if (d->L > 256) return ft_set_error("infer: L > 256 not supported");
    if (d->B > 64) return ft_set_error("infer: batch > 64 not supported");
    if (d->n_mel % 8 || d->n_attn % 8 || d->n_text % 8) return ft_set_error("infer: channel counts must be multiples of 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    InferScratch s = plan_infer(*d, static_cast<uint8_t*>(scratch));
    const int M = d->n_mel, A = d->n_attn, E = d->n_text, D = IH + A, B = d->B;
    const long long RL = static_cast<long long>(d->L) * B;
#define FT_TRYX(x) do { if ((x) != 0) return -1; } while (0)
    FT_TRYX(launch_cast(w->attn_lstm_w_ih, 2, s.w_ih_a, 0, static_cast<long long>(IG) * M, st));
    FT_TRYX(launch_cast(w->attn_lstm_w_hh, 2, s.w_hh_a, 0, static_cast<long long>(IG) * IH, st));
    FT_TRYX(launch_cast(w->lstm_w_ih0, 2, s.w_ih0, 0, static_cast<long long>(IG) * D, st));
    FT_TRYX(launch_cast(w->lstm_w_hh0, 2, s.w_hh0, 0, static_cast<long long>(IG) * IH, st));
    FT_TRYX(launch_cast(w->lstm_w_ih1, 2, s.w_ih1, 0, static_cast<long long>(IG) * IH, st));
    FT_TRYX(launch_cast(w->lstm_w_hh1, 2, s.w_hh1, 0, static_cast<long long>(IG) * IH, st));
    FT_TRYX(launch_cast(w->att_query, 2, s.wq, 0, static_cast<long long>(A) * IH, st));
    FT_TRYX(launch_cast(w->att_key, 2, s.wk, 0, static_cast<long long>(A) * E, st));
    FT_TRYX(launch_cast(w->att_value, 2, s.wv, 0, static_cast<long long>(A) * E, st));
    FT_TRYX(launch_cast(w->dense_w0, 2, s.w1, 0, static_cast<long long>(IH) * IH, st));
    FT_TRYX(launch_cast(w->dense_w1, 2, s.w2, 0, static_cast<long long>(IH) * IH, st));
    FT_TRYX(launch_cast(w->conv_w, 2, s.wc, 0, static_cast<long long>(2 * M) * IH, st));
    FT_TRYX(launch_cast(text, 2, s.text16, 0, RL * E, st));
    {   // K / V projections once per utterance (the reference recomputes them every frame, flowtron.py:568-570)
        GemmArgs g;
        g.M = static_cast<int>(RL); g.N = A; g.K = E; g.A = s.text16; g.lda = E; g.a_fmt = FMT_F16;
        g.B = s.wk; g.ldb = E; g.b_fmt = FMT_F16; g.C32 = s.Kp; g.ldc32 = A;
        FT_TRYX(launch_gemm(g, st));
        g.B = s.wv; g.C32 = s.Vp;
        FT_TRYX(launch_gemm(g, st));
    }
    if (cudaMemsetAsync(s.state, 0, s.state_floats * 4, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(s.ints, 0, (static_cast<size_t>(B) + 64) * 4, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(out, 0, sizeof(float) * d->T * B * M, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(attn_out, 0, sizeof(float) * d->T * B * d->L, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(n_frames, 0, sizeof(int) * B, st) != cudaSuccess) return ft_set_error("infer: memset failed");

    InferParams p;
    p.T = d->T; p.B = B; p.L = d->L; p.M = M; p.A = A; p.E = E; p.D = D;
    p.w_ih_a = reinterpret_cast<const __half*>(s.w_ih_a); p.w_hh_a = reinterpret_cast<const __half*>(s.w_hh_a);
    p.w_ih0 = reinterpret_cast<const __half*>(s.w_ih0); p.w_hh0 = reinterpret_cast<const __half*>(s.w_hh0);
    p.w_ih1 = reinterpret_cast<const __half*>(s.w_ih1); p.w_hh1 = reinterpret_cast<const __half*>(s.w_hh1);
    p.wq = reinterpret_cast<const __half*>(s.wq); p.w1 = reinterpret_cast<const __half*>(s.w1);
    p.w2 = reinterpret_cast<const __half*>(s.w2); p.wc = reinterpret_cast<const __half*>(s.wc);
    p.b_ih_a = w->attn_lstm_b_ih; p.b_hh_a = w->attn_lstm_b_hh; p.b_ih0 = w->lstm_b_ih0; p.b_hh0 = w->lstm_b_hh0;
    p.b_ih1 = w->lstm_b_ih1; p.b_hh1 = w->lstm_b_hh1; p.b1 = w->dense_b0; p.b2 = w->dense_b1; p.bc = w->conv_b;
    p.v = w->att_v; p.wg = w->gate_w; p.bg = w->gate_b;
    p.Kp = s.Kp; p.Vp = s.Vp; p.residual = residual; p.prior = d->has_prior ? attn_prior : nullptr;
    p.attn_forced = attn_forced;
    p.inv_temperature = 1.0f / d->temperature; p.gate_threshold = gate_threshold;
    p.has_gate = d->has_gate && w->gate_w;
    p.out = out; p.attn_out = attn_out; p.n_frames = n_frames;
    float* f = s.state;
    auto take = [&](size_t n) { float* r = f; f += n; return r; };
    const size_t BH = static_cast<size_t>(B) * IH;
    p.hA[0] = take(BH); p.hA[1] = take(BH); p.cA = take(BH);
    p.h0[0] = take(BH); p.h0[1] = take(BH); p.c0 = take(BH);
    p.h1[0] = take(BH); p.h1[1] = take(BH); p.c1 = take(BH);
    p.xprev = take(static_cast<size_t>(B) * M); p.q = take(static_cast<size_t>(B) * A); p.e = take(static_cast<size_t>(B) * d->L);
    p.d = take(static_cast<size_t>(B) * D); p.y1 = take(BH); p.y2 = take(BH);
    p.alive = s.ints; p.barrier = s.ints + B + 32;
    p.status = ft_status_word();
    // alive[b] = 1
    {
        static int ones[64];
        for (int i = 0; i < 64; ++i) ones[i] = 1;
        if (cudaMemcpyAsync(p.alive, ones, sizeof(int) * B, cudaMemcpyHostToDevice, st) != cudaSuccess) return ft_set_error("infer: memcpy failed");
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ibt = B == 1 ? 1 : (B == 2 ? 2 : (B <= 4 ? 4 : 8));
    void* fn = ibt == 1 ? reinterpret_cast<void*>(infer_kernel<1>) : ibt == 2 ? reinterpret_cast<void*>(infer_kernel<2>)
             : ibt == 4 ? reinterpret_cast<void*>(infer_kernel<4>) : reinterpret_cast<void*>(infer_kernel<8>);
    const int smem = ibt * KMAX * 2;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    TimeScope ts("infer", d->T, B, d->L, st);
    void* args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(INF_THREADS), args, smem, st);
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("infer_kernel");
}

}  // extern "C"
