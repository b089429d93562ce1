// Autoregressive inference of one flow (AR_Step.infer, flowtron.py:775-828) as ONE persistent cooperative
// kernel: the per-frame chain  attention_lstm -> Q -> score/softmax/context -> lstm x2 -> dense x2 -> conv ->
// (z - b)/exp(log_s) -> gate  runs for all T frames without returning to the host (the reference issues ~25
// launches and one host sync per frame, and re-projects K/V every frame; here K/V are projected once).
//
// Round-2 design (r1 was warp-per-4-rows CUDA-core FMAs: B=16 cost 2.9x B=1, 9 grid barriers per frame, loads issued one
// dependent iteration at a time):
//   * every matrix-vector phase runs on the TENSOR CORES with the batch as the M = 16 rows of mma.sync.m16n8k16 and 8 weight
//     rows as N, so batches up to 16 cost what B = 1 costs.  Why warp-level mma.sync and not tcgen05: a tcgen05 instruction
//     needs M >= 64 rows and BOTH operands in shared memory -- for an N <= 16 matvec that is 2.5-3 KB of shared-memory operand
//     reads (20+ clk) per K = 16 plus a shared-memory staging pass for 54 MB of weights per frame, while mma.sync takes the
//     weights straight from L2 into registers (one 16-byte load per lane per 32 k) with no padding rows.  The phase is bound
//     by L2 latency/bandwidth, not by tensor throughput, either way;
//   * weights are re-packed ONCE per call into the order the lanes consume them: [task (8 rows)][k-block of 32][row][32 k]
//     fp16, input and recurrent matrices of an LSTM fused along K (zero padded), so a warp instruction reads 512 contiguous
//     bytes and issues 8 of them back to back (4 KB in flight per warp, 16 warps per SM);
//   * a task's K range is split over the warps of a CTA (partial accumulators meet in shared memory), so all 16 warps of
//     every SM stream weights in every phase;
//   * a warp's share of the NEXT phase's weights is prefetched into registers before it waits for that phase's inputs (weights
//     do not depend on them); scores run over all warps of the grid, softmax (+ prior posterior) + context + gate of one
//     utterance inside one CTA;
//   * NO grid barriers (r2 call 7 trace: 9 barriers x 1.4 us + 9 staging round trips x 0.6 us = 18 of a frame's 39 us): the
//     DATA IS THE FLAG.  Every exchanged activation lives in a ring of 3 frame slots pre-filled with a sentinel (0xFFFF per
//     fp16, 0xFFFFFFFF per fp32 word: bit patterns no conversion produces); producers store results with st.relaxed.gpu and
//     reset the same elements of the NEXT slot to the sentinel; consumers spin with ld.relaxed.gpu on the very words they need
//     until no sentinel is left and copy them to shared memory.  A phase hand-over costs one L2 store + one L2 load instead of
//     membar + atomic + poll + barrier + load.  A slot is reset two frames after its last reader: every read is followed by a
//     store that all CTAs consume within one frame, and one gpu-scope fence per thread per frame makes that chain formal
//     (tools/trace_infer.py prints the phase table).
// Same operand precision as the training kernels (fp16 operands, fp32 accumulate/state).
#include "ptx.cuh"
#include "ft_internal.h"
#include "../../include/flowtron_b200.h"

namespace ft {

constexpr int IH = 1024, IG = 4096;
constexpr int INF_THREADS = 512, INF_WARPS = 16;
constexpr int KMAX = 1664 + 1024;              // widest phase input: [d ; h0]
constexpr int KP = KMAX + 32;                  // activation row pitch (halfs): 5440 B = 64 mod 128 -> conflict-free 16-byte reads
constexpr int XPAD = 96;                       // the 80 mel channels of the attention LSTM's input, padded to 3 k-blocks
constexpr int LMAX = 256;

struct InferParams {
    int T, B, L, M, A, E, D;
    // packed fp16 weights ([task][k-block][8 rows][32 k]) and packed fp32 biases ([task][8])
    const __half *wA, *w0, *w1, *wq, *wd1, *wd2, *wc;
    const float *bA, *b0, *b1, *bd1, *bd2, *bc;
    const float *v, *wg, *bg;
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
    // state (global scratch).  Exchanged tensors are rings of 3 frame slots (slot = frame % 3) initialised to the sentinel
    // (slot 2 of the recurrent ones to zero: the state before frame 0); cell states are private to their producing lanes.
    float *cA, *c0, *c1;
    // (slot-0 pointers; slot k of every ring lies k * slot_bytes further: no dynamically indexed arrays in the parameter struct,
    //  which would move it to local memory)
    long long slot_bytes;
    float *hA, *q, *e;                     // fp32: attention-LSTM output (gate), query, scores
    __half *hA16, *h016, *h116, *d16, *y116, *y216, *x16;   // fp16 [B][K]: what the next phases consume
    uint32_t* alive;                       // [B] per slot: 1 while the sample keeps generating after this frame, 0 once stopped
    int* status;
    long long* trace;                      // debug: [T][32] clock64 stamps of CTA 0 (tools/trace_infer.py), or null
};

static long long* g_infer_trace = nullptr;
#define IT_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[i * 32 + (slot)] = clock64(); } while (0)

constexpr uint32_t SENT = 0xFFFFFFFFu;
#define SL(ptr, k) (reinterpret_cast<decltype(ptr)>(reinterpret_cast<char*>(ptr) + (k)))     /* ring slot: k = slot index * slot_bytes */
__device__ __forceinline__ uint4 ld_rlx_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_rlx_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rlx_u32(void* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rlx_u16(void* p, uint16_t v) { asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory"); }
__device__ __forceinline__ void st_rlx_f32(float* p, float v) { st_rlx_u32(p, __float_as_uint(v)); }
__device__ __forceinline__ void st_rlx_h(__half* p, float v) { st_rlx_u16(p, __half_as_ushort(__float2half_rn(v))); }
__device__ __forceinline__ void st_rlx_h2(__half* p, float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    st_rlx_u32(p, *reinterpret_cast<const uint32_t*>(&h));
}
// any fp16 element of the 16-byte packet still the sentinel?
__device__ __forceinline__ bool sent16(const uint4& v) { return (__vcmpeq2(v.x, SENT) | __vcmpeq2(v.y, SENT) | __vcmpeq2(v.z, SENT) | __vcmpeq2(v.w, SENT)) != 0u; }
__device__ __forceinline__ uint4 poll16(const void* g, int* status, int code) {
    uint4 v = ld_rlx_v4(g);
    if (sent16(v)) {
        const long long t0 = clock64();
        do {
            v = ld_rlx_v4(g);
            if (clock64() - t0 > FT_WATCHDOG_CYCLES) watchdog_fail(status, code);
        } while (sent16(v));
    }
    return v;
}
__device__ __forceinline__ float poll_f32(const float* g, int* status, int code) {
    uint32_t v = ld_rlx_u32(g);
    if (v == SENT) {
        const long long t0 = clock64();
        do {
            v = ld_rlx_u32(g);
            if (clock64() - t0 > FT_WATCHDOG_CYCLES) watchdog_fail(status, code);
        } while (v == SENT);
    }
    return __uint_as_float(v);
}

// activations of a phase: up to two fp16 ring-slot tensors [B, pitch] (written by the producing phases' epilogues on other SMs)
// -> shared [16][KP] at columns cA / cB.  Round = every still-pending 16-byte packet of this thread is fetched with cp.async.cg
// (L2, no registers, all in flight at once), then checked in shared memory; packets that still hold a sentinel element are
// fetched again next round.  Batch rows >= B stay zero.  K, pitch and column offsets are multiples of 8.  Callers finish with
// __syncthreads().
__device__ __forceinline__ void stage2(__half* sx, int cA, const __half* srcA, int KA, int pitchA, int cB, const __half* srcB, int KB, int pitchB,
                                       int B, int* status) {
    const int k8A = KA >> 3, k8B = KB >> 3, totA = B * k8A, total = totA + B * k8B;
    uint32_t pend = 0;                                            // bit u: packet threadIdx.x + u * INF_THREADS not valid yet
    for (int u = 0, idx = threadIdx.x; idx < total; ++u, idx += INF_THREADS) pend |= 1u << u;
    long long t0 = 0;
    while (pend) {
        for (int u = 0, idx = threadIdx.x; idx < total; ++u, idx += INF_THREADS) {
            if (pend >> u & 1u) {
                const bool inA = idx < totA;
                const int j = inA ? idx : idx - totA, k8 = inA ? k8A : k8B;
                const int b = j / k8, i8 = j - b * k8;
                const __half* g = (inA ? srcA + static_cast<long long>(b) * pitchA : srcB + static_cast<long long>(b) * pitchB) + 8 * i8;
                const uint32_t dst = smem_u32(sx + b * KP + (inA ? cA : cB) + 8 * i8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int u = 0, idx = threadIdx.x; idx < total; ++u, idx += INF_THREADS) {
            if (pend >> u & 1u) {
                const bool inA = idx < totA;
                const int j = inA ? idx : idx - totA, k8 = inA ? k8A : k8B;
                const int b = j / k8, i8 = j - b * k8;
                const uint4 v = *reinterpret_cast<const uint4*>(sx + b * KP + (inA ? cA : cB) + 8 * i8);
                if (!sent16(v)) pend &= ~(1u << u);
            }
        }
        if (pend) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > FT_WATCHDOG_CYCLES) watchdog_fail(status, 302);
        }
    }
}

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// d[16 batch x 8 rows] += sum over k-blocks [m0, m1) of  x[batch][k] * W[row][k].  Wt: this task's packed weights
// ([k-block][8 rows][32 k]); lane (n = lane/4, j = lane%4) owns row n and the 8 k values 8j..8j+7 of every block; the k order
// inside a block is a fixed permutation shared by both operands (any bijection is a valid contraction order):
// MMA step s in {0,1} of a block uses halfs 4s..4s+3 of the lane's chunk as its (k, k+1, k+8, k+9) slots.
template <bool kB16>
__device__ __forceinline__ void mv_partial(const __half* __restrict__ Wt, int m0, int m1, const __half* sx, int lane, float (&d)[4]) {
    constexpr int U = 10;
    const uint4* wp = reinterpret_cast<const uint4*>(Wt) + lane;             // block m: + 32 m   (lane = 4 n + j: 16 bytes each)
    const __half* a_lo = sx + (lane >> 2) * KP + 8 * (lane & 3);
    const __half* a_hi = a_lo + 8 * KP;
    for (int m = m0; m < m1; m += U) {
        uint4 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = (m + u < m1) ? __ldg(wp + 32 * (m + u)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (m + u < m1) {
                const uint4 lo = *reinterpret_cast<const uint4*>(a_lo + 32 * (m + u));
                uint4 hi = make_uint4(0u, 0u, 0u, 0u);
                if (kB16) hi = *reinterpret_cast<const uint4*>(a_hi + 32 * (m + u));
                mma_16816(d, lo.x, hi.x, lo.y, hi.y, w[u].x, w[u].y);
                mma_16816(d, lo.z, hi.z, lo.w, hi.w, w[u].z, w[u].w);
            }
        }
    }
}

// Weights do not depend on the activations a phase waits for: the first PFN k-blocks of a warp's share of the NEXT phase are
// loaded into registers BEFORE the grid barrier and land while the CTA waits (trace: a phase spent 1-3 load round trips of
// ~1.2 us each after its barrier).
constexpr int PFN = 11;
struct Prefetch { uint4 w[PFN]; };
__device__ __forceinline__ void mv_prefetch(Prefetch& pf, const __half* __restrict__ W, int n_tasks, int nm, int S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tpc = INF_WARPS / S;
    const int tl = warp / S, split = warp - tl * S;
    const int task = blockIdx.x * tpc + tl;
    const int m0 = nm * split / S, m1 = nm * (split + 1) / S;
    const uint4* wp = reinterpret_cast<const uint4*>(W + static_cast<long long>(task) * nm * 256) + lane;
#pragma unroll
    for (int u = 0; u < PFN; ++u) pf.w[u] = (task < n_tasks && m0 + u < m1) ? __ldg(wp + 32 * (m0 + u)) : make_uint4(0u, 0u, 0u, 0u);
}

// One matrix-vector phase: n_tasks tasks of 8 rows, nm k-blocks, each task split over S warps of a CTA.  epi(task, d) is
// called by the task's first warp (all 32 lanes) with the complete accumulator fragment:
//   d[0], d[1] = batch row lane/4, rows 2j, 2j+1 of the task (j = lane%4);  d[2], d[3] = batch row lane/4 + 8, same rows.
// pf: the prefetched first PFN blocks of this warp's share of the FIRST pass (mv_prefetch with the same arguments).
template <bool kB16, class Epi>
__device__ __forceinline__ void mv_phase(const __half* __restrict__ W, int n_tasks, int nm, int S, float* spart, const __half* sx,
                                         const Prefetch& pf, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tpc = INF_WARPS / S;
    const int tl = warp / S, split = warp - tl * S;
    bool first = true;
    for (int base = blockIdx.x * tpc; base < n_tasks; base += gridDim.x * tpc) {       // CTA-uniform trip count
        const int task = base + tl;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        if (task < n_tasks) {
            int m0 = nm * split / S;
            const int m1 = nm * (split + 1) / S;
            if (first) {
                const __half* a_lo = sx + (lane >> 2) * KP + 8 * (lane & 3);
                const __half* a_hi = a_lo + 8 * KP;
#pragma unroll
                for (int u = 0; u < PFN; ++u) {
                    if (m0 + u < m1) {
                        const uint4 lo = *reinterpret_cast<const uint4*>(a_lo + 32 * (m0 + u));
                        uint4 hi = make_uint4(0u, 0u, 0u, 0u);
                        if (kB16) hi = *reinterpret_cast<const uint4*>(a_hi + 32 * (m0 + u));
                        mma_16816(d, lo.x, hi.x, lo.y, hi.y, pf.w[u].x, pf.w[u].y);
                        mma_16816(d, lo.z, hi.z, lo.w, hi.w, pf.w[u].z, pf.w[u].w);
                    }
                }
                m0 = (m0 + PFN < m1) ? m0 + PFN : m1;
            }
            mv_partial<kB16>(W + static_cast<long long>(task) * nm * 256, m0, m1, sx, lane, d);
            if (S > 1) *reinterpret_cast<float4*>(spart + (warp * 32 + lane) * 4) = make_float4(d[0], d[1], d[2], d[3]);
        }
        if (S > 1) {
            __syncthreads();
            if (task < n_tasks && split == 0) {
                for (int s2 = 1; s2 < S; ++s2) {
                    const float4 v = *reinterpret_cast<const float4*>(spart + ((warp + s2) * 32 + lane) * 4);
                    d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
                }
            }
        }
        if (task < n_tasks && split == 0) epi(task, d);
        if (S > 1) __syncthreads();
        first = false;
    }
}

// LSTM cell for task = 2 units x 4 gates (rows i0,f0,g0,o0,i1,f1,g1,o1): lane j holds (i,f) [j even] or (g,o) [j odd] of unit j/2.
// h goes to ring slot `h16` (and `hnew`, fp32, when given); the same elements of the next slot are reset to the sentinel.
__device__ __forceinline__ void lstm_epi(const InferParams& p, int task, float (&d)[4], const float* bias, float* c, float* hnew, float* hnew_next,
                                         __half* h16, __half* h16_next) {
    const int lane = threadIdx.x & 31, j = lane & 3, r = lane >> 2;
    const float b0 = bias[task * 8 + 2 * j], b1 = bias[task * 8 + 2 * j + 1];
    float v[4] = {d[0] + b0, d[1] + b1, d[2] + b0, d[3] + b1};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __shfl_xor_sync(0xffffffffu, v[i], 1);          // partner lane holds the other gate pair
    if ((j & 1) == 0) {
        const int u = 2 * task + (j >> 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int b = r + 8 * half;
            if (b < p.B) {
                const float gi = sigmoid_f(v[2 * half]), gf = sigmoid_f(v[2 * half + 1]);
                const float gg = tanh_f(o[2 * half]), go = sigmoid_f(o[2 * half + 1]);
                const float cn = gf * c[b * IH + u] + gi * gg;
                c[b * IH + u] = cn;
                const float h = go * tanh_f(cn);
                st_rlx_h(h16 + b * IH + u, h);
                st_rlx_u16(h16_next + b * IH + u, 0xFFFFu);
                if (hnew) { st_rlx_f32(hnew + b * IH + u, h); st_rlx_u32(hnew_next + b * IH + u, SENT); }
            }
        }
    }
}

// Utterance b inside CTA b: softmax (+ prior posterior) or the forced alignment, context, d = [hA ; ctx], gate decision.
// (Must stay inlined: taking the address of the kernel's parameter struct for a real call moves it to local memory and
//  turns every p.field access of the frame loop into a local load -- measured: 40 -> 57 us per frame.)
__device__ __forceinline__ void attend_one(const InferParams& p, float* sf, int i, long long cur, long long nxt, bool forced, bool was_alive) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    float* se = sf; float* sd = sf + LMAX; float* sp = sd + p.D;          // e / attn [LMAX], d [D], context partials [4][A]
    if (!forced) {
        if (warp == 0) {                         // softmax over L in registers: lane holds l = lane + 32 jj
            float w[8];
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int l = lane + 32 * jj;
                w[jj] = -INFINITY;
                if (l < p.L) w[jj] = poll_f32(SL(p.e, cur) + b * p.L + l, p.status, 303);
                mx = fmaxf(mx, w[jj]);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float s = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) { w[jj] = (lane + 32 * jj < p.L) ? expf(w[jj] - mx) : 0.f; s += w[jj]; }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float inv = 1.f / s;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) w[jj] *= inv;
            if (p.prior) {
                float m2 = -INFINITY;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int l = lane + 32 * jj;
                    if (l < p.L) {
                        const float pr = p.prior[(static_cast<long long>(b) * p.T + i) * p.L + l];
                        w[jj] = logf(w[jj] + 1e-20f) + logf(pr + 1e-20f);
                        m2 = fmaxf(m2, w[jj]);
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
                float s2 = 0.f;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) { w[jj] = (lane + 32 * jj < p.L) ? expf(w[jj] - m2) : 0.f; s2 += w[jj]; }
#pragma unroll
                for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                const float inv2 = 1.f / s2;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) w[jj] *= inv2;
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) { const int l = lane + 32 * jj; if (l < p.L) se[l] = w[jj]; }
        }
    } else {
        for (int l = threadIdx.x; l < p.L; l += INF_THREADS) se[l] = p.attn_forced[(static_cast<long long>(i) * p.B + b) * p.L + l];
    }
    // fp32 attention-LSTM output of this frame (exchanged: written by the P1 epilogues of other CTAs).  Warp 0 is busy with
    // the softmax above, the other warps wait here.
    for (int k = threadIdx.x; k < IH; k += INF_THREADS) sd[k] = poll_f32(SL(p.hA, cur) + b * IH + k, p.status, 304);
    __syncthreads();
    for (int l = threadIdx.x; l < p.L; l += INF_THREADS) p.attn_out[(static_cast<long long>(i) * p.B + b) * p.L + l] = se[l];
    const int a4n = p.A >> 2, ngrp = INF_THREADS / a4n < 4 ? INF_THREADS / a4n : 4;       // A = 640: 3 groups of 160 threads
    {   // context: `ngrp` groups of keys x (A/4) groups of 4 channels, 16-byte loads of V, partial sums meet in shared memory
        const int lg = threadIdx.x / a4n, a4 = threadIdx.x - lg * a4n;
        if (lg < ngrp) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            constexpr int CU = 6;                                     // independent 16-byte loads in flight per thread
            for (int l0 = lg; l0 < p.L; l0 += CU * ngrp) {
                float4 v[CU];
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    const int l = l0 + u * ngrp;
                    v[u] = l < p.L ? *reinterpret_cast<const float4*>(p.Vp + (static_cast<long long>(l) * p.B + b) * p.A + 4 * a4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    const int l = l0 + u * ngrp;
                    const float w = l < p.L ? se[l] : 0.f;
                    acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y); acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
                }
            }
            *reinterpret_cast<float4*>(sp + lg * p.A + 4 * a4) = acc;
        }
    }
    __syncthreads();
    for (int a = threadIdx.x; a < p.A; a += INF_THREADS) {
        float c = 0.f;
        for (int g2 = 0; g2 < ngrp; ++g2) c += sp[g2 * p.A + a];
        sd[IH + a] = c;
    }
    __syncthreads();
    for (int k = 2 * threadIdx.x; k < p.D; k += 2 * INF_THREADS) {       // d = [hA ; ctx] (fp16) for lstm layer 0 (D even)
        st_rlx_h2(SL(p.d16, cur) + b * p.D + k, sd[k], sd[k + 1]);
        st_rlx_u32(SL(p.d16, nxt) + b * p.D + k, SENT);
    }
    if (warp == 0) {                             // gate decision for this frame (the frame that trips the gate IS emitted, :823-826)
        bool alive = was_alive;
        if (p.has_gate) {
            float s = 0.f;
            for (int k = lane; k < p.D; k += 32) s = fmaf(p.wg[k], sd[k], s);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (was_alive && sigmoid_f(s + p.bg[0]) > p.gate_threshold) alive = false;
        }
        if (lane == 0) {
            if (was_alive) p.n_frames[b] = i + 1;
            st_rlx_u32(SL(p.alive, cur) + b, alive ? 1u : 0u);
            st_rlx_u32(SL(p.alive, nxt) + b, SENT);
        }
    }
}

template <bool kB16>
__global__ void __launch_bounds__(INF_THREADS, 1)
infer_kernel(InferParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __half* sx = reinterpret_cast<__half*>(smem_raw);                          // [16][KP] fp16 activations of the phase
    float* spart = reinterpret_cast<float*>(sx + 16 * KP);                     // [16 warps][32 lanes][4] partial accumulators
    float* sf = spart + INF_WARPS * 128;                                       // attention scratch: e[LMAX] d[D] partials[4][A]
    __shared__ int s_alive[16], s_nfr[16];                                     // every CTA tracks every sample's gate state
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * KP; i += INF_THREADS) sx[i] = __float2half_rn(0.f);     // batch rows >= B: zero forever
    if (threadIdx.x < 16) { s_alive[threadIdx.x] = threadIdx.x < p.B; s_nfr[threadIdx.x] = 0; }
    __syncthreads();

    const int nmA = (XPAD + IH) / 32, nm0 = (p.D + IH) / 32, nm1 = 2 * IH / 32, nmd = IH / 32;
    const int gw = blockIdx.x * INF_WARPS + warp, nw = gridDim.x * INF_WARPS;
    // which phases this CTA has tasks in (CTAs without a task skip the phase's staging: nothing they read would be followed
    // by a store, and the ring-slot reuse argument relies on read -> store chains)
    const bool act_lstm = static_cast<int>(blockIdx.x) * 4 < IH / 2, act_q = static_cast<int>(blockIdx.x) < p.A / 8;
    const bool act_dense = static_cast<int>(blockIdx.x) < IH / 8, act_conv = static_cast<int>(blockIdx.x) < p.M / 4;
    const bool forced = p.attn_forced != nullptr;    // `attn` given (flowtron.py:585-588): no query / score / softmax / prior
    Prefetch pf;
    mv_prefetch(pf, p.wA, IH / 2, nmA, 4);
    for (int i = 0; i < p.T; ++i) {
        const long long cur = (i % 3) * p.slot_bytes, prv = ((i + 2) % 3) * p.slot_bytes, nxt = ((i + 1) % 3) * p.slot_bytes;
        // all samples stopped?  (s_alive: the flags every CTA read in the previous frame's P4 -- identical everywhere)
        bool any = false;
        for (int b = 0; b < p.B; ++b) any |= (s_alive[b] != 0);
        if (!any) break;
        __threadfence();                         // one gpu-scope fence per thread per frame: see the header (slot reuse)

        // ---- P1 attention_lstm step on the previous output frame (zeros at i == 0): x = [out_{i-1} (80 -> 96) ; hA_{i-1}]
        IT_TRACE(0);
        if (act_lstm) {
            if (threadIdx.x < 16 * 2) *reinterpret_cast<uint4*>(sx + (threadIdx.x >> 1) * KP + p.M + 8 * (threadIdx.x & 1)) = make_uint4(0u, 0u, 0u, 0u);
            stage2(sx, 0, SL(p.x16, prv), p.M, XPAD, XPAD, SL(p.hA16, prv), IH, IH, p.B, p.status);
        }
        __syncthreads();
        IT_TRACE(1);
        mv_phase<kB16>(p.wA, IH / 2, nmA, 4, spart, sx, pf, [&](int task, float (&d)[4]) {
            lstm_epi(p, task, d, p.bA, p.cA, SL(p.hA, cur), SL(p.hA, nxt), SL(p.hA16, cur), SL(p.hA16, nxt));
        });
        if (forced) mv_prefetch(pf, p.w0, IH / 2, nm0, 4); else mv_prefetch(pf, p.wq, p.A / 8, nmd, INF_WARPS);
        IT_TRACE(2);

        if (!forced) {
            // ---- P2 query projection (no bias)
            if (act_q) stage2(sx, 0, SL(p.hA16, cur), IH, IH, 0, nullptr, 0, 0, p.B, p.status);
            __syncthreads();
            IT_TRACE(4);
            mv_phase<kB16>(p.wq, p.A / 8, nmd, INF_WARPS, spart, sx, pf, [&](int task, float (&d)[4]) {
                const int j = lane & 3, r = lane >> 2;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int b = r + 8 * half;
                    if (b < p.B && (half == 0 || kB16)) {
                        const int o = b * p.A + 8 * task + 2 * j;
                        st_rlx_f32(SL(p.q, cur) + o, d[2 * half]); st_rlx_f32(SL(p.q, cur) + o + 1, d[2 * half + 1]);
                        st_rlx_u32(SL(p.q, nxt) + o, SENT); st_rlx_u32(SL(p.q, nxt) + o + 1, SENT);
                    }
                }
            });
            mv_prefetch(pf, p.w0, IH / 2, nm0, 4);       // lstm layer 0's weights ride through the two attention phases
            IT_TRACE(5);
            // ---- P3a scores e[b,l] = v . tanh(q[b] + K[l,b]) / temperature over ALL warps of the grid (no key mask in
            //      inference, flowtron.py:800-803); every warp polls the query row it needs
            for (int t = gw; t < p.B * p.L; t += nw) {
                const int b = t / p.L, l = t - b * p.L;
                const float* kr = p.Kp + (static_cast<long long>(l) * p.B + b) * p.A;
                const float* qr = SL(p.q, cur) + b * p.A;
                float s = 0.f;
                for (int a0 = 0; a0 < p.A; a0 += 128) {                      // 4 independent loads per operand in flight
                    float kv[4]; uint32_t qv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int a = a0 + 32 * u + lane; kv[u] = a < p.A ? kr[a] : 0.f; qv[u] = a < p.A ? ld_rlx_u32(qr + a) : 0u; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int a = a0 + 32 * u + lane;
                        if (a < p.A) {
                            const float qf = qv[u] == SENT ? poll_f32(qr + a, p.status, 305) : __uint_as_float(qv[u]);
                            s = fmaf(p.v[a], tanh_f(qf + kv[u]), s);
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) { st_rlx_f32(SL(p.e, cur) + t, s * p.inv_temperature); st_rlx_u32(SL(p.e, nxt) + t, SENT); }
            }
            IT_TRACE(6);
        }
        // ---- P3b utterance b inside CTA b (attend_one)
        if (static_cast<int>(blockIdx.x) < p.B) attend_one(p, sf, i, cur, nxt, forced, s_alive[blockIdx.x] != 0);
        IT_TRACE(7);
        // ---- P4 lstm layer 0 on [d ; h0_{i-1}]; every CTA reads the samples' gate flags of this frame
        if (act_lstm) {
            stage2(sx, 0, SL(p.d16, cur), p.D, p.D, p.D, SL(p.h016, prv), IH, IH, p.B, p.status);
        }
        if (threadIdx.x < p.B) {
            const float fl = poll_f32(reinterpret_cast<const float*>(SL(p.alive, cur)) + threadIdx.x, p.status, 306);
            if (s_alive[threadIdx.x]) s_nfr[threadIdx.x] = i + 1;
            s_alive[threadIdx.x] = __float_as_uint(fl) != 0u;
        }
        __syncthreads();
        IT_TRACE(9);
        mv_phase<kB16>(p.w0, IH / 2, nm0, 4, spart, sx, pf, [&](int task, float (&d)[4]) {
            lstm_epi(p, task, d, p.b0, p.c0, nullptr, nullptr, SL(p.h016, cur), SL(p.h016, nxt));
        });
        mv_prefetch(pf, p.w1, IH / 2, nm1, 4);
        IT_TRACE(10);
        // ---- P5 lstm layer 1 on [h0 ; h1_{i-1}]
        if (act_lstm) {
            stage2(sx, 0, SL(p.h016, cur), IH, IH, IH, SL(p.h116, prv), IH, IH, p.B, p.status);
        }
        __syncthreads();
        IT_TRACE(12);
        mv_phase<kB16>(p.w1, IH / 2, nm1, 4, spart, sx, pf, [&](int task, float (&d)[4]) {
            lstm_epi(p, task, d, p.b1, p.c1, nullptr, nullptr, SL(p.h116, cur), SL(p.h116, nxt));
        });
        mv_prefetch(pf, p.wd1, IH / 8, nmd, INF_WARPS);
        IT_TRACE(13);
        // ---- P6/P7 dense layers (tanh)
        auto dense = [&](const __half* W, const float* bias, const __half* x, __half* y, __half* y_next) {
            if (act_dense) stage2(sx, 0, x, IH, IH, 0, nullptr, 0, 0, p.B, p.status);
            __syncthreads();
            mv_phase<kB16>(W, IH / 8, nmd, INF_WARPS, spart, sx, pf, [&](int task, float (&d)[4]) {
                const int j = lane & 3, r = lane >> 2;
                const float b0 = bias[task * 8 + 2 * j], b1 = bias[task * 8 + 2 * j + 1];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int b = r + 8 * half;
                    if (b < p.B && (half == 0 || kB16)) {
                        const int o = b * IH + 8 * task + 2 * j;
                        st_rlx_h2(y + o, tanh_f(d[2 * half] + b0), tanh_f(d[2 * half + 1] + b1));
                        st_rlx_u32(y_next + o, SENT);
                    }
                }
            });
        };
        dense(p.wd1, p.bd1, SL(p.h116, cur), SL(p.y116, cur), SL(p.y116, nxt));
        mv_prefetch(pf, p.wd2, IH / 8, nmd, INF_WARPS);
        IT_TRACE(15);
        dense(p.wd2, p.bd2, SL(p.y116, cur), SL(p.y216, cur), SL(p.y216, nxt));
        mv_prefetch(pf, p.wc, p.M / 4, nmd, INF_WARPS);
        IT_TRACE(17);
        // ---- P8 conv + inverse affine: out = (residual - b) / exp(log_s); task rows = (log_s, b) of 4 consecutive channels
        if (act_conv) stage2(sx, 0, SL(p.y216, cur), IH, IH, 0, nullptr, 0, 0, p.B, p.status);
        __syncthreads();
        IT_TRACE(18);
        mv_phase<kB16>(p.wc, p.M / 4, nmd, INF_WARPS, spart, sx, pf, [&](int task, float (&d)[4]) {
            const int j = lane & 3, r = lane >> 2;
            const int m = 4 * task + j;
            const float bl = p.bc[task * 8 + 2 * j], bb = p.bc[task * 8 + 2 * j + 1];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int b = r + 8 * half;
                if (b < p.B && (half == 0 || kB16)) {
                    const long long ro = (static_cast<long long>(i) * p.B + b) * p.M;
                    // a sample that stopped at an earlier frame emits zeros; the frame that trips the gate is still emitted
                    // (s_nfr: updated in this frame's P4 from the flags every CTA read)
                    const bool emit = (i < s_nfr[b]);
                    const float val = (p.residual[ro + m] - (d[2 * half + 1] + bb)) / expf(d[2 * half] + bl);
                    p.out[ro + m] = emit ? val : 0.f;
                    st_rlx_h(SL(p.x16, cur) + b * XPAD + m, val);
                    st_rlx_u16(SL(p.x16, nxt) + b * XPAD + m, 0xFFFFu);
                }
            }
        });
        mv_prefetch(pf, p.wA, IH / 2, nmA, 4);           // next frame's attention LSTM
        IT_TRACE(19);
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
// dst[task][m][n][kk] (fp16) <- W[row(task, n)][32 m + kk] over the virtual K = [A part (Ka real columns, padded to Kap) ; B part]
//   kind 0: LSTM      row = (n % 4) * 1024 + 2 task + n / 4        (rows i0,f0,g0,o0,i1,f1,g1,o1)
//   kind 1: dense     row = 8 task + n
//   kind 2: conv pair row = (n % 2) * M + 4 task + n / 2           (log_s, b of 4 consecutive channels)
// bias_dst[task][n] = b1[row] (+ b2[row])
__global__ void pack_weights_kernel(__half* __restrict__ dst, float* __restrict__ bias_dst, const float* __restrict__ WA, int Ka, int Kap,
                                    const float* __restrict__ WB, int Kb, const float* __restrict__ b1, const float* __restrict__ b2,
                                    int n_tasks, int nm, int kind, int M) {
    const long long total = static_cast<long long>(n_tasks) * nm * 256;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int kk = static_cast<int>(i & 31), n = static_cast<int>((i >> 5) & 7);
        const long long tm = i >> 8;
        const int m = static_cast<int>(tm % nm), task = static_cast<int>(tm / nm);
        const int row = kind == 0 ? (n & 3) * IH + 2 * task + (n >> 2) : kind == 1 ? 8 * task + n : (n & 1) * M + 4 * task + (n >> 1);
        const int k = 32 * m + kk;
        float v = 0.f;
        if (k < Kap) { if (k < Ka) v = WA[static_cast<long long>(row) * Ka + k]; }
        else if (WB && k - Kap < Kb) v = WB[static_cast<long long>(row) * Kb + (k - Kap)];
        dst[i] = __float2half_rn(v);
        if (m == 0 && kk == 0 && bias_dst) bias_dst[task * 8 + n] = (b1 ? b1[row] : 0.f) + (b2 ? b2[row] : 0.f);
    }
}

static int pack(__half* dst, float* bias_dst, const float* WA, int Ka, int Kap, const float* WB, int Kb, const float* b1, const float* b2,
                int n_tasks, int nm, int kind, int M, cudaStream_t st) {
    const long long total = static_cast<long long>(n_tasks) * nm * 256;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_weights_kernel<<<blocks, 256, 0, st>>>(dst, bias_dst, WA, Ka, Kap, WB, Kb, b1, b2, n_tasks, nm, kind, M);
    ft_count_launch(1);
    return ft_check_launch("pack_weights_kernel");
}

// ------------------------------------------------------------------------------------------------ host
struct InferScratch {
    __half *wA, *w0, *w1, *wq, *wd1, *wd2, *wc;
    float *bA, *b0, *b1, *bd1, *bd2, *bc;
    uint16_t *wk, *wv, *text16;
    float *Kp, *Vp, *state;
    int* ints;
    size_t state_floats, total;
};

static InferScratch plan_infer(const FtArStepDesc& d, uint8_t* base) {
    InferScratch s;
    size_t off = 0;
    auto get = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return base ? base + o : nullptr; };
    const int M = d.n_mel, A = d.n_attn, E = d.n_text, D = IH + A;
    auto wbytes = [](size_t tasks, size_t nm) { return tasks * nm * 256 * 2; };
    s.wA = reinterpret_cast<__half*>(get(wbytes(IH / 2, (XPAD + IH) / 32)));
    s.w0 = reinterpret_cast<__half*>(get(wbytes(IH / 2, (D + IH) / 32)));
    s.w1 = reinterpret_cast<__half*>(get(wbytes(IH / 2, 2 * IH / 32)));
    s.wq = reinterpret_cast<__half*>(get(wbytes(A / 8, IH / 32)));
    s.wd1 = reinterpret_cast<__half*>(get(wbytes(IH / 8, IH / 32)));
    s.wd2 = reinterpret_cast<__half*>(get(wbytes(IH / 8, IH / 32)));
    s.wc = reinterpret_cast<__half*>(get(wbytes(M / 4, IH / 32)));
    s.bA = reinterpret_cast<float*>(get(IG * 4)); s.b0 = reinterpret_cast<float*>(get(IG * 4)); s.b1 = reinterpret_cast<float*>(get(IG * 4));
    s.bd1 = reinterpret_cast<float*>(get(IH * 4)); s.bd2 = reinterpret_cast<float*>(get(IH * 4)); s.bc = reinterpret_cast<float*>(get(2 * M * 4));
    s.wk = reinterpret_cast<uint16_t*>(get(size_t(A) * E * 2));
    s.wv = reinterpret_cast<uint16_t*>(get(size_t(A) * E * 2));
    s.text16 = reinterpret_cast<uint16_t*>(get(size_t(d.L) * d.B * E * 2));
    s.Kp = reinterpret_cast<float*>(get(size_t(d.L) * d.B * A * 4));
    s.Vp = reinterpret_cast<float*>(get(size_t(d.L) * d.B * A * 4));
    // state: cell states cA, c0, c1 (3 x B*IH fp32), then the exchange rings (3 frame slots each): hA32 (B*IH), q (B*A), e (B*L),
    // alive (B) as 32-bit words; hA16, h016, h116, y116, y216 (B*IH), d16 (B*D), x16 (B*96) as fp16 (counted in floats, each piece
    // padded to 16 bytes)
    s.state_floats = size_t(d.B) * 3 * IH + 3 * (size_t(d.B) * (IH + A + d.L + 1) + 16) +
                     3 * ((size_t(d.B) * (5 * IH + D + XPAD) + 1) / 2 + 32) + 256;
    s.state = reinterpret_cast<float*>(get(s.state_floats * 4));
    s.ints = nullptr;
    s.total = off;
    return s;
}

}  // namespace ft

extern "C" {

/* debug: device buffer [T][32] receiving clock64 stamps of CTA 0 of the next ft_ar_step_infer launches (NULL disables) */
void ft_debug_set_infer_trace(long long* buf) { ft::g_infer_trace = buf; }

size_t ft_ar_step_infer_scratch_bytes(const FtArStepDesc* d) { return ft::plan_infer(*d, nullptr).total + 256; }

int ft_ar_step_infer(const FtArStepDesc* d, const FtArStepWeights* w, const float* residual, const float* text,
                     const float* attn_prior, const float* attn_forced, float gate_threshold, float* out, float* attn_out,
                     int* n_frames, void* scratch, void* stream) {
    using namespace ft;
    if (!d || !w || !residual || !text || !out || !attn_out || !n_frames || !scratch) return ft_set_error("ft_ar_step_infer: NULL argument");
    if (d->n_hidden != IH) return ft_set_error("infer: n_hidden must be 1024");
    if (d->L > LMAX) return ft_set_error("infer: L > 256 not supported");
    if (d->B > 16) return ft_set_error("infer: batch > 16 per call not supported (run the batch in slices of 16)");
    if (d->n_mel > XPAD || d->n_mel % 8 || d->n_attn % 32 || d->n_text % 8 || d->n_attn > 4 * INF_THREADS)
        return ft_set_error("infer: n_mel <= 96 and %8, n_attn %32 and <= 2048, n_text %8 required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    InferScratch s = plan_infer(*d, static_cast<uint8_t*>(scratch));
    const int M = d->n_mel, A = d->n_attn, E = d->n_text, D = IH + A, B = d->B;
    const long long RL = static_cast<long long>(d->L) * B;
#define FT_TRYX(x) do { if ((x) != 0) return -1; } while (0)
    // weights in consumption order (fp16), LSTM input + recurrent matrices fused along K, biases pre-added
    FT_TRYX(pack(s.wA, s.bA, w->attn_lstm_w_ih, M, XPAD, w->attn_lstm_w_hh, IH, w->attn_lstm_b_ih, w->attn_lstm_b_hh, IH / 2, (XPAD + IH) / 32, 0, M, st));
    FT_TRYX(pack(s.w0, s.b0, w->lstm_w_ih0, D, D, w->lstm_w_hh0, IH, w->lstm_b_ih0, w->lstm_b_hh0, IH / 2, (D + IH) / 32, 0, M, st));
    FT_TRYX(pack(s.w1, s.b1, w->lstm_w_ih1, IH, IH, w->lstm_w_hh1, IH, w->lstm_b_ih1, w->lstm_b_hh1, IH / 2, 2 * IH / 32, 0, M, st));
    FT_TRYX(pack(s.wq, nullptr, w->att_query, IH, IH, nullptr, 0, nullptr, nullptr, A / 8, IH / 32, 1, M, st));
    FT_TRYX(pack(s.wd1, s.bd1, w->dense_w0, IH, IH, nullptr, 0, w->dense_b0, nullptr, IH / 8, IH / 32, 1, M, st));
    FT_TRYX(pack(s.wd2, s.bd2, w->dense_w1, IH, IH, nullptr, 0, w->dense_b1, nullptr, IH / 8, IH / 32, 1, M, st));
    FT_TRYX(pack(s.wc, s.bc, w->conv_w, IH, IH, nullptr, 0, w->conv_b, nullptr, M / 4, IH / 32, 2, M, st));
    FT_TRYX(launch_cast(w->att_key, 2, s.wk, 0, static_cast<long long>(A) * E, st));
    FT_TRYX(launch_cast(w->att_value, 2, s.wv, 0, static_cast<long long>(A) * E, st));
    FT_TRYX(launch_cast(text, 2, s.text16, 0, RL * E, st));
    {   // K / V projections once per utterance (the reference recomputes them every frame, flowtron.py:568-570)
        GemmArgs g;
        g.M = static_cast<int>(RL); g.N = A; g.K = E; g.A = s.text16; g.lda = E; g.a_fmt = FMT_F16;
        g.B = s.wk; g.ldb = E; g.b_fmt = FMT_F16; g.C32 = s.Kp; g.ldc32 = A;
        FT_TRYX(launch_gemm(g, st));
        g.B = s.wv; g.C32 = s.Vp;
        FT_TRYX(launch_gemm(g, st));
    }
    if (cudaMemsetAsync(out, 0, sizeof(float) * d->T * B * M, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(attn_out, 0, sizeof(float) * d->T * B * d->L, st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(n_frames, 0, sizeof(int) * B, st) != cudaSuccess) return ft_set_error("infer: memset failed");

    InferParams p;
    p.T = d->T; p.B = B; p.L = d->L; p.M = M; p.A = A; p.E = E; p.D = D;
    p.wA = s.wA; p.w0 = s.w0; p.w1 = s.w1; p.wq = s.wq; p.wd1 = s.wd1; p.wd2 = s.wd2; p.wc = s.wc;
    p.bA = s.bA; p.b0 = s.b0; p.b1 = s.b1; p.bd1 = s.bd1; p.bd2 = s.bd2; p.bc = s.bc;
    p.v = w->att_v; p.wg = w->gate_w; p.bg = w->gate_b;
    p.Kp = s.Kp; p.Vp = s.Vp; p.residual = residual; p.prior = d->has_prior ? attn_prior : nullptr;
    p.attn_forced = attn_forced;
    p.inv_temperature = 1.0f / d->temperature; p.gate_threshold = gate_threshold;
    p.has_gate = d->has_gate && w->gate_w;
    p.out = out; p.attn_out = attn_out; p.n_frames = n_frames;
    float* f = s.state;
    auto take = [&](size_t n) { float* r = f; f += (n + 3) & ~size_t(3); return r; };     // 16-byte aligned pieces (16-byte polls)
    const size_t BH = static_cast<size_t>(B) * IH;
    auto take16 = [&](size_t n) { return reinterpret_cast<__half*>(take((n + 1) / 2)); };
    p.cA = take(BH); p.c0 = take(BH); p.c1 = take(BH);
    float* ring0 = f;                                            // everything from here on is exchanged: sentinel-filled
    p.hA = take(BH); p.q = take(static_cast<size_t>(B) * A); p.e = take(static_cast<size_t>(B) * d->L);
    p.alive = reinterpret_cast<uint32_t*>(take(B));
    p.hA16 = take16(BH); p.h016 = take16(BH); p.h116 = take16(BH); p.y116 = take16(BH); p.y216 = take16(BH);
    p.d16 = take16(static_cast<size_t>(B) * D); p.x16 = take16(static_cast<size_t>(B) * XPAD);
    p.slot_bytes = (f - ring0) * static_cast<long long>(sizeof(float));
    f = ring0 + 3 * (f - ring0);
    if (static_cast<size_t>(f - s.state) > s.state_floats) return ft_set_error("infer: scratch plan overflow");
    if (cudaMemsetAsync(s.state, 0, (ring0 - s.state) * sizeof(float), st) != cudaSuccess) return ft_set_error("infer: memset failed");
    if (cudaMemsetAsync(ring0, 0xFF, (f - ring0) * sizeof(float), st) != cudaSuccess) return ft_set_error("infer: memset failed");
    // the state before frame 0 (slot 2 = frame -1): zero hidden states and a zero previous output frame
    auto slot2 = [&](void* ptr) { return static_cast<char*>(ptr) + 2 * p.slot_bytes; };
    if (cudaMemsetAsync(slot2(p.hA16), 0, BH * 2, st) != cudaSuccess || cudaMemsetAsync(slot2(p.h016), 0, BH * 2, st) != cudaSuccess ||
        cudaMemsetAsync(slot2(p.h116), 0, BH * 2, st) != cudaSuccess || cudaMemsetAsync(slot2(p.x16), 0, static_cast<size_t>(B) * XPAD * 2, st) != cudaSuccess)
        return ft_set_error("infer: memset failed");
    p.status = ft_status_word();
    p.trace = g_infer_trace;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < B) return ft_set_error("infer: fewer SMs than utterances");
    void* fn = B > 8 ? reinterpret_cast<void*>(infer_kernel<true>) : reinterpret_cast<void*>(infer_kernel<false>);
    const int smem = 16 * KP * 2 + INF_WARPS * 128 * 4 + (LMAX + D + 4 * A + 64) * 4;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    TimeScope ts("infer", d->T, B, d->L, st);
    void* args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(INF_THREADS), args, smem, st);
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("infer_kernel");
}

}  // extern "C"
